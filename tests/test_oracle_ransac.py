"""Known-answer tests of the numpy oracle (the reference ships none — SURVEY.md section 4 — so every
degenerate branch listed in section 8a gets a hand-computed case here)."""
import math

import numpy as np
import pytest

from oracle import philox_np
from oracle import ransac_voting_np as O

F = np.float32


def test_generate_hypothesis_hand_case():
    # pixel A (0.5,0.5) looking along +x, pixel B (2.5,2.5) looking along -y -> they meet at (2.5, 0.5)
    coords = np.array([[0.5, 0.5], [2.5, 2.5]], F)
    direct = np.array([[[1, 0]], [[0, -1]]], F)  # [tn=2, vn=1, (dx,dy)]
    hyp = O.generate_hypothesis(direct, coords, np.array([[[0, 1]]], np.int32))
    assert np.array_equal(hyp, np.array([[[2.5, 0.5]]], F))
    # det = d1x*d0y - d1y*d0x = 0*0 - (-1)*1 = 1, u = ((2)*0 - (2)*(-1))/1 = 2


def test_generate_hypothesis_degenerate_pairs_give_origin():
    coords = np.array([[0.5, 0.5], [2.5, 2.5], [4.5, 0.5]], F)
    direct = np.array([[[1, 0]], [[0, -1]], [[2, 0]]], F)
    same = O.generate_hypothesis(direct, coords, np.array([[[1, 1]]], np.int32))  # same pixel twice -> det 0
    par = O.generate_hypothesis(direct, coords, np.array([[[0, 2]]], np.int32))  # parallel vectors -> det 0
    assert np.array_equal(same, np.zeros((1, 1, 2), F))
    assert np.array_equal(par, np.zeros((1, 1, 2), F))


def test_voting_angle_threshold_and_guards():
    coords = np.array([[0.5, 0.5]], F)
    hyp = np.array([[[10.5, 0.5]]], F)  # 10 px to the right

    def vote(ang_deg, hyp=hyp, scale=1.0):
        a = math.radians(ang_deg)
        direct = np.array([[[scale * math.cos(a), scale * math.sin(a)]]], F)
        return int(O.voting_for_hypothesis(direct, coords, hyp, 0.99)[0, 0, 0])

    assert vote(0.0) == 1
    assert vote(8.0) == 1  # cos 8.0 deg = 0.99027
    assert vote(-8.0) == 1
    assert vote(8.2) == 0  # cos 8.2 deg = 0.98978
    assert vote(180.0) == 0
    assert vote(3.0, scale=1e-7) == 0  # |direct| <= 1e-6            :240
    assert vote(0.0, scale=0.0) == 0  # zero vector at a masked pixel
    assert vote(0.0, hyp=np.array([[[0.5, 0.5]]], F)) == 0  # hypothesis on the pixel, |hd| = 0   :240
    assert vote(-45.0, hyp=np.array([[[3.0, -3.0]]], F)) == 0  # |hx + hy| <= 1e-6 guard         :241-243
    assert vote(0.0, hyp=np.zeros((1, 1, 2), F)) == 0  # the (0,0) hypothesis of a degenerate pair


def _disk(h, w, cx, cy, r):
    ys, xs = np.mgrid[0:h, 0:w]
    return (((xs + 0.5 - cx) ** 2 + (ys + 0.5 - cy) ** 2) <= r * r).astype(F)


def _field(h, w, kps):
    ys, xs = np.mgrid[0:h, 0:w]
    v = np.zeros((h, w, len(kps), 2), F)
    for k, (kx, ky) in enumerate(kps):
        dx, dy = kx - (xs + 0.5), ky - (ys + 0.5)
        n = np.sqrt(dx * dx + dy * dy) + 1e-12
        v[:, :, k, 0] = dy / n
        v[:, :, k, 1] = dx / n
    return v


KPS = [(20.3 + 3.1 * k, 31.7 - 2.3 * k) for k in range(9)]


def test_planted_keypoints_are_recovered():
    h, w = 48, 64
    mask = _disk(h, w, 30, 22, 12)
    r = O.ransac_voting_batch(mask, _field(h, w, KPS), 0.99, 0.99, 20, 5, 30000, 64, 9, seed=3)
    assert r["tn"] == int(mask.sum()) and r["rounds"] == 1 and r["refined"]
    assert np.abs(r["points"] - np.array(KPS, F)).max() < 1e-2  # returns (x, y)
    assert r["counts"][0].shape == (64, 9) and r["counts"][0].dtype == np.int32


def test_gate_too_few_pixels_gives_zeros():
    h, w = 16, 16
    mask = np.zeros((h, w), F)
    mask[3, 3:7] = 1  # 4 pixels < min_num 5
    r = O.ransac_voting_batch(mask, _field(h, w, KPS), 0.99, 0.99, 20, 5, 30000, 16, 9)
    assert r["tn0"] == 4 and r["tn"] == 0 and r["rounds"] == 0
    assert np.array_equal(r["points"], np.zeros((9, 2), F))
    mask[3, 7] = 1  # exactly min_num is NOT less than min_num -> runs
    r = O.ransac_voting_batch(mask, _field(h, w, KPS), 0.99, 0.99, 2, 5, 30000, 16, 9)
    assert r["tn"] == 5 and r["rounds"] >= 1


def test_cap_downsamples_with_the_selection_stream():
    h, w = 48, 64
    mask = _disk(h, w, 30, 22, 14)
    n0 = int(mask.sum())
    r = O.ransac_voting_batch(mask, _field(h, w, KPS), 0.99, 0.99, 20, 5, 200, 32, 9, seed=5, image=2, cls=1)
    sel = philox_np.draw_selection(5, 2, 1, h, w)
    keep = (sel < F(200) / F(n0)) & (mask != 0)
    assert r["tn0"] == n0 and r["tn"] == int(keep.sum()) and 100 < r["tn"] < 300
    # caller-supplied selection: keep everything below 0.5
    sel2 = np.full((h, w), 0.9, F)
    sel2[:, :32] = 0.0
    r2 = O.ransac_voting_batch(mask, _field(h, w, KPS), 0.99, 0.99, 20, 5, 200, 32, 9, selection=sel2)
    assert r2["tn"] == int((mask[:, :32] != 0).sum())


def test_argmax_ties_take_the_lowest_hypothesis_and_earlier_round_wins():
    h, w = 24, 24
    mask = _disk(h, w, 12, 12, 6)
    tn = int(mask.sum())
    vf = _field(h, w, [(12.2, 40.3)])
    # identical hypotheses in every slot and in both rounds: all counts tie
    idx = np.zeros((2, 8, 1, 2), np.int32)
    idx[..., 1] = tn - 1
    r = O.ransac_voting_batch(mask, vf, 0.99, 1.1, 2, 5, 30000, 8, 1, idxs=idx)  # confidence 1.1 -> never stops early
    assert r["rounds"] == 2
    assert all(int(wi[0]) == 0 for wi in r["win_idx"])
    assert np.array_equal(r["counts"][0], r["counts"][1])
    assert np.array_equal(r["win_pts"][0], r["hyps"][0][0, 0])


def test_singular_normal_matrix_returns_all_winners_unrefined():
    h, w = 32, 32
    mask = np.zeros((h, w), F)
    mask[10, 4:28] = 1  # one pixel row
    vf = _field(h, w, KPS[:2])
    # keypoint 1: every vector points along +x -> all normals parallel -> ATA singular
    vf[:, :, 1, 0] = 0.0
    vf[:, :, 1, 1] = 1.0
    r = O.ransac_voting_batch(mask, vf, 0.99, 0.99, 3, 5, 30000, 32, 2, seed=1)
    assert not r["refined"]
    assert np.array_equal(r["points"], r["win_pts"])  # BOTH keypoints unrefined (:364-365)
    assert not r["invertible"][1]


def test_zero_vectors_never_vote_and_force_max_iter():
    h, w = 24, 24
    mask = _disk(h, w, 12, 12, 5)
    vf = _field(h, w, KPS[:2])
    vf[:, :, 1] = 0.0  # keypoint 1 has no direction anywhere
    r = O.ransac_voting_batch(mask, vf, 0.99, 0.99, 4, 5, 30000, 16, 2, seed=2)
    assert r["rounds"] == 4  # min ratio stays 0 -> the stop test never fires
    assert all(int(c[:, 1].max()) == 0 for c in r["counts"])
    assert np.array_equal(r["win_pts"][1], np.zeros(2, F))


def test_stop_test_hand_values():
    assert O.stop_test(F(0.1), 512, 0.99)  # 1 - 0.99^512 = 0.99418
    assert not O.stop_test(F(0.09), 512, 0.99)  # 1 - 0.9919^512 = 0.98445
    assert O.stop_test(F(0.09), 1024, 0.99)
    assert not O.stop_test(F(0.0), 512 * 20, 0.99)
    assert O.stop_test(F(1.0), 512, 0.99)
    for base, n in [(0.9919, 512), (0.5, 37), (0.999999, 10240)]:
        assert abs(O._ipow_f64(base, n) - math.pow(base, n)) <= 1e-12 * math.pow(base, n)


def test_condition_number_closed_form():
    assert O.is_invertible(2.0, 0.0, 1.0)
    assert not O.is_invertible(1.0, 1.0, 1.0)  # rank 1
    assert not O.is_invertible(0.0, 0.0, 0.0)
    assert not O.is_invertible(1.0, 0.0, 1e-7)  # cond 1e7 > 1e6
    assert O.is_invertible(1.0, 0.0, 1e-5)
    a = np.array([[3.0, 1.0], [1.0, 2.0]])
    assert abs(O.cond_2x2_sym_f64(3.0, 1.0, 2.0) - np.linalg.cond(a)) < 1e-12


def test_float32_vs_float64_accumulation_differs_below_tolerance():
    h, w = 48, 64
    mask = _disk(h, w, 30, 22, 12)
    vf = _field(h, w, KPS)
    r64 = O.ransac_voting_batch(mask, vf, 0.99, 0.99, 20, 5, 30000, 64, 9, seed=3)
    r32 = O.ransac_voting_batch(mask, vf, 0.99, 0.99, 20, 5, 30000, 64, 9, seed=3, accumulate="float32")
    assert np.abs(r64["points"] - r32["points"]).max() < 1e-3


def test_all_masks_layer_shapes_and_image_offset():
    from casapose_b200 import synthetic

    d = synthetic.make_frames(2, 60, 80, (1, 5), variant="easy")
    out = O.ransac_voting_layer_all_masks(d["mask"], d["vertex"], 32, seed=11)
    assert out.shape == (2, 2, 9, 2) and out.dtype == F
    # image 1 of the batch == image 0 of a batch that starts at offset 1 (sharding invariance)
    out1 = O.ransac_voting_layer_all_masks(d["mask"][1:], d["vertex"][1:], 32, seed=11, image_offset=1)
    assert np.array_equal(out[1], out1[0])
