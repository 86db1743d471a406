"""Known-answer tests of the numpy restatement of CoordLSVotingWeighted (voting_layers_2d.py:5-122)."""
import numpy as np

from oracle import ls_voting_np as L

F = np.float32


def _scene(h=40, w=56, kps=((12.3, 30.6), (25.1, 9.4)), boxes=((5, 20, 8, 30),)):
    """seg logits with one class on `boxes` (y0,y1,x0,x1); perfect unit vectors to `kps` (y, x)."""
    nk = len(kps)
    seg = np.zeros((1, h, w, 2), F)
    seg[..., 0] = 1.0
    for (y0, y1, x0, x1) in boxes:
        seg[0, y0:y1, x0:x1, 1] = 3.0
    ys, xs = np.mgrid[0:h, 0:w]
    direct = np.zeros((1, h, w, 2 * nk), F)
    for k, (ky, kx) in enumerate(kps):
        dy, dx = ky - (ys + 0.5), kx - (xs + 0.5)
        n = np.sqrt(dy * dy + dx * dx)
        direct[0, :, :, 2 * k], direct[0, :, :, 2 * k + 1] = dy / n, dx / n
    conf = np.zeros((1, h, w, nk), F)
    return seg, direct, conf


def test_perfect_field_recovers_keypoints_in_yx_order():
    seg, direct, conf = _scene()
    out = L.coord_ls_voting_weighted(seg, direct, conf, num_points=2)
    assert out.shape == (1, 1, 2, 2) and out.dtype == F
    assert np.abs(out[0, 0] - np.array([(12.3, 30.6), (25.1, 9.4)], F)).max() < 1e-3  # (y, x), pixels


def test_both_grid_axes_are_divided_by_the_height_and_scaled_back_by_it():
    # non-square image: a bug that divided x by the width would move the x estimate by w/h
    seg, direct, conf = _scene(h=32, w=96, kps=((10.5, 70.25),), boxes=((4, 28, 10, 60),))
    out = L.coord_ls_voting_weighted(seg, direct, conf, num_points=1)
    assert np.abs(out[0, 0, 0] - np.array((10.5, 70.25), F)).max() < 1e-3


def test_softplus_and_sigmoid_weights():
    x = np.array([-20.0, -1.0, 0.0, 1.0, 20.0], F)
    assert np.allclose(L.softplus_f32(x), np.log1p(np.exp(x.astype(np.float64))), rtol=1e-6)
    assert np.allclose(L.sigmoid_f32(x), 1 / (1 + np.exp(-x.astype(np.float64))), rtol=1e-6)
    seg, direct, conf = _scene()
    rng = np.random.default_rng(0)
    conf = rng.normal(size=conf.shape).astype(F)
    a = L.coord_ls_voting_weighted(seg, direct, conf, num_points=2)
    b = L.coord_ls_voting_weighted(seg, direct, conf, num_points=2, sigmoid_weights=True)
    assert np.abs(a - b).max() < 1e-3  # perfect field: the weights do not move the solution


def test_hard_softmax_is_one_hot_except_at_near_ties():
    seg = np.array([[[[0.3, 0.1, 0.2], [0.5, 0.5 + 1e-7, 0.1]]]], F)
    hot = L.hard_softmax_f32(seg)
    assert np.array_equal(hot[0, 0, 0], np.array([1, 0, 0], F))
    assert 0.4 < hot[0, 0, 1, 0] < 0.6 and 0.4 < hot[0, 0, 1, 1] < 0.6  # 1e6 * 1e-7 = 0.1 apart


def test_component_filter_keeps_the_largest_component_and_its_quirks():
    big, small = (5, 25, 8, 30), (30, 36, 40, 50)  # 440 px and 60 px
    seg, direct, conf = _scene(boxes=(big, small))
    # without the filter both blobs vote; break the small blob's vectors to make that visible
    direct[0, 30:36, 40:50, :] = np.array([0.0, 1.0, 1.0, 0.0], F)
    free = L.coord_ls_voting_weighted(seg, direct, conf, num_points=2)
    filt, dbg = L.coord_ls_voting_weighted(seg, direct, conf, num_points=2, filter_estimates=True, return_debug=True)
    assert np.abs(filt[0, 0] - np.array([(12.3, 30.6), (25.1, 9.4)], F)).max() < 1e-3
    assert np.abs(free - filt).max() > 0.05
    assert dbg["hot"][0, :, :, 0].sum() == 440
    # second largest component on request (:58-73)
    second, dbg2 = L.coord_ls_voting_weighted(seg, direct, conf, num_points=2, filter_estimates=True,
                                              output_second_largest_component=True, return_debug=True)
    assert dbg2["hot"][0, :, :, 0].sum() == 60
    # quirk 1: every component below 50 px -> counts zeroed -> top_k ties pick label 1, the first blob in raster order
    seg3, direct3, conf3 = _scene(boxes=((2, 6, 40, 46), (20, 26, 10, 17)))  # 24 px (first), 42 px
    _, dbg3 = L.coord_ls_voting_weighted(seg3, direct3, conf3, num_points=2, filter_estimates=True, return_debug=True)
    assert dbg3["hot"][0, :, :, 0].sum() == 24
    # quirk 2: the object is larger than the background -> index[1] is label 0 -> nothing of the class survives
    seg4, direct4, conf4 = _scene(boxes=((0, 40, 0, 40),))  # 1600 of 2240 px
    out4, dbg4 = L.coord_ls_voting_weighted(seg4, direct4, conf4, num_points=2, filter_estimates=True, return_debug=True)
    assert dbg4["hot"].sum() == 0 and np.array_equal(out4, np.zeros_like(out4))


def test_empty_class_gives_zeros():
    seg, direct, conf = _scene(boxes=())
    out = L.coord_ls_voting_weighted(seg, direct, conf, num_points=2, filter_estimates=True)
    assert np.array_equal(out, np.zeros((1, 1, 2, 2), F))


def test_parallel_vectors_use_the_pseudo_inverse():
    seg, direct, conf = _scene(kps=((12.3, 30.6),))
    direct[..., 0], direct[..., 1] = 0.0, 1.0  # every vector along +x: R = diag(1, 0) * w, rank one
    out = L.coord_ls_voting_weighted(seg, direct, conf, num_points=1)
    ys = np.arange(5, 20) + 0.5
    assert abs(out[0, 0, 0, 0] - ys.mean()) < 1e-3 and out[0, 0, 0, 1] == 0.0  # y = mean row, x unconstrained -> 0
