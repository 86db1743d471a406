"""GPU parity tests: the sm_100a path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): vote counts, winning-hypothesis indices, tn and rounds bit-exact;
refined keypoints within 1e-3 px."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from casapose_b200 import synthetic  # noqa: E402
from oracle import philox_np  # noqa: E402
from oracle import ransac_voting_np as O  # noqa: E402

pytestmark = pytest.mark.gpu
TOL_PX = 1e-3


@pytest.fixture(scope="module")
def vote(cuda_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from casapose_b200.pose_estimation import ransac_voting_layer_all_masks

    return ransac_voting_layer_all_masks


def _compare(vote, mask, vertex, hn, seed=7, okw=None, **kw):
    okw = dict(okw or {})
    pts, dbg = vote(torch.from_numpy(mask).cuda(), torch.from_numpy(vertex).cuda(), hn, seed=seed, return_debug=True, **kw)
    torch.cuda.synchronize()
    for k in ("inlier_thresh", "confidence", "max_iter", "min_num", "max_num", "image_offset"):
        if k in kw:
            okw[k] = kw[k]
    ref, rdbg = O.ransac_voting_layer_all_masks(mask, vertex, hn, seed=seed, return_debug=True, **okw)
    b, oc = mask.shape[0], mask.shape[3]
    for i in range(b):
        for c in range(oc):
            r = rdbg[i][c]
            assert int(dbg["tn0"][i, c]) == r["tn0"]
            assert int(dbg["tn"][i, c]) == r["tn"], (i, c)
            assert int(dbg["rounds"][i, c]) == r["rounds"], (i, c)
            for k in range(r["rounds"]):
                assert np.array_equal(dbg["counts"][i, c, k].cpu().numpy(), r["counts"][k]), ("counts", i, c, k)
                assert np.array_equal(dbg["win_idx"][i, c, k].cpu().numpy(), r["win_idx"][k]), ("win_idx", i, c, k)
            assert np.array_equal(dbg["win_pts"][i, c].cpu().numpy(), r["win_pts"]), ("win_pts", i, c)
            assert int(dbg["refined"][i, c]) == int(r["refined"])
    err = float(np.abs(pts.cpu().numpy() - ref).max())
    assert err <= TOL_PX, "refined keypoints differ by %g px" % err
    return pts, dbg, ref, rdbg


@pytest.mark.parametrize("variant", ["easy", "hard", "clean"])
def test_votes_bit_exact_small(vote, variant):
    d = synthetic.make_frames(2, 120, 160, (1, 5, 6), variant=variant)
    _, dbg, _, rdbg = _compare(vote, d["mask"], d["vertex"], 64, max_iter=6 if variant == "hard" else 20)
    if variant == "hard":
        assert max(r["rounds"] for row in rdbg for r in row) > 1, "the hard variant must exercise several rounds"


def test_filter_equals_exact_predicate(vote):
    d = synthetic.make_frames(1, 120, 160, (1, 5, 6), variant="easy")
    m, v = torch.from_numpy(d["mask"]).cuda(), torch.from_numpy(d["vertex"]).cuda()
    p1, d1 = vote(m, v, 96, seed=3, return_debug=True)
    p2, d2 = vote(m, v, 96, seed=3, return_debug=True, force_exact=True)
    assert torch.equal(d1["counts"], d2["counts"]) and torch.equal(d1["win_idx"], d2["win_idx"])
    assert torch.equal(p1, p2)
    assert int(d2["stats"][1]) == int(d2["stats"][0])  # force_exact sends every unit to the exact predicate
    assert int(d1["stats"][1]) < int(d1["stats"][0]) // 1000


def test_config3_shape_13_objects(vote):
    """BASELINE config 3 (13 LM objects), reduced in size so the oracle finishes in seconds."""
    d = synthetic.make_frames(1, 240, 320, synthetic.CONFIG_13_IDS, variant="easy")
    _compare(vote, d["mask"], d["vertex"], 128)


def test_full_resolution_frame_512_hypotheses(vote):
    """One 480x640, 8-object frame at the reference's hn=512 (BASELINE config 2 shape, b=1)."""
    d = synthetic.make_frames(1, 480, 640, synthetic.CONFIG_8_IDS, variant="easy")
    _compare(vote, d["mask"], d["vertex"], 512)


def test_cap_downsampling_philox_and_external_selection(vote):
    d = synthetic.make_frames(1, 240, 320, synthetic.CONFIG_8_IDS, variant="easy")
    _, dbg, _, _ = _compare(vote, d["mask"], d["vertex"], 64, max_num=400)
    assert int(dbg["tn"].max()) < int(dbg["tn0"].max())
    rng = np.random.default_rng(0)
    sel = rng.random((1, 8, 240, 320), dtype=np.float32)
    _compare(vote, d["mask"], d["vertex"], 64, max_num=400, selection=torch.from_numpy(sel).cuda(), okw={"selection": sel})


def test_external_idxs(vote):
    d = synthetic.make_frames(1, 120, 160, (1, 5), variant="hard")
    tn = d["mask"].sum((1, 2)).astype(np.int64)[0]
    hn, mi = 48, 5
    idxs = np.zeros((1, 2, mi, hn, 9, 2), np.int32)
    rng = np.random.default_rng(5)
    for c in range(2):
        idxs[0, c] = rng.integers(0, max(int(tn[c]), 1), size=(mi, hn, 9, 2), dtype=np.int32)
    _compare(vote, d["mask"], d["vertex"], hn, max_iter=mi, idxs=torch.from_numpy(idxs).cuda(), okw={"idxs": idxs})


def test_image_offset_makes_sharding_invisible(vote):
    d = synthetic.make_frames(3, 96, 128, (1, 5), variant="easy")
    m, v = torch.from_numpy(d["mask"]).cuda(), torch.from_numpy(d["vertex"]).cuda()
    full = vote(m, v, 64, seed=9)
    part = vote(m[2:].contiguous(), v[2:].contiguous(), 64, seed=9, image_offset=2)
    assert torch.equal(full[2], part[0])


@pytest.mark.parametrize("variant", ["easy", "hard"])
def test_graph_replay_equals_debug_path_and_oracle(vote, variant):
    """Calls without debug outputs replay a cached CUDA graph whose kernel nodes are patched per call; calls with
    debug outputs launch directly.  Same bits either way, in the multi-round case too, for repeated calls and for
    a shifted / split batch."""
    d = synthetic.make_frames(3, 120, 160, (1, 5, 6), variant=variant)
    mask = np.tile(d["mask"], (3, 1, 1, 1))
    vertex = np.tile(d["vertex"], (3, 1, 1, 1, 1))
    mask[7] = 0.0  # an image without any object
    mi = 6 if variant == "hard" else 20
    single, dbg, _, rdbg = _compare(vote, mask, vertex, 64, seed=5, max_iter=mi)
    if variant == "hard":
        assert max(r["rounds"] for row in rdbg for r in row) > 1
    m, v = torch.from_numpy(mask).cuda(), torch.from_numpy(vertex).cuda()
    for _ in range(3):  # first call builds the graph, later calls patch its kernel nodes
        again = vote(m, v, 64, seed=5, max_iter=mi)
        assert torch.equal(again, single)
    shifted = vote(m, v, 64, seed=5, max_iter=mi, image_offset=3)
    part = vote(m[5:].contiguous(), v[5:].contiguous(), 64, seed=5, max_iter=mi, image_offset=8)
    assert torch.equal(shifted[5:], part)


@pytest.mark.parametrize("vn,hn,ids", [(5, 100, (1,)), (12, 33, (1, 5)), (1, 300, (5, 6, 8))])
def test_other_shapes_keypoint_counts_and_ragged_hypothesis_numbers(vote, vn, hn, ids):
    """vn != 9 (generic direction gather), hypothesis counts that are not a multiple of the 256-wide groups,
    a single class: still bit-exact against the oracle."""
    d = synthetic.make_frames(2, 96, 128, ids, variant="hard")
    v9 = d["vertex"]
    vertex = np.ascontiguousarray(np.concatenate([v9, v9[:, :, :, ::-1]], axis=3)[:, :, :, :vn])  # up to 18 keypoint fields
    _compare(vote, d["mask"], vertex, hn, seed=3, max_iter=4)


def test_degenerate_inputs(vote):
    h, w = 64, 80
    mask = np.zeros((1, h, w, 4), np.float32)
    vertex = np.zeros((1, h, w, 9, 2), np.float32)
    ys, xs = np.mgrid[0:h, 0:w]
    for k in range(9):
        dx, dy = (40.3 + k) - (xs + 0.5), (30.7 - k) - (ys + 0.5)
        n = np.sqrt(dx * dx + dy * dy)
        vertex[0, :, :, k, 0], vertex[0, :, :, k, 1] = dy / n, dx / n
    mask[0, 5, 5:9, 0] = 1  # class 0: 4 pixels < min_num -> zeros
    mask[0, 20:40, 20:50, 1] = 1  # class 1: regular
    mask[0, 50, 10:70, 2] = 1  # class 2: one pixel row
    # class 3: empty
    vertex[0, 50, :, 3] = (0.0, 1.0)  # keypoint 3 of the row: parallel vectors -> singular ATA -> all unrefined
    vertex[0, 20:40, 20:50, 5] = 0.0  # keypoint 5 of class 1: zero vectors -> no votes, max_iter rounds
    pts, dbg, ref, rdbg = _compare(vote, mask, vertex, 32, max_iter=3)
    assert np.array_equal(pts[0, 0].cpu().numpy(), np.zeros((9, 2), np.float32))
    assert np.array_equal(pts[0, 3].cpu().numpy(), np.zeros((9, 2), np.float32))
    assert int(dbg["rounds"][0, 1]) == 3 and int(dbg["refined"][0, 2]) == 0


def test_keypoint_exactly_on_a_pixel_centre_uses_the_exact_list(vote):
    """A noise-free field aimed at a pixel centre makes hypotheses coincide with that pixel (|hd| = 0, :240 guard)."""
    h, w = 48, 64
    mask = np.zeros((1, h, w, 1), np.float32)
    ys, xs = np.mgrid[0:h, 0:w]
    mask[0, ((xs - 30) ** 2 + (ys - 20) ** 2) <= 100, 0] = 1
    vertex = np.zeros((1, h, w, 9, 2), np.float32)
    for k in range(9):
        dx, dy = (28.5 + k) - (xs + 0.5), (18.5 + k % 3) - (ys + 0.5)  # integer + .5 = pixel centres inside the mask
        n = np.sqrt(dx * dx + dy * dy) + 1e-30
        vertex[0, :, :, k, 0], vertex[0, :, :, k, 1] = dy / n, dx / n
    _, dbg, _, _ = _compare(vote, mask, vertex, 64)
    assert int(dbg["stats"][2]) > 0  # some hypotheses went through the exact list


def test_non_finite_and_huge_directions_fall_back_to_exact(vote):
    d = synthetic.make_frames(1, 96, 128, (1, 5), variant="easy")
    v = d["vertex"].copy()
    ys, xs = np.nonzero(d["mask"][0, :, :, 0])
    v[0, ys[0], xs[0], 2] = (np.inf, 1.0)
    v[0, ys[1], xs[1], 3] = (np.nan, np.nan)
    v[0, ys[2], xs[2], 4] = (3e30, -2e30)
    with np.errstate(all="ignore"):
        _compare(vote, d["mask"], v, 48, max_iter=2)


def test_inlier_threshold_outside_filter_range_is_still_exact(vote):
    d = synthetic.make_frames(1, 96, 128, (1, 5), variant="easy")
    _compare(vote, d["mask"], d["vertex"], 48, inlier_thresh=0.3, max_iter=2)
    _compare(vote, d["mask"], d["vertex"], 48, inlier_thresh=0.999, max_iter=2)


def test_filter_selftest_no_mismatch(cuda_lib):
    from casapose_b200 import _lib

    h = _lib.handle(0)
    # negative spread = extreme magnitudes (coordinates to 65535, |d| in 2^[-19,29], distances 2^[-17,59])
    for thr, spread in ((0.99, 2e-5), (0.99, 1e-2), (0.9, 2e-5), (0.999, 5e-5), (0.6, 1e-4), (0.99, -2e-5), (0.99, -1e-3),
                        (0.75, -5e-5)):
        res = (C.c_uint64 * 4)()
        _lib.check(cuda_lib.casa_selftest_filter(h, 1 << 26, 99, thr, spread, res))
        tested, bad, unc, inl = list(res)
        assert tested > (1 << 25) and bad == 0, (thr, spread, list(res))
        assert 0 < inl < tested


def test_errors_are_reported_not_swallowed(vote):
    from casapose_b200._lib import CasaError

    m = torch.zeros(1, 8, 8, 2, device="cuda")
    v = torch.zeros(1, 8, 8, 9, 2, device="cuda")
    with pytest.raises(CasaError):
        vote(m, v, 0)  # round_hyp_num out of range
    with pytest.raises(TypeError):
        vote(m.double(), v, 16)
    with pytest.raises(ValueError):
        vote(m.cpu(), v, 16)
    multi = torch.ones(1, 8, 8, 2, device="cuda")  # every pixel in both classes: 2*h*w list entries per image
    with pytest.raises(CasaError):
        vote(multi, v, 16, pix_capacity=64)  # an explicit capacity that is too small is an error
    out = vote(multi, v, 16, pix_capacity=2 * 64)  # enough room: runs
    assert out.shape == (1, 2, 9, 2)
    # default capacity (h*w): the reference treats every channel independently and accepts overlapping channels
    # (ransac_voting.py:458-470) — the synchronous call is repeated with room for every channel
    rng = np.random.default_rng(3)
    vv = torch.from_numpy(rng.normal(size=(1, 8, 8, 9, 2)).astype(np.float32)).cuda()
    auto = vote(multi, vv, 16, seed=5)
    assert torch.equal(auto, vote(multi, vv, 16, seed=5, pix_capacity=2 * 64))


def test_host_buffer_entry_point_matches_device_path(vote):
    from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks_host

    d = synthetic.make_frames(2, 96, 128, (1, 5, 6), variant="easy")
    dev = vote(torch.from_numpy(d["mask"]).cuda(), torch.from_numpy(d["vertex"]).cuda(), 64, seed=4)
    host = ransac_voting_layer_all_masks_host(torch.from_numpy(d["mask"]).pin_memory(), torch.from_numpy(d["vertex"]).pin_memory(), 64, seed=4)
    assert torch.equal(dev.cpu(), host)


def test_host_entry_packs_a_pageable_mask_beside_a_pinned_vertex_field(vote):
    """The host entry packs the mask into membership words on host threads (any host memory) and reads the vector
    field in place (pinned): 8 classes take the 64-bit fast path of the packer, batch 9 splits into 4 image ranges."""
    from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks_host

    d = synthetic.make_frames(3, 96, 128, synthetic.CONFIG_8_IDS, variant="easy")
    mask = np.tile(d["mask"], (3, 1, 1, 1))
    vertex = np.tile(d["vertex"], (3, 1, 1, 1, 1))
    dev = vote(torch.from_numpy(mask).cuda(), torch.from_numpy(vertex).cuda(), 64, seed=9)
    host = ransac_voting_layer_all_masks_host(torch.from_numpy(mask), torch.from_numpy(vertex).pin_memory(), 64, seed=9)
    assert torch.equal(dev.cpu(), host)
    again = ransac_voting_layer_all_masks_host(torch.from_numpy(mask), torch.from_numpy(vertex).pin_memory(), 64, seed=9)
    assert torch.equal(host, again)


@pytest.mark.parametrize("variant", ["easy", "hard"])
def test_two_lane_mode_equals_the_synchronous_calls(vote, variant):
    """casa_set_async(h, 2): consecutive votes alternate between two lanes (own workspace, own stream) and overlap;
    after casa_join the outputs are those of the synchronous calls, bit for bit (multi-round frames included), and
    casa_sync reports no error."""
    from casapose_b200 import _lib

    d = synthetic.make_frames(2, 96, 128, synthetic.CONFIG_8_IDS, variant=variant)
    mask, vertex = torch.from_numpy(d["mask"]).cuda(), torch.from_numpy(d["vertex"]).cuda()
    seeds = [3, 4, 5, 6, 7]
    expect = [vote(mask, vertex, 64, seed=s).clone() for s in seeds]
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    _lib.set_async(0, 2, stream)
    try:
        outs = [torch.full_like(expect[0], -1.0) for _ in seeds]
        for s, o in zip(seeds, outs):
            vote(mask, vertex, 64, seed=s, out=o)
        _lib.join(0, stream)
        got = [o.clone() for o in outs]  # on the caller's stream, behind the join
        _lib.sync(0, stream)
    finally:
        _lib.set_async(0, 0, stream)
    for e, g in zip(expect, got):
        assert torch.equal(e, g)
    again = vote(mask, vertex, 64, seed=seeds[0])
    assert torch.equal(again, expect[0])


def test_overlapping_channels_vote_independently_like_the_reference(vote):
    """Channels of `mask` need not be mutually exclusive in the reference (every channel is voted on by itself,
    ransac_voting.py:458-470): a pixel in two channels is listed twice; counts, winners and keypoints are the oracle's."""
    d = synthetic.make_frames(1, 64, 96, (1, 5), variant="easy")
    mask = d["mask"].copy()
    mask[0, :, :, 1] = np.maximum(mask[0, :, :, 1], mask[0, :, :, 0])  # channel 1 = union of both objects
    _compare(vote, mask, d["vertex"], 64, seed=11, pix_capacity=2 * 64 * 96)


def test_lanes_with_changing_shapes_and_a_failing_call(vote):
    """Lane mode with calls of different shapes (each lane keeps its own graphs and workspace), and a call whose pixel
    lists overflow: the error surfaces at casa_sync (asynchronous calls cannot repeat themselves), the other calls'
    results are the eager ones, and the handle works on afterwards."""
    from casapose_b200 import _lib
    from casapose_b200._lib import CasaError

    frames = [synthetic.make_frames(b, hh, ww, ids, variant="easy", seed=3 + b)
              for b, hh, ww, ids in ((1, 64, 96, (1, 5)), (2, 96, 128, synthetic.CONFIG_8_IDS), (3, 48, 64, (6,)), (1, 96, 128, (1, 5, 6)))]
    ins = [(torch.from_numpy(d["mask"]).cuda(), torch.from_numpy(d["vertex"]).cuda()) for d in frames]
    expect = [vote(m, v, 64, seed=2).clone() for m, v in ins]
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    _lib.set_async(0, 3, stream)
    try:
        outs = [vote(m, v, 64, seed=2) for m, v in ins + ins]
        _lib.join(0, stream)
        got = [o.clone() for o in outs]
        _lib.sync(0, stream)
        for k, g in enumerate(got):
            assert torch.equal(expect[k % len(ins)], g), k
        multi = torch.ones(1, 8, 8, 2, device="cuda")  # every pixel in both channels: 2*h*w entries > h*w slots
        vote(multi, torch.zeros(1, 8, 8, 9, 2, device="cuda"), 16)
        ok = vote(ins[0][0], ins[0][1], 64, seed=2)
        with pytest.raises(CasaError):
            _lib.sync(0, stream)
        assert torch.equal(ok, expect[0])
        _lib.sync(0, stream)  # the error was reported once
    finally:
        _lib.set_async(0, 0, stream)
    assert torch.equal(vote(ins[1][0], ins[1][1], 64, seed=2), expect[1])
