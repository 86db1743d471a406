"""GPU parity against the golden vectors in tests/golden/ — outputs of the REFERENCE'S OWN SOURCE FILES executed
over a numpy TensorFlow stand-in (oracle/make_golden.py; tests/test_golden_oracle.py pins the CPU oracle to the
same files).  Everything goes through the C ABI (casapose_b200.pose_estimation -> libcasapose_b200.so).

Bar (BASELINE.json north_star): vote counts / rounds bit-exact, refined keypoints within 1e-3 px, identical
ADD / ADD-S / 2-D verdicts after PnP."""
import ast

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import golden_inputs as GI  # noqa: E402
from oracle import ls_voting_np as OL  # noqa: E402

from .test_golden_oracle import LS_GRAD_CASES, RANSAC_CASES, _well_conditioned, golden_counts, load, ransac_case_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
TOL_PX = 1e-3
F = np.float32


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("name", RANSAC_CASES)
def test_ransac_votes_equal_reference_code(cuda_lib, name):
    from casapose_b200.pose_estimation import ransac_voting_layer_all_masks

    g = load(name)
    mask, vertex, hn, seed, kw = ransac_case_inputs(g)
    pts, dbg = ransac_voting_layer_all_masks(cu(mask), cu(vertex), hn, seed=seed, return_debug=True, **kw)
    torch.cuda.synchronize()
    b, oc = mask.shape[0], mask.shape[3]
    assert np.array_equal(dbg["rounds"].cpu().numpy().astype(np.int32), g["rounds"]), "rounds differ"
    counts = dbg["counts"].cpu().numpy()
    for i in range(b):
        for c in range(oc):
            for k, ref in enumerate(golden_counts(g, i, c)):
                assert np.array_equal(counts[i, c, k], ref), ("vote counts", i, c, k)
                win = dbg["win_idx"][i, c, k].cpu().numpy()
                assert np.array_equal(win, ref.argmax(axis=0)), ("winning hypothesis", i, c, k)
    err = float(np.abs(pts.cpu().numpy() - g["points"]).max())
    assert err <= TOL_PX, "refined keypoints differ from the reference code's by %g px" % err
    # the production call (no debug outputs) replays the CUDA graph: same keypoints
    again = ransac_voting_layer_all_masks(cu(mask), cu(vertex), hn, seed=seed, **kw)
    assert torch.equal(again, pts)


def _ls(layer_kw, seg, direct, conf):
    from casapose_b200.pose_estimation import CoordLSVotingWeighted

    layer = CoordLSVotingWeighted("ls", seg.shape[-1], num_points=9, **layer_kw)
    out = layer([cu(seg), cu(direct), cu(conf)])
    _, dbg = OL.coord_ls_voting_weighted(seg, direct, conf, return_debug=True, **layer_kw)
    return out.cpu().numpy(), _well_conditioned(dbg)


@pytest.mark.parametrize("case,key,layer_kw", [
    ("ls_plain", "points", {}),
    ("ls_plain", "points_sigmoid", {"sigmoid_weights": True}),
    ("ls_filter", "points", {"filter_estimates": True}),
    ("ls_filter", "points_second", {"filter_estimates": True, "output_second_largest_component": True}),
])
def test_ls_layer_equals_reference_code(cuda_lib, case, key, layer_kw):
    g = load(case)
    out, ok = _ls(layer_kw, g["seg"], g["direct"], g["conf"])
    assert np.abs(out - g[key])[ok].max() <= TOL_PX


@pytest.mark.parametrize("key,layer_kw", [("points", {}), ("points_filter", {"filter_estimates": True})])
def test_ls_layer_full_size_equals_reference_code(cuda_lib, key, layer_kw):
    g = load("ls_full_480x640")
    seg, direct, conf = GI.ls_inputs(**ast.literal_eval(str(g["gen"])))
    assert GI.sha(seg, direct, conf) == str(g["input_sha"])
    out, ok = _ls(layer_kw, seg, direct, conf)
    assert ok.all() and np.abs(out - g[key]).max() <= TOL_PX


@pytest.mark.parametrize("tag,kw", LS_GRAD_CASES)
def test_ls_backward_equals_autodiff_of_reference_code(cuda_lib, tag, kw):
    """casa_ls_vote_backward vs torch.autograd over the reference's own layer code (tests/golden/ls_grad.npz)."""
    from casapose_b200.pose_estimation import CoordLSVotingWeighted

    g = load("ls_grad")
    seg, direct, conf, go = GI.ls_grad_inputs()
    assert GI.sha(seg, direct, conf, go) == str(g["input_sha"])
    layer = CoordLSVotingWeighted("ls", seg.shape[3], **kw)
    gd, gw = layer.backward([cu(seg), cu(direct), cu(conf)], cu(go))
    for mine, ref, what in ((gd, g["grad_direct_" + tag], "direct"), (gw, g["grad_conf_" + tag], "conf")):
        mine = mine.cpu().numpy()
        assert np.abs(mine - ref).max() <= 2e-4 * np.abs(ref).max(), what
        differ = (mine != 0) != (ref != 0)
        assert not differ.any() or max(np.abs(mine[differ]).max(), np.abs(ref[differ]).max()) <= 1e-6 * np.abs(ref).max(), what


@pytest.mark.parametrize("pnp_backend,metric_backend", [("cv2", "numpy"), ("cuda", "cuda")])
def test_pose_pipeline_equals_reference_code(cuda_lib, pnp_backend, metric_backend):
    """estimate_and_evaluate_poses: identical ADD / 2-D verdicts and bookkeeping, with the reference's host stages
    (OpenCV PnP, numpy metrics) and with the device-resident ones (casa_pnp, casa_pose_errors)."""
    from casapose_b200.pose_estimation import estimate_and_evaluate_poses

    g = load("pose_eval")
    gen = ast.literal_eval(str(g["gen"]))
    d, cams, offsets, kp3, target_seg, poses_gt, diam = GI.pose_inputs(**gen)
    b, h, w = gen["b"], gen["h"], gen["w"]
    vertex18 = d["vertex"].reshape(b, h, w, 18)
    assert GI.sha(d["seg_logits"], vertex18, target_seg, offsets) == str(g["input_sha"])
    stats, poses, pts = estimate_and_evaluate_poses(cu(d["seg_logits"]), cu(target_seg), cu(vertex18), poses_gt, kp3, cams,
                                                    diam, offsets, seed=int(g["seed"]), pnp_backend=pnp_backend,
                                                    metric_backend=metric_backend)
    valid_2d, valid_3d, valid_count, fp_mask, err_2d, err_3d, missing, fp_pose = [np.asarray(s, F) for s in stats]
    assert np.abs(pts.cpu().numpy() - g["points"]).max() <= TOL_PX
    assert np.array_equal(valid_3d, g["valid_3d"]), "ADD verdicts differ from the reference code's"
    assert np.array_equal(valid_2d, g["valid_2d"]) and np.array_equal(missing, g["missing_object"])
    assert np.array_equal(valid_count, g["valid_pose_count"]) and np.array_equal(fp_pose, g["false_positive_pose"])
    assert np.array_equal(np.atleast_1d(fp_mask), np.atleast_1d(g["false_positive_mask"]))
    assert np.allclose(err_3d, g["err_3d"], rtol=2e-2, atol=0.5) and np.allclose(err_2d, g["err_2d"], rtol=2e-2, atol=0.05)
    p = poses.cpu().numpy() if isinstance(poses, torch.Tensor) else np.asarray(poses)
    assert np.abs(p[..., :3] - g["poses"][..., :3]).max() < 2e-2 and np.abs(p[..., 3] - g["poses"][..., 3]).max() < 5.0


def test_pose_metric_kernel_equals_reference_code(cuda_lib):
    """casa_pose_errors (ADD, ADD-S on 7862 / 3417 points, 2-D, bookkeeping) vs evaluate_poses of the reference."""
    from casapose_b200.pose_estimation.ransac_voting import evaluate_poses

    g = load("pose_metric")
    s = GI.metric_scene()
    assert GI.sha(*[s[k] for k in sorted(s)]) == str(g["input_sha"])
    b, oc = s["valid"].shape
    ev = np.broadcast_to(s["evaluation_points"][None, :, None], (b, oc, 1) + s["evaluation_points"].shape[1:])
    cnt = np.broadcast_to(s["counts"][None], (b, oc, 1))
    for backend in ("cuda", "numpy"):
        err_2d, err_3d, valid_2d, valid_3d, missing, count, fp = evaluate_poses(
            s["poses"], s["poses_gt"], s["points_estimated"], ev, cnt, s["cams"], s["diameters"], s["valid"], 5.0, backend=backend)
        assert np.array_equal(valid_3d, g["valid_3d"]) and np.array_equal(valid_2d, g["valid_2d"]), backend
        assert np.array_equal(missing, g["missing_object"]) and np.array_equal(fp, g["false_positive_pose"]), backend
        assert np.array_equal(count, g["valid_points_count"])
        assert np.allclose(err_3d, g["err_3d"], rtol=1e-4, atol=1e-3) and np.allclose(err_2d, g["err_2d"], rtol=1e-4, atol=1e-3)


def test_ls_layer_plus_poses_pnp_equals_reference_code(cuda_lib):
    from casapose_b200.pose_estimation import CoordLSVotingWeighted, poses_pnp

    g = load("poses_pnp")
    gen = ast.literal_eval(str(g["gen"]))
    d, cams, offsets, kp3, target_seg, poses_gt, diam = GI.pose_inputs(**gen)
    b, h, w, oc = gen["b"], gen["h"], gen["w"], len(gen["ids"])
    vertex18 = d["vertex"].reshape(b, h, w, 18)
    assert GI.sha(d["seg_logits"], vertex18, d["conf_logits"]) == str(g["input_sha"])
    layer = CoordLSVotingWeighted("ls", oc + 1, num_points=9, filter_estimates=True)
    coords = layer([cu(d["seg_logits"]), cu(vertex18), cu(d["conf_logits"])])
    assert np.abs(coords.cpu().numpy() - g["coords"]).max() <= TOL_PX
    for backend in ("cv2", "cuda"):
        poses = poses_pnp(coords, cu(d["seg_logits"]), kp3, cams, oc, min_num=20, pnp_backend=backend).numpy()
        assert poses.shape == g["poses"].shape
        assert np.array_equal(poses.any(axis=(2, 3, 4)), g["poses"].any(axis=(2, 3, 4))), "availability mask differs"
        assert np.abs(poses[..., :3] - g["poses"][..., :3]).max() < 2e-2, backend
        assert np.abs(poses[..., 3] - g["poses"][..., 3]).max() < 5.0, backend


def test_pvnet_style_fields_equal_reference_code(cuda_lib):
    """vertex [b,h,w,oc*vn*2] through estimate_and_evaluate_poses and pose_estimation (pose_evaluation.py:38-45, 222-269)."""
    from casapose_b200.pose_estimation import estimate_and_evaluate_poses, pose_estimation

    g = load("pose_eval_pvnet")
    d, fields, cams, offsets, kp3, target_seg, poses_gt, diam = GI.pvnet_inputs()
    assert GI.sha(d["seg_logits"], fields, target_seg, offsets) == str(g["input_sha"])
    stats, poses, pts = estimate_and_evaluate_poses(cu(d["seg_logits"]), cu(target_seg), cu(fields), poses_gt, kp3, cams, diam,
                                                    offsets, seed=int(g["seed"]))
    assert np.abs(pts.cpu().numpy() - g["points"]).max() <= TOL_PX
    valid_2d, valid_3d, valid_count, fp_mask, err_2d, err_3d, missing, fp_pose = [np.asarray(s, F) for s in stats]
    assert np.array_equal(valid_3d, g["valid_3d"]) and np.array_equal(valid_2d, g["valid_2d"])
    assert np.array_equal(missing, g["missing_object"]) and np.array_equal(valid_count, g["valid_pose_count"])
    assert np.allclose(err_3d, g["err_3d"], rtol=2e-2, atol=0.5) and np.allclose(err_2d, g["err_2d"], rtol=2e-2, atol=0.05)
    p2 = pose_estimation(cu(d["seg_logits"]), cu(target_seg), cu(fields), poses_gt, kp3, cams, offsets, seed=int(g["seed"]))
    p2 = p2.cpu().numpy() if isinstance(p2, torch.Tensor) else np.asarray(p2)
    assert np.abs(p2[..., :3] - g["poses_pose_estimation"][..., :3]).max() < 2e-2
    assert np.abs(p2[..., 3] - g["poses_pose_estimation"][..., 3]).max() < 5.0


@pytest.mark.parametrize("name,pinned", [("ransac_full_480x640", True), ("ransac_hard", True), ("ransac_cap", False)])
def test_host_buffer_entry_equals_reference_code(cuda_lib, name, pinned):
    """casa_ransac_vote_host (the benchmark's e2e path: zero-copy reads of pinned buffers, staged copy of pageable
    ones) against the reference code's keypoints."""
    from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks_host

    g = load(name)
    mask, vertex, hn, seed, kw = ransac_case_inputs(g)
    tm, tv = torch.from_numpy(np.ascontiguousarray(mask)), torch.from_numpy(np.ascontiguousarray(vertex))
    if pinned:
        tm, tv = tm.pin_memory(), tv.pin_memory()
    out = ransac_voting_layer_all_masks_host(tm, tv, hn, seed=seed, **kw)
    assert np.abs(out.numpy() - g["points"]).max() <= TOL_PX
