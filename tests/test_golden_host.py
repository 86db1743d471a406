"""Host-side stages of the drop-in (the code around the CUDA voting: un-mapping, OpenCV PnP, ADD / ADD-S
bookkeeping, poses_pnp) against the golden vectors produced by the reference's own source
(oracle/make_golden.py).  No GPU: the voting results are taken from the golden files."""
import ast

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import golden_inputs as GI  # noqa: E402

from .test_golden_oracle import load  # noqa: E402

F = np.float32


def test_transform_points_back_equals_reference_code():
    from casapose_b200.pose_estimation.ransac_voting import transform_points_back

    g = load("unmap_offsets")
    for pts, o, ref in zip(g["points"], g["offsets"], g["unmapped"]):
        if abs(pts.sum(dtype=F)) < 0.01:
            continue
        out = transform_points_back(pts, o[0], o[1], o[8], o[9], o[4], o[5], o[6], o[7])
        assert np.abs(out - ref).max() <= 2e-3


def test_estimate_and_evaluate_host_stages_equal_reference_code():
    """points_estimated given (normalised, pose_evaluation.py:60) -> un-mapping -> OpenCV PnP -> metrics."""
    from casapose_b200.pose_estimation.ransac_voting import estimate_poses, evaluate_poses

    g = load("pose_eval")
    gen = ast.literal_eval(str(g["gen"]))
    d, cams, offsets, kp3, target_seg, poses_gt, diam = GI.pose_inputs(**gen)
    b, oc = gen["b"], len(gen["ids"])
    valid = ((target_seg[..., 1:] != 0).sum((1, 2)) > 20).astype(np.int32)
    poses, fp = estimate_poses(g["points"], kp3, cams, valid, offsets)
    assert np.abs(poses[..., :3] - g["poses"][..., :3]).max() < 2e-2 and np.abs(poses[..., 3] - g["poses"][..., 3]).max() < 5.0
    err_2d, err_3d, valid_2d, valid_3d, missing, count, fp_pose = evaluate_poses(
        poses, poses_gt, g["points"], kp3, np.full((b, oc, 1), 9, np.int32), cams, diam, valid, 5.0)
    assert np.array_equal(valid_3d, g["valid_3d"]) and np.array_equal(valid_2d, g["valid_2d"])
    assert np.array_equal(missing, g["missing_object"]) and np.array_equal(count, g["valid_pose_count"])
    assert np.array_equal(fp_pose, g["false_positive_pose"]) and np.array_equal(np.atleast_1d(fp), np.atleast_1d(g["false_positive_mask"]))
    assert np.allclose(err_3d, g["err_3d"], rtol=2e-2, atol=0.5) and np.allclose(err_2d, g["err_2d"], rtol=2e-2, atol=0.05)


def test_evaluate_poses_numpy_backend_equals_reference_code():
    from casapose_b200.pose_estimation.ransac_voting import evaluate_poses

    g = load("pose_metric")
    s = GI.metric_scene()
    b, oc = s["valid"].shape
    ev = np.broadcast_to(s["evaluation_points"][None, :, None], (b, oc, 1) + s["evaluation_points"].shape[1:])
    cnt = np.broadcast_to(s["counts"][None], (b, oc, 1))
    err_2d, err_3d, valid_2d, valid_3d, missing, count, fp = evaluate_poses(
        s["poses"], s["poses_gt"], s["points_estimated"], ev, cnt, s["cams"], s["diameters"], s["valid"], 5.0)
    assert np.array_equal(valid_3d, g["valid_3d"]) and np.array_equal(valid_2d, g["valid_2d"])
    assert np.array_equal(missing, g["missing_object"]) and np.array_equal(fp, g["false_positive_pose"])
    assert np.array_equal(count, g["valid_points_count"])
    assert np.allclose(err_3d, g["err_3d"], rtol=1e-4, atol=1e-3) and np.allclose(err_2d, g["err_2d"], rtol=1e-4, atol=1e-3)


def test_poses_pnp_host_equals_reference_code():
    from casapose_b200.pose_estimation import poses_pnp

    g = load("poses_pnp")
    gen = ast.literal_eval(str(g["gen"]))
    d, cams, offsets, kp3, target_seg, poses_gt, diam = GI.pose_inputs(**gen)
    oc = len(gen["ids"])
    poses = poses_pnp(g["coords"], torch.from_numpy(d["seg_logits"]), kp3, cams, oc, min_num=20).numpy()
    assert poses.shape == g["poses"].shape
    assert np.array_equal(poses.any(axis=(2, 3, 4)), g["poses"].any(axis=(2, 3, 4)))
    assert np.abs(poses[..., :3] - g["poses"][..., :3]).max() < 2e-2 and np.abs(poses[..., 3] - g["poses"][..., 3]).max() < 5.0
