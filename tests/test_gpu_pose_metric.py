"""ADD / ADD-S / 2-D errors on the GPU (casa_pose_errors) against the oracle's restatement of
map_estimates / evaluate_poses (ransac_voting.py:561-687) and against the product's own numpy backend."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from casapose_b200 import synthetic  # noqa: E402
from oracle import pose_np as OP  # noqa: E402

pytestmark = pytest.mark.gpu
F = np.float32


def _scene(b=3, seed=0, counts=(500, 7862, 9, 3417, 1200)):
    """Random model clouds (two with the symmetric meshes' vertex counts), ground-truth poses, perturbed estimates."""
    rng = np.random.default_rng(seed)
    oc, maxp = len(counts), max(counts)
    pts = np.zeros((oc, maxp, 3), F)
    diam = np.zeros(oc, F)
    for c, n in enumerate(counts):
        ext = rng.uniform(30, 120, size=3)
        pts[c, :n] = (rng.uniform(-0.5, 0.5, size=(n, 3)) * ext).astype(F)
        diam[c] = np.linalg.norm(ext)
    K = synthetic.camera_matrix(480).astype(F)
    cams = np.broadcast_to(K, (b, 3, 3)).copy()
    gt = np.zeros((b, oc, 1, 3, 4), F)
    est = np.zeros((b, oc, 3, 4), F)
    valid = np.ones((b, oc), np.int32)
    for i in range(b):
        for c in range(oc):
            R = synthetic._random_rotation(rng)
            t = np.array([rng.uniform(-150, 150), rng.uniform(-100, 100), rng.uniform(600, 1200)])
            gt[i, c, 0] = np.concatenate([R, t[:, None]], 1)
            mag = [0.002, 0.02, 0.08][(i + c) % 3]  # well inside, near and beyond the 0.1 d threshold
            w = rng.normal(size=3) * mag
            th = np.linalg.norm(w)
            k = w / th
            Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            dR = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
            est[i, c] = np.concatenate([dR @ R, (t + rng.normal(size=3) * mag * 100)[:, None]], 1)
    est[0, 0] = 0.0  # missing object (:577)
    valid[1, 2] = 0  # no ground truth, pose present -> false positive (:573)
    valid[-1, -1] = 0
    est[-1, -1] = 0.0  # no ground truth, no pose
    return pts, np.array(counts, np.int32), diam, cams, gt, est, valid


def test_rows_match_the_oracle(cuda_lib):
    from casapose_b200.pose_estimation.ransac_voting import evaluate_poses, pose_errors_cuda

    pts, counts, diam, cams, gt, est, valid = _scene()
    b, oc = valid.shape
    rows = pose_errors_cuda(est.reshape(-1, 3, 4), gt.reshape(-1, 3, 4), np.broadcast_to(cams[:, None], (b, oc, 3, 3)).reshape(-1, 3, 3),
                            pts, counts, np.broadcast_to(diam[None], (b, oc)).reshape(-1), valid.reshape(-1), 5.0)
    rows = rows.cpu().numpy().reshape(b, oc, 6)
    pts5 = np.broadcast_to(pts[None, :, None], (b, oc, 1) + pts.shape[1:])
    cnt3 = np.broadcast_to(counts[None, :, None], (b, oc, 1))
    ref = OP.evaluate_poses(est, gt, pts5, cnt3, cams, np.broadcast_to(diam[None], (b, oc)), valid)
    assert np.array_equal(rows[:, :, 2].sum(0), ref["valid_3d"]) and np.array_equal(rows[:, :, 3].sum(0), ref["valid_2d"])
    assert np.array_equal(rows[:, :, 4].sum(0), ref["missing"]) and np.array_equal(rows[:, :, 5].sum(0), ref["false_positive"])
    assert np.array_equal(rows[:, :, 3:1:-1], ref["flags"])  # (valid_2d, valid_3d) per object
    assert np.allclose(rows[:, :, 0].sum(0), ref["err_2d"], rtol=1e-5, atol=1e-4)
    assert np.allclose(rows[:, :, 1].sum(0), ref["err_3d"], rtol=1e-5, atol=1e-4)
    assert rows[0, 0].tolist() == [F(99.9), F(999.9), 0, 0, 1, 0] and rows[1, 2].tolist() == [0, 0, 0, 0, 0, 1]
    assert not rows[-1, -1].any()
    assert 0 < ref["valid_3d"].sum() < valid.sum()  # the scene straddles the threshold
    # the drop-in signature with both backends
    a = evaluate_poses(est, gt, None, pts5, cnt3, cams, diam, valid, 5.0, backend="cuda")
    c = evaluate_poses(est, gt, None, pts5, cnt3, cams, diam, valid, 5.0)
    for x, y in zip(a, c):
        assert np.allclose(x, y, rtol=1e-5, atol=1e-4)
    for k in (2, 3, 4, 5, 6):
        assert np.array_equal(a[k], c[k])


def test_per_object_model_table_and_keypoint_models(cuda_lib):
    from casapose_b200.pose_estimation.ransac_voting import evaluate_poses

    pts, counts, diam, cams, gt, est, valid = _scene(b=2, seed=5, counts=(9, 9, 9))
    b, oc = valid.shape
    pts5 = np.broadcast_to(pts[None, :, None], (b, oc, 1) + pts.shape[1:]).copy()
    pts5[1] *= F(1.1)  # batch entries differ -> the general (b*oc models) route
    cnt3 = np.full((b, oc, 1), 9, np.int32)
    a = evaluate_poses(est, gt, None, pts5, cnt3, cams, diam, valid, 5.0, backend="cuda")
    c = evaluate_poses(est, gt, None, pts5, cnt3, cams, diam, valid, 5.0)
    for x, y in zip(a, c):
        assert np.allclose(x, y, rtol=1e-5, atol=1e-4)


def test_pipeline_with_gpu_metrics_gives_identical_statistics(cuda_lib):
    from casapose_b200.pose_estimation import estimate_and_evaluate_poses
    from tests.test_gpu_pipeline import _inputs

    d, cams, offsets, kp3, target_seg, poses_gt, diam = _inputs()
    b = diam.shape[0]
    args = (torch.from_numpy(d["seg_logits"]).cuda(), torch.from_numpy(target_seg).cuda(),
            torch.from_numpy(d["vertex"].reshape(b, 240, 320, 18)).cuda(), poses_gt, kp3, cams, diam, offsets)
    s0, p0, _ = estimate_and_evaluate_poses(*args, seed=11)
    s1, p1, _ = estimate_and_evaluate_poses(*args, seed=11, metric_backend="cuda")
    assert torch.equal(p0, p1)
    for k in (0, 1, 2, 3, 6, 7):
        assert np.array_equal(np.asarray(s0[k]), np.asarray(s1[k])), k
    assert np.allclose(s0[4], s1[4], rtol=1e-5, atol=1e-4) and np.allclose(s0[5], s1[5], rtol=1e-5, atol=1e-4)


def test_device_evaluator_equals_the_drop_in_with_gpu_backends(cuda_lib):
    from casapose_b200.pose_estimation import DeviceEvaluator, estimate_and_evaluate_poses
    from tests.test_gpu_pipeline import _inputs

    d, cams, offsets, kp3, target_seg, poses_gt, diam = _inputs()
    b, oc = diam.shape
    seg = torch.from_numpy(d["seg_logits"]).cuda()
    tgt = torch.from_numpy(target_seg).cuda()
    vert = torch.from_numpy(d["vertex"].reshape(b, 240, 320, 18)).cuda()
    s0, p0, k0 = estimate_and_evaluate_poses(seg, tgt, vert, poses_gt, kp3, cams, diam, offsets, seed=11,
                                             pnp_backend="cuda", metric_backend="cuda")
    ev = DeviceEvaluator(kp3[0, :, 0], cams[0], diam[0])
    s1, p1, k1 = ev(seg, tgt, vert, poses_gt[:, :, 0], offsets, seed=11)
    assert torch.equal(k0, k1) and torch.equal(p0, p1.cpu())
    names = ["valid_2d", "valid_3d", "valid_pose_count", "false_positive_mask", "err_2d", "err_3d", "missing_object", "false_positive_pose"]
    for i, n in enumerate(names):
        a, c = np.atleast_1d(np.asarray(s0[i], np.float32)), s1[n]
        assert np.allclose(a, c, rtol=1e-6, atol=1e-5), n
