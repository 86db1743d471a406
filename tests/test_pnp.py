"""Batched PnP: the numpy restatement of the GPU algorithm against the reference's cv2 sequence (CPU), and the
GPU kernel against both (GPU)."""
import numpy as np
import pytest

from casapose_b200 import synthetic
from oracle import pnp_np
from oracle import pose_np as OP

F = np.float32


def _cases(n, seed=0, outliers=True):
    rng = np.random.default_rng(seed)
    models = synthetic.lm_models()
    K = synthetic.camera_matrix(480).astype(F)
    out = []
    for trial in range(n):
        oid = synthetic.CONFIG_13_IDS[trial % 13]
        X = np.array(models["obj_%06d" % oid]["keypoints"], F)
        R = synthetic._random_rotation(rng)
        t = np.array([rng.uniform(-200, 200), rng.uniform(-150, 150), rng.uniform(600, 1200)])
        cam = X @ R.T + t
        uv = cam @ K.T
        uv = (uv[:, :2] / uv[:, 2:]).astype(F) + rng.normal(scale=[0.1, 0.5, 2.0][trial % 3], size=(9, 2)).astype(F)
        if outliers and trial % 4 == 3:
            uv[rng.integers(0, 9)] += rng.normal(scale=40, size=2).astype(F)
        out.append((X, uv, K, np.concatenate([R, t[:, None]], 1)))
    return out


def _close(a, b):
    return np.abs(a[:, 3] - b[:, 3]).max() < 0.05 and np.abs(a[:, :3] - b[:, :3]).max() < 1e-4  # 0.05 mm, 1e-4


def test_restatement_matches_the_reference_cv2_sequence_on_clean_keypoints():
    for X, uv, K, _ in _cases(40, outliers=False):
        assert _close(pnp_np.pnp(X, uv, K), OP.pnp(X, uv, K))


def test_restatement_with_outlier_keypoints_mostly_matches_cv2():
    cases = [c for i, c in enumerate(_cases(120)) if i % 4 == 3]
    agree = sum(_close(pnp_np.pnp(X, uv, K), OP.pnp(X, uv, K)) for X, uv, K, _ in cases)
    assert agree >= 0.9 * len(cases)  # OpenCV's RANSAC is randomised: identical minima in the vast majority


def test_guards():
    X, uv, K, _ = _cases(1)[0]
    assert np.array_equal(pnp_np.pnp(X, np.zeros_like(uv), K), np.zeros((3, 4), F))
    assert np.array_equal(OP.pnp(X, np.zeros_like(uv), K), np.zeros((3, 4), F))


@pytest.mark.gpu
def test_gpu_pnp_matches_restatement_and_cv2(cuda_lib):
    import torch

    from casapose_b200.pose_estimation.ransac_voting import pnp_cuda

    cases = _cases(96)
    p2 = np.stack([c[1] for c in cases])
    p3 = np.stack([c[0] for c in cases])
    cam = np.stack([c[2] for c in cases])
    p2[5] = 0.0  # guard: zero keypoints -> zero pose
    got = pnp_cuda(torch.from_numpy(p2).cuda(), p3, cam).cpu().numpy()
    assert np.array_equal(got[5], np.zeros((3, 4), F))
    n_cv = 0
    for i, (X, uv, K, _) in enumerate(cases):
        if i == 5:
            continue
        ref = pnp_np.pnp(X, p2[i], K)
        assert np.abs(got[i] - ref).max() < 1e-3, i  # same algorithm, float64 on both sides (mm / rotation entries)
        n_cv += _close(got[i], OP.pnp(X, p2[i], K))
    assert n_cv >= 0.95 * (len(cases) - 1)


@pytest.mark.gpu
def test_gpu_pnp_offsets_unmapping(cuda_lib):
    import torch

    from casapose_b200.pose_estimation.ransac_voting import pnp_cuda, transform_points_back

    X, uv, K, _ = _cases(1, seed=3, outliers=False)[0]
    off = np.array([12.0, 30.0, 0, 0, 5.0, -3.0, 17.0, 0.8, 640.0, 480.0], F)
    # forward-map the keypoints so that un-mapping gives back uv: brute-force inverse through the product's own function
    back = transform_points_back(uv, off[0], off[1], off[8], off[9], off[4], off[5], off[6], off[7])
    a = pnp_cuda(torch.from_numpy(uv[None]).cuda(), X[None], K[None], off[None]).cpu().numpy()[0]
    b = pnp_np.pnp(X, back, K)
    assert np.abs(a - b).max() < 1e-2


@pytest.mark.gpu
def test_pipeline_with_gpu_pnp_gives_identical_add_verdicts(cuda_lib):
    import torch

    from casapose_b200.pose_estimation import estimate_and_evaluate_poses
    from tests.test_gpu_pipeline import _inputs

    d, cams, offsets, kp3, target_seg, poses_gt, diam = _inputs()
    b = diam.shape[0]
    args = (torch.from_numpy(d["seg_logits"]).cuda(), torch.from_numpy(target_seg).cuda(),
            torch.from_numpy(d["vertex"].reshape(b, 240, 320, 18)).cuda(), poses_gt, kp3, cams, diam, offsets)
    s_cv, p_cv, k_cv = estimate_and_evaluate_poses(*args, seed=11)
    s_gpu, p_gpu, k_gpu = estimate_and_evaluate_poses(*args, seed=11, pnp_backend="cuda")
    assert torch.equal(k_cv, k_gpu)
    assert np.array_equal(s_cv[1], s_gpu[1]) and np.array_equal(s_cv[0], s_gpu[0])  # valid_3d (ADD) and valid_2d
    assert np.abs(p_cv.numpy()[..., 3] - p_gpu.numpy()[..., 3]).max() < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("vn", [6, 12, 16])
def test_gpu_pnp_other_keypoint_counts(cuda_lib, vn):
    import torch

    from casapose_b200.pose_estimation.ransac_voting import pnp_cuda

    rng = np.random.default_rng(vn)
    K = synthetic.camera_matrix(480).astype(F)
    n = 12
    X = (rng.uniform(-0.5, 0.5, size=(n, vn, 3)) * 150).astype(F)
    p2 = np.zeros((n, vn, 2), F)
    for i in range(n):
        R = synthetic._random_rotation(rng)
        t = np.array([rng.uniform(-150, 150), rng.uniform(-100, 100), rng.uniform(600, 1200)])
        cam = X[i] @ R.T + t
        uv = cam @ K.T
        p2[i] = (uv[:, :2] / uv[:, 2:]).astype(F) + rng.normal(scale=0.3, size=(vn, 2)).astype(F)
    got = pnp_cuda(torch.from_numpy(p2).cuda(), X, np.broadcast_to(K, (n, 3, 3)).copy()).cpu().numpy()
    for i in range(n):
        ref = pnp_np.pnp(X[i], p2[i], K)
        assert np.abs(got[i] - ref).max() < 1e-3, (vn, i)
