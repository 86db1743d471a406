"""Host-side post-step (PnP, un-mapping, ADD / ADD-S bookkeeping): product functions vs the loop oracle.
These run on the CPU in both (the reference itself leaves TensorFlow for numpy + OpenCV here)."""
import numpy as np

from casapose_b200 import synthetic
from casapose_b200.pose_estimation import ransac_voting as RV
from casapose_b200.pose_estimation.bpnp_layers import rodrigues_batch
from oracle import pose_np as OP

F = np.float32


def _scene(b=3, ids=synthetic.CONFIG_8_IDS, seed=5):
    d = synthetic.make_frames(b, 60, 80, ids, seed=seed, variant="clean")  # small image: only poses / keypoints used
    oc = len(ids)
    K = synthetic.camera_matrix(480).astype(F)
    rng = np.random.default_rng(seed)
    kp3 = d["keypoints_3d"]
    poses = d["poses_gt"].copy()
    poses[:, :, :, 3] *= 1.0
    kp2 = np.zeros((b, oc, 9, 2), F)
    for i in range(b):
        for c in range(oc):
            cam = kp3[c] @ poses[i, c, :, :3].T + poses[i, c, :, 3]
            uvw = cam @ K.T
            kp2[i, c] = uvw[:, :2] / uvw[:, 2:]
    kp2 += rng.normal(scale=0.3, size=kp2.shape).astype(F)
    cams = np.broadcast_to(K, (b, 3, 3)).copy()
    offsets = np.zeros((b, 10), F)
    offsets[:, 7] = 1.0  # scale
    offsets[:, 8], offsets[:, 9] = 640, 480
    keypoints = np.broadcast_to(kp3[None, :, None], (b, oc, 1, 9, 3)).copy()
    return kp2, keypoints, cams, offsets, poses[:, :, None].astype(F), d["diameters"]


def test_pnp_recovers_pose_and_matches_oracle():
    kp2, keypoints, cams, offsets, poses_gt, _ = _scene()
    valid = np.ones(kp2.shape[:2], np.int32)
    poses, fp = RV.estimate_poses(kp2, keypoints, cams, valid, offsets)
    ref, rfp = OP.estimate_poses(kp2, keypoints, cams, valid, offsets)
    assert np.array_equal(poses, ref) and np.array_equal(fp, rfp)
    assert np.abs(poses[..., 3] - poses_gt[:, :, 0, :, 3]).max() < 25.0  # mm, 0.3 px keypoint noise


def test_guards_zero_points_and_false_positives():
    kp2, keypoints, cams, offsets, poses_gt, _ = _scene(b=2)
    kp2[0, 1] = 0.0  # voting returned zeros (gated class) -> zero pose
    valid = np.ones(kp2.shape[:2], np.int32)
    valid[1, 2] = 0  # object not in the ground truth but keypoints found -> false positive
    poses, fp = RV.estimate_poses(kp2, keypoints, cams, valid, offsets)
    ref, rfp = OP.estimate_poses(kp2, keypoints, cams, valid, offsets)
    assert np.array_equal(poses, ref) and np.array_equal(fp, rfp)
    assert np.array_equal(poses[0, 1], np.zeros((3, 4), F)) and fp[2] == 1 and fp.sum() == 1


def test_offsets_unmapping_matches_oracle():
    rng = np.random.default_rng(1)
    pts = rng.uniform(0, 400, size=(9, 2)).astype(F)
    args = (12.0, 30.0, 640.0, 480.0, 5.0, -3.0, 17.0, 0.8)
    a = RV.transform_points_back(pts, *args)
    b = OP.transform_points_back(pts, *args)
    assert np.abs(a - b).max() < 1e-3
    ident = RV.transform_points_back(pts, 0, 0, 640, 480, 0, 0, 0, 1.0)
    assert np.abs(ident - pts).max() < 1e-4


def test_add_and_adds_metrics_match_oracle():
    kp2, keypoints, cams, offsets, poses_gt, diam = _scene(b=2)
    b, oc = kp2.shape[:2]
    valid = np.ones((b, oc), np.int32)
    valid[0, 3] = 0
    poses, _ = RV.estimate_poses(kp2, keypoints, cams, valid, offsets)
    poses[1, 4] = 0.0  # a missed object
    rng = np.random.default_rng(3)
    n_pts = 3417  # eggbox vertex count -> ADD-S branch (:618)
    ev = rng.uniform(-40, 40, size=(oc, n_pts, 3)).astype(F)
    pts3d = np.broadcast_to(ev[None, :, None], (b, oc, 1, n_pts, 3)).copy()
    for cnt_val in (3417, 9):
        counts = np.full((b, oc, 1), cnt_val, np.int32)
        dm = np.broadcast_to(diam[None], (b, oc)).copy()
        e2, e3, v2, v3, miss, vcount, fpp = RV.evaluate_poses(poses, poses_gt, kp2, pts3d, counts, cams, dm, valid, 5.0)
        ref = OP.evaluate_poses(poses, poses_gt, pts3d, counts, cams, dm, valid, 5.0)
        assert np.allclose(e2, ref["err_2d"], rtol=1e-4, atol=1e-3) and np.allclose(e3, ref["err_3d"], rtol=1e-4, atol=1e-3)
        assert np.array_equal(v2, ref["valid_2d"]) and np.array_equal(v3, ref["valid_3d"])  # identical ADD(-S) verdicts
        assert np.array_equal(miss, ref["missing"]) and miss[4] == 1
        assert np.array_equal(vcount, ref["valid_count"]) and np.array_equal(fpp, ref["false_positive"])


def test_rodrigues_batch():
    import cv2

    rng = np.random.default_rng(0)
    r = rng.normal(size=(5, 3)).astype(F)
    r[2] = 0
    R = rodrigues_batch(r)
    for i in range(5):
        assert np.abs(R[i] - cv2.Rodrigues(r[i].astype(np.float64))[0]).max() < 1e-5
