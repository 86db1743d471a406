"""Plain-loop restatement of the reference's post-voting PnP and pose metrics.

TEST INFRASTRUCTURE (see oracle/__init__.py); pinned by tests/golden/ (oracle/make_golden.py).
Follows /root/reference/casapose/pose_estimation/ransac_voting.py:13-57 (pnp), :92-121
(transform_points_back_tf), :173-182 (project_tf), :487-558 (map_offsets / map_pnp / estimate_poses),
:561-687 (map_estimates / evaluate_poses) and pose_evaluation.py:11-101.  OpenCV is a third-party
dependency of the reference (opencv-python 4.5.5.62 pinned in requirements.txt:3; this image has 4.13.0):
oracle and product call the same cv2 functions with the same arguments, so PnP parity is by construction
and what is tested is everything around it (guards, un-mapping, ADD / ADD-S, bookkeeping)."""
import math

import cv2
import numpy as np

F = np.float32


def pnp(points_3d, points_2d, camera_matrix):
    if np.abs(np.sum(points_2d)) < 1e-4:  # :17
        return np.zeros([3, 4], F)
    p3 = np.ascontiguousarray(points_3d[None].astype(np.float64))
    p2 = np.ascontiguousarray(points_2d[None].astype(np.float64))
    K = camera_matrix.astype(np.float64)
    _, rvec0, T0, _ = cv2.solvePnPRansac(p3, p2, K, None, flags=cv2.SOLVEPNP_EPNP, confidence=0.9999, reprojectionError=12)
    ret, R_exp, t = cv2.solvePnP(p3, p2, K, None, flags=cv2.SOLVEPNP_ITERATIVE, useExtrinsicGuess=True, rvec=rvec0, tvec=T0)
    if ret is False or np.isnan(np.sum(t)):
        return np.zeros([3, 4], F)
    R, _ = cv2.Rodrigues(R_exp)
    if t[2] < 0:
        t, R = -t, -R
    return np.concatenate([R, t], axis=-1).astype(F)


def transform_points_back(points, h_crop, w_crop, sx, sy, dx, dy, angle, scale):
    out = np.zeros_like(points, dtype=F)
    ang = F(-angle) * F(math.pi / 180)
    a, b = F(np.cos(ang)), F(np.sin(ang))
    cx, cy = F(sx) / F(2), F(sy) / F(2)
    c = (F(1) - a) * cx - b * cy
    d = b * cx + (F(1) - a) * cy
    for i, (x, y) in enumerate(points.astype(F)):
        x = F(x / F(scale)) + F(w_crop)
        y = F(y / F(scale)) + F(h_crop)
        x, y = F(x - F(dx)), F(y - F(dy))  # tm
        out[i] = (F(a * x + b * y) + c, F(-b * x + a * y) + d)  # rm
    return out


def project(xyz, K, RT):
    cam = xyz.astype(F) @ RT[:, :3].astype(F).T + RT[:, 3].astype(F)
    uvw = cam @ K.astype(F).T
    with np.errstate(divide="ignore", invalid="ignore"):
        return (uvw[:, :2] / uvw[:, 2:]).astype(F), cam.astype(F)


def estimate_poses(points, keypoints, cams, valid, offsets):
    b, oc = points.shape[:2]
    poses = np.zeros((b, oc, 3, 4), F)
    fp = np.zeros(oc, F)
    for i in range(b):
        for c in range(oc):
            pts = points[i, c].astype(F)
            if valid[i, c] == 0 and pts.sum(dtype=F) > 0:
                fp[c] += 1
            if abs(pts.sum(dtype=F)) < 0.01:
                continue
            o = offsets[i]
            pts = transform_points_back(pts, o[0], o[1], o[8], o[9], o[4], o[5], o[6], o[7])
            if abs(pts.sum(dtype=F)) < 0.01:
                continue
            poses[i, c] = pnp(keypoints[i, c, 0], pts, cams[i])
    return poses, fp


def evaluate_poses(poses, poses_gt, pts3d, counts, cams, diameters, valid, allowed_error_2d=5.0):
    b, oc = poses.shape[:2]
    out = {k: np.zeros(oc, F) for k in ("err_2d", "err_3d", "valid_2d", "valid_3d", "missing", "false_positive")}
    flags = np.zeros((b, oc, 2), F)
    for i in range(b):
        for c in range(oc):
            pose = poses[i, c]
            if valid[i, c] == 0:
                if abs(pose.sum(dtype=F)) > 0.0001:
                    out["false_positive"][c] += 1
                continue
            if abs(pose.sum(dtype=F)) < 0.0001:
                out["err_2d"][c] += F(99.9)
                out["err_3d"][c] += F(999.9)
                out["missing"][c] += 1
                continue
            n = int(counts[i, c, 0])
            pts = pts3d[i, c, 0][:n]
            p2, p3 = project(pts, cams[i], pose)
            t2, t3 = project(pts, cams[i], poses_gt[i, c, 0])
            e2 = F(np.linalg.norm(t2 - p2, axis=1).mean())
            if n in (7862, 3417):
                A, B = t3.astype(np.float64), p3.astype(np.float64)
                dist = np.array([np.min(((a[None] - B) ** 2).sum(1)) for a in A])  # exact closest-point form
                e3 = F(np.sqrt(np.abs(dist) + 1e-5).astype(F).mean())
            else:
                e3 = F(np.linalg.norm(t3 - p3, axis=1).mean())
            v3 = F(e3 < F(diameters[i, c]) * F(0.1))
            v2 = F(e2 < allowed_error_2d)
            out["err_2d"][c] += e2
            out["err_3d"][c] += e3
            out["valid_3d"][c] += v3
            out["valid_2d"][c] += v2
            flags[i, c] = (v2, v3)
    out["valid_count"] = valid.sum(axis=0).astype(F)
    out["flags"] = flags
    return out
