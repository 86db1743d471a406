"""Seeded input builders shared by oracle/make_golden.py (which feeds them to the reference's code) and the
golden-vector tests (which feed them to the oracle and the CUDA path).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Cases that do not store their inputs in the .npz store the
builder's arguments and a SHA-256 of what it produced; ``sha`` is checked again by the tests."""
import hashlib

import numpy as np

from casapose_b200 import synthetic

F = np.float32


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes())
    return h.hexdigest()


def mask_from_labels(labels, oc):
    return np.stack([(labels == c + 1) for c in range(oc)], -1).astype(F)


def degenerate_scene():
    """One 40x48 image, 5 classes: empty; 3 pixels (< min_num); a parallel field (det = 0 -> (0,0) hypotheses,
    zero votes, singular AtA -> unrefined winners); a clean field with 10 % zero and 5 % huge vectors; keypoints
    exactly on pixel centres with an exact field (norm_hyp guard)."""
    h, w, oc, vn = 40, 48, 5, 9
    rng = np.random.default_rng(42)
    mask = np.zeros((1, h, w, oc), F)
    vertex = np.zeros((1, h, w, vn, 2), F)
    ys, xs = np.mgrid[0:h, 0:w]
    mask[0, 2, 2:5, 1] = 1
    vertex[0, 2, 2:5] = rng.normal(size=(3, vn, 2))
    sel = (ys >= 4) & (ys < 12) & (xs >= 30) & (xs < 44)
    mask[0, sel, 2] = 1
    vertex[0, sel] = np.array([0.6, 0.8], F)
    sel = (ys >= 16) & (ys < 36) & (xs >= 4) & (xs < 24)
    mask[0, sel, 3] = 1
    kp = rng.uniform([4, 16], [24, 36], size=(vn, 2))  # (x, y)
    py, px = ys[sel] + 0.5, xs[sel] + 0.5
    ang = np.arctan2(kp[None, :, 1] - py[:, None], kp[None, :, 0] - px[:, None])
    ang += np.deg2rad(2.0) * rng.normal(size=ang.shape)
    v = np.stack([np.sin(ang), np.cos(ang)], -1).astype(F)  # (dy, dx)
    r = rng.uniform(size=v.shape[0])
    v[r < 0.10] = 0
    v[(r >= 0.10) & (r < 0.15)] *= F(1e4)
    vertex[0, sel] = v
    sel = (ys >= 16) & (ys < 36) & (xs >= 28) & (xs < 46)
    mask[0, sel, 4] = 1
    kp = np.stack([rng.integers(28, 46, vn) + 0.5, rng.integers(16, 36, vn) + 0.5], 1)
    py, px = ys[sel] + 0.5, xs[sel] + 0.5
    dy, dx = kp[None, :, 1] - py[:, None], kp[None, :, 0] - px[:, None]
    n = np.hypot(dy, dx)
    n[n == 0] = 1
    vertex[0, sel] = np.stack([dy / n, dx / n], -1).astype(F)
    return mask, vertex


def ransac_inputs(b, h, w, ids, variant="easy", seed=synthetic.SEED_BASE):
    d = synthetic.make_frames(b, h, w, tuple(ids), seed=seed, variant=variant)
    return d["mask"], d["vertex"]


def ls_inputs(b, h, w, ids, variant="easy", seed=synthetic.SEED_BASE):
    d = synthetic.make_frames(b, h, w, tuple(ids), seed=seed, variant=variant, with_logits=True)
    return d["seg_logits"], d["vertex"].reshape(b, h, w, 18), d["conf_logits"]


def ls_filter_inputs():
    """Two 64x80 frames whose class 1 carries extra blobs: 30 px (below the 50 px floor of
    voting_layers_2d.py:66) in image 0 and 560 px (larger than the object itself) in image 1."""
    seg, direct, conf = ls_inputs(2, 64, 80, (1, 5, 6))
    seg[0, 2:8, 2:7, :] = 0
    seg[0, 2:8, 2:7, 1] = 12.0
    seg[1, 40:60, 50:78, :] = 0
    seg[1, 40:60, 50:78, 1] = 12.0
    direct[1, 40:60, 50:78] = np.random.default_rng(3).normal(size=(20, 28, 18)).astype(F)
    return seg, direct, conf


def ls_grad_inputs():
    """One 64x80 frame, 3 objects, plus an incoming gradient g [1,3,9,2].  The background carries small random
    vectors instead of exact zeros: at an exactly zero vector TensorFlow's autodiff of sqrt gives NaN, the product
    defines the gradient as 0, and the comparison should not depend on that convention."""
    seg, direct, conf = ls_inputs(1, 64, 80, (1, 5, 6), seed=77)
    rng = np.random.default_rng(77)
    bg = ~direct.any(axis=-1)
    direct[bg] = (0.05 * rng.normal(size=(int(bg.sum()), 18))).astype(F)
    g = rng.normal(size=(1, 3, 9, 2)).astype(F)
    return seg, direct, conf, g


def pose_inputs(b, h, w, ids, variant="easy", crop=(0.0, 0.0)):
    """Arguments of estimate_and_evaluate_poses.  crop = (cx, cy): w_crop = dx = cx and h_crop = dy = cy, which
    cancel in transform_points_back_tf (ransac_voting.py:92-121) only if offsets[0,1,4,5] are read in the
    reference's order (:494-504)."""
    ids = tuple(ids)
    d = synthetic.make_frames(b, h, w, ids, variant=variant, with_logits=True)
    oc = len(ids)
    K = synthetic.camera_matrix(h).astype(F)
    cams = np.broadcast_to(K, (b, 3, 3)).copy()
    offsets = np.zeros((b, 10), F)
    offsets[:, 0], offsets[:, 5] = crop[1], crop[1]  # h_crop, dy
    offsets[:, 1], offsets[:, 4] = crop[0], crop[0]  # w_crop, dx
    offsets[:, 7] = 1.0
    offsets[:, 8], offsets[:, 9] = w, h
    kp3 = np.broadcast_to(d["keypoints_3d"][None, :, None], (b, oc, 1, 9, 3)).copy()
    target_seg = np.concatenate([(d["labels"] == 0)[..., None].astype(F), d["mask"]], axis=-1)
    diam = np.broadcast_to(d["diameters"][None, :, None], (b, oc, 1)).copy()
    return d, cams, offsets, kp3, target_seg, d["poses_gt"][:, :, None].astype(F), diam


def pvnet_inputs():
    """pose_inputs with a PVNet-style vector field [b,h,w,oc*vn*2] — one field per class, wrong (random) everywhere
    except on the class's own pixels (pose_evaluation.py:38-45 gathers the arg-max class's field and zeroes the
    background)."""
    gen = dict(b=1, h=120, w=160, ids=(1, 5, 6, 8))
    d, cams, offsets, kp3, target_seg, poses_gt, diam = pose_inputs(**gen)
    oc = len(gen["ids"])
    rng = np.random.default_rng(21)
    fields = rng.normal(size=(1, gen["h"], gen["w"], oc, 9, 2)).astype(F)
    lab = d["seg_logits"].argmax(-1)  # the classes the pre-step will see
    for c in range(oc):
        sel = lab[0] == c + 1
        fields[0, sel, c] = d["vertex"][0, sel]
    return d, fields.reshape(1, gen["h"], gen["w"], oc * 18), cams, offsets, kp3, target_seg, poses_gt, diam


def unmap_inputs(n=12, vn=9, seed=5):
    """Random keypoints and crop / rotate / scale offsets for map_offsets (ransac_voting.py:487-504); rows 0-1
    are all-zero keypoints (the |sum| < 0.01 guard)."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0, 400, size=(n, vn, 2)).astype(F)
    pts[:2] = 0
    off = np.zeros((n, 10), F)
    off[:, 0] = rng.uniform(0, 40, n)  # h_crop
    off[:, 1] = rng.uniform(0, 60, n)  # w_crop
    off[:, 4] = rng.uniform(-20, 20, n)  # dx
    off[:, 5] = rng.uniform(-20, 20, n)  # dy
    off[:, 6] = rng.uniform(-30, 30, n)  # angle (deg)
    off[:, 7] = rng.uniform(0.5, 2.0, n)  # scale
    off[:, 8], off[:, 9] = 640, 480
    return pts, off


def metric_scene(seed=9):
    """Arguments of evaluate_poses (ransac_voting.py:628-687) on 2 images x 4 objects with evaluation clouds of
    7862 (ADD-S, 'glue'), 3417 (ADD-S, 'eggbox'), 500 and 300 points: estimates straddle the 0.1 * diameter and
    5 px thresholds; one object is absent with a non-zero pose (false positive), one present with a zero pose
    (missing), one absent with a zero pose."""
    rng = np.random.default_rng(seed)
    b, oc, nmax = 2, 4, 7862
    counts = np.array([7862, 3417, 500, 300], np.int32)
    diam = np.array([170.0, 160.0, 120.0, 250.0], F)
    pts = np.zeros((oc, nmax, 3), F)
    for c in range(oc):
        u = rng.normal(size=(counts[c], 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        pts[c, : counts[c]] = (u * (diam[c] / 2) * np.array([1.0, 0.7, 0.4])).astype(F)
    K = synthetic.camera_matrix(480).astype(F)
    cams = np.broadcast_to(K, (b, 3, 3)).copy()
    gt = np.zeros((b, oc, 1, 3, 4), F)
    est = np.zeros((b, oc, 3, 4), F)
    scale = [[0.2, 1.5, 0.9, 0.02], [1.1, 0.3, 3.0, 1.0]]  # pose error in units of 0.1 * diameter, roughly
    for i in range(b):
        for c in range(oc):
            R = synthetic._random_rotation(rng)
            t = np.array([rng.uniform(-100, 100), rng.uniform(-80, 80), rng.uniform(700, 1100)])
            gt[i, c, 0, :, :3], gt[i, c, 0, :, 3] = R, t
            w = rng.normal(size=3)
            w *= 0.02 * scale[i][c] / np.linalg.norm(w)
            wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
            dR = np.eye(3) + wx + 0.5 * wx @ wx
            u_, _, vt = np.linalg.svd(dR @ R)
            est[i, c, :, :3] = u_ @ vt
            est[i, c, :, 3] = t + rng.normal(size=3) * 0.1 * diam[c] * scale[i][c] * 0.6
    valid = np.ones((b, oc), np.int32)
    valid[0, 2] = 0  # absent, pose estimated anyway -> false positive
    valid[1, 3] = 0
    est[1, 3] = 0  # absent, nothing estimated
    est[1, 0] = 0  # present, not found -> missing
    pts_est = rng.uniform(0, 400, size=(b, oc, 9, 2)).astype(F)
    return dict(poses=est, poses_gt=gt, points_estimated=pts_est, evaluation_points=pts, counts=counts.reshape(oc, 1),
                cams=cams, diameters=np.broadcast_to(diam[None, :, None], (b, oc, 1)).copy(), valid=valid)
