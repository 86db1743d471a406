"""Seeded synthetic "LM-O-shaped" frames for tests and benchmarks (no dataset, no network).

What the real pipeline would hand to the voting path is the network output split of
/root/reference/test_casapose.py:302-304: segmentation logits [b,h,w,1+oc], a shared
2*9-channel unit-vector field and 9 confidence logits.  This module fabricates tensors of
that shape whose statistics resemble LM-O frames:

* objects: the LINEMOD ids of config_8.ini:13 / config_13.ini:13 with their 9 3-D
  keypoints (mm), diameters and bounding boxes (casapose_b200/data/lm_models.json,
  extracted from the reference's data files by scripts/make_lm_fixture.py);
* camera: the public BOP LINEMOD intrinsics, scaled with the image height;
* pose: uniform random rotation, t_z ~ U[600,1200] mm, centre projected inside the image;
* mask: every object's bbox-inscribed ellipsoid is ray-cast and painted far-to-near, so
  nearer objects occlude farther ones (LM-O-like);
* vector field: unit vector from the pixel centre (x+.5, y+.5) to the projected keypoint,
  stored (dy, dx) at channels [2k, 2k+1] (/root/reference/casapose/utils/image_utils.py:29-63),
  zero on background; "easy": 80 % of the pixels get Gaussian angular noise sigma = 3 deg and
  20 % a uniformly random direction; "hard": sigma = 20 deg, 60 % random; "multi": sigma = 35 deg, 85 % random
  (the best inlier ratio stays below 0.0946, so the stop test of ransac_voting.py:344-347 asks for further rounds).
"""
import json
import os

import numpy as np

CONFIG_8_IDS = (1, 5, 6, 8, 9, 10, 11, 12)  # config_8.ini:13
CONFIG_13_IDS = (1, 2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15)  # config_13.ini:13
LM_K_480 = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]])
SEED_BASE = 1237  # config_8.ini:22 manualseed

VARIANTS = {"easy": (3.0, 0.2), "hard": (20.0, 0.6), "multi": (35.0, 0.85), "clean": (0.0, 0.0)}

_MODELS = None


def lm_models():
    global _MODELS
    if _MODELS is None:
        path = os.path.join(os.path.dirname(__file__), "data", "lm_models.json")
        with open(path) as f:
            _MODELS = json.load(f)
    return _MODELS


def camera_matrix(h):
    s = h / 480.0
    k = LM_K_480.copy()
    k[:2] *= s
    return k


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def make_frame(h, w, object_ids, rng, variant="easy", depth_range=(600.0, 1200.0)):
    """One frame.  Returns labels [h,w] uint8 (0 = background, c+1 = class c), vertex [h,w,18] f32,
    poses [oc,3,4] f64, kp2d [oc,9,2] f64 (x,y)."""
    sigma_deg, outlier_frac = VARIANTS[variant]
    models = lm_models()
    oc = len(object_ids)
    K = camera_matrix(h)
    Kinv = np.linalg.inv(K)
    xs, ys = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5)
    rays = np.stack([xs, ys, np.ones_like(xs)], -1) @ Kinv.T  # [h,w,3]

    poses = np.zeros((oc, 3, 4))
    kp2d = np.zeros((oc, 9, 2))
    hits = []
    for c, oid in enumerate(object_ids):
        m = models["obj_%06d" % oid]
        R = _random_rotation(rng)
        tz = rng.uniform(*depth_range)
        u = rng.uniform(0.1 * w, 0.9 * w)
        v = rng.uniform(0.1 * h, 0.9 * h)
        t = tz * (Kinv @ np.array([u, v, 1.0]))
        poses[c, :, :3] = R
        poses[c, :, 3] = t
        kp = np.asarray(m["keypoints"])  # [9,3]
        cam = kp @ R.T + t
        pr = cam @ K.T
        kp2d[c] = pr[:, :2] / pr[:, 2:3]
        # ray-cast the bbox-inscribed ellipsoid
        axes = np.asarray(m["size"]) / 2.0
        centre = np.asarray(m["min"]) + axes
        o = (R.T @ (-t)) - centre  # camera origin in the ellipsoid frame
        d = rays @ R  # = R^T r for every pixel
        oq = o / axes
        dq = d / axes
        a = (dq * dq).sum(-1)
        bq = (dq * oq).sum(-1)
        cq = (oq * oq).sum() - 1.0
        hits.append((tz, c, (bq * bq - a * cq) >= 0.0))
    labels = np.zeros((h, w), np.uint8)
    for _, c, hit in sorted(hits, key=lambda e: -e[0]):  # far to near
        labels[hit] = c + 1

    vertex = np.zeros((h, w, 18), np.float32)
    fg = labels > 0
    cls = labels[fg].astype(np.int64) - 1
    px = xs[fg]
    py = ys[fg]
    n = cls.shape[0]
    target = kp2d[cls]  # [n,9,2]
    ang = np.arctan2(target[..., 1] - py[:, None], target[..., 0] - px[:, None])  # [n,9]
    if sigma_deg > 0:
        ang = ang + np.deg2rad(sigma_deg) * rng.normal(size=ang.shape)
    if outlier_frac > 0:
        outl = rng.uniform(size=n) < outlier_frac
        ang[outl] = rng.uniform(-np.pi, np.pi, size=(int(outl.sum()), 9))
    vf = np.empty((n, 18), np.float32)
    vf[:, 0::2] = np.sin(ang)  # dy
    vf[:, 1::2] = np.cos(ang)  # dx
    vertex[fg] = vf
    return labels, vertex, poses, kp2d


def make_frames(b, h=480, w=640, object_ids=CONFIG_8_IDS, seed=SEED_BASE, variant="easy", with_logits=False):
    """Batch of frames as float32 numpy arrays in the layouts the reference's entry points take."""
    oc = len(object_ids)
    mask = np.zeros((b, h, w, oc), np.float32)
    vertex = np.zeros((b, h, w, 18), np.float32)
    labels = np.zeros((b, h, w), np.uint8)
    poses = np.zeros((b, oc, 3, 4), np.float32)
    kp2d = np.zeros((b, oc, 9, 2), np.float32)
    for i in range(b):
        rng = np.random.Generator(np.random.PCG64([seed, i]))
        lab, vf, ps, k2 = make_frame(h, w, object_ids, rng, variant)
        labels[i] = lab
        vertex[i] = vf
        poses[i] = ps
        kp2d[i] = k2
        for c in range(oc):
            mask[i, :, :, c] = lab == c + 1
    models = lm_models()
    out = {
        "mask": mask,
        "vertex": vertex.reshape(b, h, w, 9, 2),
        "labels": labels,
        "poses_gt": poses,
        "kp2d_gt": kp2d,
        "keypoints_3d": np.array([models["obj_%06d" % o]["keypoints"] for o in object_ids], np.float32),
        "diameters": np.array([models["obj_%06d" % o]["diameter"] for o in object_ids], np.float32),
        "camera": camera_matrix(h).astype(np.float32),
    }
    if with_logits:
        rng = np.random.Generator(np.random.PCG64([seed, 1 << 20]))
        seg = rng.normal(size=(b, h, w, oc + 1)).astype(np.float32)
        onehot = np.eye(oc + 1, dtype=np.float32)[labels]
        out["seg_logits"] = seg + 10.0 * onehot
        out["conf_logits"] = rng.normal(size=(b, h, w, 9)).astype(np.float32)
    return out
