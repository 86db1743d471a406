"""Timing of the section-8(f) kernels at BASELINE-config shapes (CUDA events; also the command ncu wraps for
profiles/r01_next_rows_launches.csv — numbers printed under ncu are not bench values).
usage: gpu_next_rows.py [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import synthetic  # noqa: E402
from casapose_b200.pose_estimation import CoordLSVotingWeighted  # noqa: E402
from casapose_b200.pose_estimation.ransac_voting import pnp_cuda, pose_errors_cuda  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, H, W, OC, VN = 16, 480, 640, 8, 9


def timed(fn, n=reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


d = synthetic.make_frames(4, H, W, synthetic.CONFIG_8_IDS, variant="easy", with_logits=True)
rep = B // 4
seg = torch.from_numpy(np.tile(d["seg_logits"], (rep, 1, 1, 1))).cuda()
direct = torch.from_numpy(np.tile(d["vertex"].reshape(4, H, W, 2 * VN), (rep, 1, 1, 1))).cuda()
conf = torch.from_numpy(np.tile(d["conf_logits"], (rep, 1, 1, 1))).cuda()
layer = CoordLSVotingWeighted("ls", OC + 1, num_points=VN, filter_estimates=True)
g = torch.randn((B, OC, VN, 2), device="cuda")
fwd = timed(lambda: layer([seg, direct, conf], check_finite=False))
bwd = timed(lambda: layer.backward([seg, direct, conf], g))
print("LS layer  b=%d: forward %.3f ms, backward (incl. forward recomputation) %.3f ms" % (B, fwd, bwd))

# PnP: 128 objects, keypoints = projected model keypoints + 1 px noise
rng = np.random.default_rng(0)
K = synthetic.camera_matrix(H).astype(np.float32)
kp3 = np.tile(d["keypoints_3d"][None], (B, 1, 1, 1)).reshape(B * OC, VN, 3).astype(np.float32)
poses_gt = np.tile(d["poses_gt"], (rep, 1, 1, 1)).reshape(B * OC, 3, 4).astype(np.float32)
cam = kp3 @ poses_gt[:, :, :3].transpose(0, 2, 1) + poses_gt[:, None, :, 3]
uv = cam @ K.T
uv = (uv[..., :2] / uv[..., 2:]).astype(np.float32) + rng.normal(scale=1.0, size=(B * OC, VN, 2)).astype(np.float32)
p2, p3 = torch.from_numpy(uv).cuda(), torch.from_numpy(kp3).cuda()
cams = torch.from_numpy(np.broadcast_to(K, (B * OC, 3, 3)).copy()).cuda()
poses = pnp_cuda(p2, p3, cams)
pnp_ms = timed(lambda: pnp_cuda(p2, p3, cams))
t0 = time.perf_counter()
from casapose_b200.pose_estimation.ransac_voting import pnp as pnp_cv2  # noqa: E402
for i in range(B * OC):
    pnp_cv2(kp3[i], uv[i], K)
cv_ms = (time.perf_counter() - t0) * 1e3
print("PnP  %d objects: GPU %.3f ms, host OpenCV sequence %.1f ms" % (B * OC, pnp_ms, cv_ms))

# ADD / ADD-S: 8 models, two of them with the symmetric meshes' vertex counts
counts = np.array([5841, 7862, 5000, 4000, 3417, 6000, 5500, 4500], np.int32)
maxp = int(counts.max())
pts = (rng.uniform(-0.5, 0.5, size=(OC, maxp, 3)) * 100).astype(np.float32)
diam = torch.full((B * OC,), 170.0, device="cuda")
valid = torch.ones((B * OC,), dtype=torch.int32, device="cuda")
tp, tc = torch.from_numpy(pts).cuda(), torch.from_numpy(counts).cuda()
gt = torch.from_numpy(poses_gt).cuda()
err_ms = timed(lambda: pose_errors_cuda(poses, gt, cams, tp, tc, diam, valid, 5.0))
print("ADD / ADD-S  %d objects (%d symmetric, up to %d points): %.3f ms" % (B * OC, 2 * B, maxp, err_ms))
