"""Small end-to-end invocations for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import synthetic  # noqa: E402
from casapose_b200.pose_estimation import CoordLSVotingWeighted, ransac_voting_layer_all_masks  # noqa: E402
from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks_host  # noqa: E402

d = synthetic.make_frames(2, 96, 128, (1, 5, 6), variant="hard", with_logits=True)
m, v = torch.from_numpy(d["mask"]).cuda(), torch.from_numpy(d["vertex"]).cuda()
out, dbg = ransac_voting_layer_all_masks(m, v, 64, seed=1, max_iter=3, return_debug=True)
out2 = ransac_voting_layer_all_masks(m, v, 64, seed=1, max_iter=3, max_num=150)
out3 = ransac_voting_layer_all_masks(torch.from_numpy(d["seg_logits"]).cuda(), v, 64, seed=1, max_iter=2, seg_scores=True)
out4 = ransac_voting_layer_all_masks(m, v, 32, seed=1, max_iter=1, force_exact=True)
host = ransac_voting_layer_all_masks_host(torch.from_numpy(d["mask"]).pin_memory(), torch.from_numpy(d["vertex"]).pin_memory(), 64, seed=1, max_iter=3)
layer = CoordLSVotingWeighted("ls", 4, num_points=9, filter_estimates=True)
ls = layer([torch.from_numpy(d["seg_logits"]).cuda(), torch.from_numpy(d["vertex"].reshape(2, 96, 128, 18)).cuda(),
            torch.from_numpy(d["conf_logits"]).cuda()])
# section-8(f) kernels: LS backward, batched PnP, ADD / ADD-S (one model with the symmetric meshes' vertex count)
from casapose_b200.pose_estimation.ransac_voting import pnp_cuda, pose_errors_cuda  # noqa: E402

seg_t = torch.from_numpy(d["seg_logits"]).cuda()
dir_t = torch.from_numpy(d["vertex"].reshape(2, 96, 128, 18)).cuda()
conf_t = torch.from_numpy(d["conf_logits"]).cuda()
gd, gw = layer.backward([seg_t, dir_t, conf_t], torch.randn((2, 3, 9, 2), device="cuda"))
ls_plain = CoordLSVotingWeighted("ls2", 4, num_points=9, sigmoid_weights=True)([seg_t, dir_t, conf_t])
rng = np.random.default_rng(0)
K = synthetic.camera_matrix(96).astype(np.float32)
kp3 = d["keypoints_3d"].astype(np.float32)  # [oc,9,3]
RT = d["poses_gt"][0].astype(np.float32)    # [oc,3,4]
cam = kp3 @ RT[:, :, :3].transpose(0, 2, 1) + RT[:, None, :, 3]
uv = cam @ K.T
uv = (uv[..., :2] / uv[..., 2:]).astype(np.float32) + rng.normal(scale=0.5, size=(3, 9, 2)).astype(np.float32)
cams = np.broadcast_to(K, (3, 3, 3)).copy()
poses = pnp_cuda(torch.from_numpy(uv).cuda(), kp3, cams, np.tile(np.array([0, 0, 0, 0, 0, 0, 0, 1, 128, 96], np.float32), (3, 1)))
counts = np.array([700, 3417, 9], np.int32)
pts = (rng.uniform(-0.5, 0.5, size=(3, 3417, 3)) * 80).astype(np.float32)
rows = pose_errors_cuda(poses, RT, cams, pts, counts, np.full(3, 100.0, np.float32), np.array([1, 1, 0], np.int32), 5.0)
# round 2: lanes (three votes / LS calls in flight, joined on the caller's stream), the run-based components with more
# runs than shared memory holds, overlapping mask channels (call repeated with room for every channel)
from casapose_b200 import _lib  # noqa: E402

_lib.set_async(0, 3)
lane_outs = [ransac_voting_layer_all_masks(m, v, 64, seed=s, max_iter=3) for s in (1, 2, 3, 4)]
lane_ls = [layer([seg_t, dir_t, conf_t], check_finite=False) for _ in range(3)]
_lib.join(0)
_lib.sync(0)
_lib.set_async(0, 0)
assert torch.equal(lane_outs[0], out) and torch.equal(lane_ls[0], ls)
noise = rng.integers(0, 3, size=(96, 128))
seg_n = np.zeros((1, 96, 128, 3), np.float32)
for c in range(3):
    seg_n[0, :, :, c] = np.where(noise == c, 3.0, 0.0)
ls_noise = CoordLSVotingWeighted("ls3", 3, num_points=9, filter_estimates=True)(
    [torch.from_numpy(seg_n).cuda(), dir_t[:1].contiguous(), conf_t[:1].contiguous()])
multi = np.maximum(d["mask"], d["mask"][..., ::-1].copy())  # every channel also holds another one's pixels
out5 = ransac_voting_layer_all_masks(torch.from_numpy(multi).cuda(), v, 32, seed=1, max_iter=2)
# session 3 of round 2: the pipelined host entry (driver threads, shared packer: three calls queued, two in flight)
mh, vh = torch.from_numpy(d["mask"]).pin_memory(), torch.from_numpy(d["vertex"]).pin_memory()
pend = [ransac_voting_layer_all_masks_host(mh, vh, 64, seed=1, max_iter=3, wait=False) for _ in range(3)]
for pnd in pend:
    assert torch.equal(pnd.result(), host)
torch.cuda.synchronize()
assert torch.isfinite(ls_noise).all() and torch.isfinite(out5).all()
assert torch.isfinite(gd).all() and torch.isfinite(gw).all() and torch.isfinite(rows).all() and torch.isfinite(ls_plain).all()
assert torch.equal(out.cpu(), host)
print("sanitize run ok", float(out.abs().sum()), float(ls.abs().sum()))
