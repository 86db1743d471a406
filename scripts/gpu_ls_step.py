"""A few LS-layer steps for profiling (ncu wraps this).  usage: gpu_ls_step.py [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import synthetic  # noqa: E402
from casapose_b200.pose_estimation import CoordLSVotingWeighted  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, H, W, OC, VN = 16, 480, 640, 8, 9
d = synthetic.make_frames(4, H, W, synthetic.CONFIG_8_IDS, variant="easy", with_logits=True)
rep = B // 4
seg = torch.from_numpy(np.tile(d["seg_logits"], (rep, 1, 1, 1))).cuda()
direct = torch.from_numpy(np.tile(d["vertex"].reshape(4, H, W, 2 * VN), (rep, 1, 1, 1))).cuda()
conf = torch.from_numpy(np.tile(d["conf_logits"], (rep, 1, 1, 1))).cuda()
layer = CoordLSVotingWeighted("ls", OC + 1, num_points=VN, filter_estimates=True)
for _ in range(steps):
    layer([seg, direct, conf], check_finite=False)
torch.cuda.synchronize()
print("done")
