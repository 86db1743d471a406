"""A few config-2 steps for profiling (ncu wraps this; numbers printed under ncu are not bench values).
usage: gpu_step.py [steps] [device|host]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import synthetic  # noqa: E402
from casapose_b200.pose_estimation.ransac_voting import (ransac_voting_layer_all_masks,  # noqa: E402
                                                           ransac_voting_layer_all_masks_host)

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "device"
d = synthetic.make_frames(4, 480, 640, synthetic.CONFIG_8_IDS, variant="easy")
mask_h = torch.from_numpy(np.tile(d["mask"], (4, 1, 1, 1))).pin_memory()
vertex_h = torch.from_numpy(np.tile(d["vertex"], (4, 1, 1, 1, 1))).pin_memory()
if mode == "host":
    for it in range(steps):
        ransac_voting_layer_all_masks_host(mask_h, vertex_h, 512, seed=it)
else:
    mask, vertex = mask_h.cuda(), vertex_h.cuda()
    for it in range(steps):
        ransac_voting_layer_all_masks(mask, vertex, 512, seed=it)
torch.cuda.synchronize()
print("done")
