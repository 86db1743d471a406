"""A few config-2 steps for profiling (ncu wraps this; numbers printed under ncu are not bench values)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import synthetic  # noqa: E402
from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
d = synthetic.make_frames(4, 480, 640, synthetic.CONFIG_8_IDS, variant="easy")
mask = torch.from_numpy(np.tile(d["mask"], (4, 1, 1, 1))).cuda()
vertex = torch.from_numpy(np.tile(d["vertex"], (4, 1, 1, 1, 1))).cuda()
for it in range(steps):
    ransac_voting_layer_all_masks(mask, vertex, 512, seed=it)
torch.cuda.synchronize()
print("done")
