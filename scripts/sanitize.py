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
torch.cuda.synchronize()
assert torch.equal(out.cpu(), host)
print("sanitize run ok", float(out.abs().sum()), float(ls.abs().sum()))
