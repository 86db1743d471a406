"""Timeline of the kernels of consecutive config-2 votes issued on n lanes (casa_set_async(h, n)): CUPTI activity records
(through torch.profiler, which sees every kernel of the process incl. graph kernel nodes) of the steady state, printed as
one line per kernel: start and end in microseconds, stream, name.  Design exploration only.
usage: python scripts/trace_lanes.py [lanes] [traced calls] [out file]"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import _lib, synthetic  # noqa: E402
from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks  # noqa: E402

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out_path = sys.argv[3] if len(sys.argv) > 3 else None
d = synthetic.make_frames(4, 480, 640, synthetic.CONFIG_8_IDS, seed=synthetic.SEED_BASE, variant="easy")
mask = torch.from_numpy(np.tile(d["mask"], (4, 1, 1, 1))).cuda()
vertex = torch.from_numpy(np.tile(d["vertex"], (4, 1, 1, 1, 1))).cuda()
lib = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
hdl = _lib.handle(0, stream)
outs = [torch.empty((16, 8, 9, 2), device="cuda") for _ in range(8)]
_lib.check(lib.casa_set_async(hdl, lanes))


def issue(n, seed0):
    for it in range(n):
        ransac_voting_layer_all_masks(mask, vertex, 512, seed=seed0 + it, out=outs[it % len(outs)])


issue(12, 0)
_lib.check(lib.casa_sync(hdl))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    issue(calls, 100)
    if lanes >= 2:
        _lib.check(lib.casa_join(hdl, stream))
    torch.cuda.synchronize()
_lib.check(lib.casa_sync(hdl))
_lib.check(lib.casa_set_async(hdl, 0))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start if ev else 0
lines = []
for e in ev:
    name = e.name.split("(")[0].replace("void ", "").replace("casa::", "")
    lines.append("%9.1f %9.1f %7.1f  s%-3s %s" % (e.time_range.start - t0, e.time_range.end - t0,
                                                 e.time_range.end - e.time_range.start, getattr(e, "device_resource_id", "?"), name))
txt = "\n".join(lines)
if out_path:
    with open(out_path, "w") as f:
        f.write(txt + "\n")
print(txt[-6000:])
