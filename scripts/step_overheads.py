"""Where the non-scoring time of a config-2 step goes: ms per step (CUDA events around 100 queued asynchronous calls)
with the seed varying / fixed (fixed = no kernel node is re-patched), k_score timing events on / off, and with / without
the WHILE node (max_iter 1).  usage: python scripts/step_overheads.py [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import _lib, synthetic  # noqa: E402
from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
gen = int(sys.argv[2]) if len(sys.argv) > 2 else 4  # distinct frames (bench.py config 2 generates 16)
d = synthetic.make_frames(gen, 480, 640, synthetic.CONFIG_8_IDS, seed=synthetic.SEED_BASE, variant="easy")
mask = torch.from_numpy(np.tile(d["mask"], (16 // gen, 1, 1, 1))).cuda()
vertex = torch.from_numpy(np.tile(d["vertex"], (16 // gen, 1, 1, 1, 1))).cuda()
lib = _lib.lib()
hdl = _lib.handle(0, torch.cuda.current_stream().cuda_stream)
out = torch.empty((16, 8, 9, 2), device="cuda")


def run(label, vary_seed, timing, max_iter, mode=1):
    _lib.check(lib.casa_set_async(hdl, mode))
    _lib.check(lib.casa_set_timing(hdl, timing))
    for it in range(5):
        ransac_voting_layer_all_masks(mask, vertex, 512, max_iter=max_iter, seed=it if vary_seed else 7, out=out)
    _lib.check(lib.casa_sync(hdl))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(steps):
        ransac_voting_layer_all_masks(mask, vertex, 512, max_iter=max_iter, seed=100 + it if vary_seed else 7, out=out)
    if mode == 2:
        _lib.check(lib.casa_join(hdl, torch.cuda.current_stream().cuda_stream))
    e1.record()
    torch.cuda.synchronize()
    _lib.check(lib.casa_sync(hdl))
    sm, sl, st = C.c_double(), C.c_int64(), (C.c_uint64 * 4)()
    lib.casa_get_timing(hdl, C.byref(sm), C.byref(sl), st)
    ms = e0.elapsed_time(e1) / steps
    score = sm.value / sl.value if sl.value else float("nan")
    print("%-44s %.4f ms/step   k_score %.4f   rest %.4f" % (label, ms, score, ms - score if sl.value else float("nan")))


run("seed varies, timing on, max_iter 20", True, 1, 20)
if len(sys.argv) > 3:  # NVML polling beside the loop, like bench.py's ClockSampler
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    smp = bench.ClockSampler(0)
    smp.start()
    import time
    time.sleep(0.5)
    run("  ... with the NVML clock sampler (5 ms)", True, 1, 20)
    smp.stop()
run("seed fixed,  timing on, max_iter 20", False, 1, 20)
run("seed fixed,  timing off, max_iter 20", False, 0, 20)
run("seed varies, timing off, max_iter 20", True, 0, 20)
run("seed fixed,  timing off, max_iter 1", False, 0, 1)
run("seed varies, timing on, max_iter 1", True, 1, 1)
if hasattr(lib, "casa_join"):
    run("two lanes: seed varies, timing on, max_iter 20", True, 1, 20, mode=2)
    run("two lanes: seed varies, timing off, max_iter 20", True, 0, 20, mode=2)
_lib.check(lib.casa_set_async(hdl, 0))
_lib.check(lib.casa_set_timing(hdl, 0))
