"""First on-GPU shake-down: filter self-test, FP32 peak, parity vs oracle, rough timing."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import _lib, synthetic  # noqa: E402
from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks  # noqa: E402
from oracle import ransac_voting_np as O  # noqa: E402

out = {}
L = _lib.lib()
h = _lib.handle(0)
for v in range(4):
    tf, ms = C.c_double(), C.c_double()
    _lib.check(L.casa_measure_fp32_peak(h, v, C.byref(tf), C.byref(ms)))
    out["fp32_peak_v%d" % v] = (tf.value, ms.value)
    print("fma variant", v, tf.value, "TFLOP/s", ms.value, "ms", flush=True)
for spread in (2e-5, 1e-3, 0.3):
    res = (C.c_uint64 * 4)()
    _lib.check(L.casa_selftest_filter(h, 1 << 28, 1234, 0.99, spread, res))
    out["selftest_%g" % spread] = list(res)
    print("selftest spread", spread, list(res), flush=True)


def compare(name, d, hn, **kw):
    mask = torch.from_numpy(d["mask"]).cuda()
    vertex = torch.from_numpy(d["vertex"]).cuda()
    pts, dbg = ransac_voting_layer_all_masks(mask, vertex, hn, return_debug=True, seed=7, **kw)
    torch.cuda.synchronize()
    okw = {k: v for k, v in kw.items() if k in ("inlier_thresh", "confidence", "max_iter", "min_num", "max_num")}
    ref, rdbg = O.ransac_voting_layer_all_masks(d["mask"], d["vertex"], hn, seed=7, return_debug=True, **okw)
    b, oc = d["mask"].shape[0], d["mask"].shape[3]
    bad = 0
    for i in range(b):
        for c in range(oc):
            r = rdbg[i][c]
            g_tn = int(dbg["tn"][i, c]); g_r = int(dbg["rounds"][i, c])
            if g_tn != r["tn"] or g_r != r["rounds"]:
                print(name, "MISMATCH tn/rounds", i, c, g_tn, r["tn"], g_r, r["rounds"]); bad += 1; continue
            for k in range(r["rounds"]):
                gc = dbg["counts"][i, c, k].cpu().numpy()
                if not np.array_equal(gc, r["counts"][k]):
                    nd = int((gc != r["counts"][k]).sum())
                    print(name, "COUNT MISMATCH", i, c, k, nd, "max diff", np.abs(gc - r["counts"][k]).max()); bad += 1
                if not np.array_equal(dbg["win_idx"][i, c, k].cpu().numpy(), r["win_idx"][k]):
                    print(name, "WINIDX MISMATCH", i, c, k); bad += 1
    err = np.abs(pts.cpu().numpy() - ref).max()
    print(name, "jobs", b * oc, "bad", bad, "max |pts - oracle| px", err, "stats", dbg["stats"].tolist(), "status", dbg["status"], flush=True)
    out[name] = {"bad": bad, "err": float(err), "stats": dbg["stats"].tolist()}


d = synthetic.make_frames(2, 120, 160, (1, 5, 6), variant="easy")
compare("small_easy", d, 64)
compare("small_easy_exact", d, 64, force_exact=True)
d = synthetic.make_frames(2, 120, 160, (1, 5, 6), variant="hard")
compare("small_hard", d, 64)
d = synthetic.make_frames(1, 240, 320, synthetic.CONFIG_8_IDS, variant="easy")
compare("mid_easy", d, 128)
compare("mid_cap", d, 128, max_num=500)
d = synthetic.make_frames(1, 480, 640, synthetic.CONFIG_8_IDS, variant="easy")
compare("full_easy_512", d, 512)

# rough timing, config 2
d = synthetic.make_frames(4, 480, 640, synthetic.CONFIG_8_IDS, variant="easy")
mask = torch.from_numpy(np.tile(d["mask"], (4, 1, 1, 1))).cuda()
vertex = torch.from_numpy(np.tile(d["vertex"], (4, 1, 1, 1, 1))).cuda()
print("sum tn per frame", d["mask"].sum((1, 2, 3)))
for it in range(3):
    ransac_voting_layer_all_masks(mask, vertex, 512, seed=it)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 10
for it in range(K):
    ransac_voting_layer_all_masks(mask, vertex, 512, seed=100 + it)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
units = 9 * float(d["mask"].sum()) * 4 * 513
print("config2 batch16: %.3f ms/step -> %.1f frames/s ; %.2f Gunits -> %.2f Tunits/s -> %.1f TFLOP/s(11/unit)" % (ms, 16 / ms * 1e3, units / 1e9, units / ms / 1e9, 11 * units / ms / 1e9))
out["config2_ms"] = ms
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_first.json", "w"), indent=1)
