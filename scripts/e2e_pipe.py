"""Host entry point (casa_ransac_vote_host) with k callers: k Python threads, each with its own handle, call the synchronous
entry point on the same pinned inputs (ctypes releases the GIL), so that the host packing of one call runs beside the GPU
tail of another.  Design exploration for the pipelined host entry.  usage: python scripts/e2e_pipe.py [steps]"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import synthetic  # noqa: E402
from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks_host  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
d = synthetic.make_frames(16, 480, 640, synthetic.CONFIG_8_IDS, seed=synthetic.SEED_BASE, variant="easy")
mask_h = torch.from_numpy(d["mask"]).pin_memory()
vertex_h = torch.from_numpy(d["vertex"]).pin_memory()


def run(callers, pack_threads, label=""):
    if pack_threads:
        os.environ["CASA_HOST_THREADS"] = str(pack_threads)
    else:
        os.environ.pop("CASA_HOST_THREADS", None)
    outs = [torch.empty((16, 8, 9, 2), dtype=torch.float32).pin_memory() for _ in range(callers)]
    start = threading.Barrier(callers + 1)
    done = threading.Barrier(callers + 1)

    def work(i):
        for it in range(2):
            ransac_voting_layer_all_masks_host(mask_h, vertex_h, 512, seed=it, out=outs[i])
        start.wait()
        for it in range(i, steps, callers):
            ransac_voting_layer_all_masks_host(mask_h, vertex_h, 512, seed=2000 + it, out=outs[i])
        done.wait()

    th = [threading.Thread(target=work, args=(i,)) for i in range(callers)]
    for t in th:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    done.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    print("%d caller(s), %s packer threads each %s: %.3f ms per 16 frames, %.0f frames/s" % (
        callers, pack_threads or "default", label, dt / steps * 1e3, 16 * steps / dt), flush=True)


print("host cores", os.cpu_count())
run(1, 0)
run(2, 0)
run(2, 6)
run(2, 8)
run(3, 6)
run(3, 4)
run(1, 12)
run(2, 12)
