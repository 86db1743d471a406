# host-side phase trace of the pipelined host entry (CASA_HOST_TRACE): last 3 calls
CASA_HOST_TRACE=1 python - <<'P' 2>&1 | tail -70
import sys, torch
sys.path.insert(0, ".")
from casapose_b200 import synthetic, _lib
from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks_host as host
d = synthetic.make_frames(16, 480, 640, synthetic.CONFIG_8_IDS, seed=synthetic.SEED_BASE, variant="easy")
m = torch.from_numpy(d["mask"]).pin_memory(); v = torch.from_numpy(d["vertex"]).pin_memory()
outs = [torch.empty((16, 8, 9, 2)).pin_memory() for _ in range(3)]
from collections import deque
import os; DEPTH = int(os.environ.get("CASA_HOST_DEPTH", "2")); pend = deque()
for it in range(16):
    if len(pend) == DEPTH: pend.popleft().result()
    pend.append(host(m, v, 512, seed=it, out=outs[it % 3], wait=False))
while pend: pend.popleft().result()
P
