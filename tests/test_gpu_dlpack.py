"""The DLPack seam of the C ABI (casa_ransac_vote_dlpack) fed by a hand-rolled ``__dlpack__`` exporter that is not a
torch tensor, in a subprocess that never imports torch: results equal the reference code's golden vector, every
capsule's deleter runs exactly once, malformed tensors are rejected in C."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

SCRIPT = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
from dlpack_exporter import CudaBuffer, cudart
from casapose_b200 import _lib, dlpack_api

g = np.load(os.path.join(%(root)r, "tests", "golden", "ransac_easy.npz"))
oc = g["points"].shape[1]
mask = np.stack([g["labels"] == c + 1 for c in range(oc)], -1).astype(np.float32)
vertex = g["vertex"].astype(np.float32)
b = mask.shape[0]
m, v, o = CudaBuffer(mask), CudaBuffer(vertex), CudaBuffer(shape=(b, oc, vertex.shape[3], 2))
dlpack_api.ransac_voting_layer_all_masks_dlpack(m, v, o, int(g["hn"]), seed=int(g["seed"]))
_lib.sync(0)
err = float(np.abs(o.numpy() - g["points"]).max())
assert err < 1e-3, err
assert (m.deleter_calls, v.deleter_calls, o.deleter_calls) == (1, 1, 1), (m.deleter_calls, v.deleter_calls, o.deleter_calls)

# rejected in C: wrong dtype, non-contiguous strides, wrong out shape — and the capsules are still released once
for bad_mask, bad_out in ((CudaBuffer(mask, dtype_code=0), None),
                          (CudaBuffer(mask, strides=(mask[0].size, mask.shape[2] * oc, 1, mask.shape[2])), None),
                          (None, CudaBuffer(shape=(b, oc, 3, 2)))):
    mm = bad_mask or CudaBuffer(mask)
    oo = bad_out or CudaBuffer(shape=(b, oc, vertex.shape[3], 2))
    vv = CudaBuffer(vertex)
    try:
        dlpack_api.ransac_voting_layer_all_masks_dlpack(mm, vv, oo, int(g["hn"]), seed=1)
        raise SystemExit("a malformed tensor was accepted")
    except _lib.CasaError as e:
        assert "error -1" in str(e), str(e)
    assert (mm.deleter_calls, vv.deleter_calls, oo.deleter_calls) == (1, 1, 1)
assert "torch" not in sys.modules, "the DLPack path must not need torch"
print("dlpack ok", err)
'''


def test_vote_through_a_non_torch_dlpack_exporter(cuda_lib):
    res = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "dlpack ok" in res.stdout
