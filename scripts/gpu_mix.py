"""Instruction-mix exploration for the scoring loop (casa_measure_fp32_peak variants)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import _lib  # noqa: E402

L = _lib.lib()
h = _lib.handle(0)
names = {0: "FFMA", 1: "FFMA2", 2: "FFMA 3-reg", 3: "shipped loop"}
for k in (1, 2, 3, 4, 5, 6, 8):
    names[3 + 100 * k] = "shipped loop, %d blocks (%d warps)/SM" % (k, 8 * k)
for v in sorted(names):
    tf, ms = C.c_double(), C.c_double()
    _lib.check(L.casa_measure_fp32_peak(h, v, C.byref(tf), C.byref(ms)))
    print("%-40s %7.2f TFLOP/s-equiv  %.3f ms" % (names[v], tf.value, ms.value), flush=True)
