"""Instruction-mix exploration for the scoring loop (casa_measure_fp32_peak variants)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import _lib  # noqa: E402

L = _lib.lib()
h = _lib.handle(0)
names = {0: "FFMA", 3: "shipped loop", 40: "packed FFMA2 loop", 50: "mma.sync m16n8k8 tf32 alone (TFLOP/s)",
         51: "tensor-path loop: count + min", 52: "tensor-path loop: count", 53: "tensor-path loop: no count/min",
         54: "hybrid loop: s on the tensor path, p as FFMA"}
if len(sys.argv) > 1:
    names = {int(v): names.get(int(v), "variant %s" % v) for v in sys.argv[1:]}
mins = ["no min", "FMNMX3/pair", "FMNMX/unit"]
cnts = ["no count", "LEA.HI", "IMAD.HI", "LEA/IMAD alternate"]
if len(sys.argv) <= 1:
    for m in range(3):
        for c in range(4):
            names[20 + m * 4 + c] = "loop: %s, %s" % (mins[m], cnts[c])
for v in sorted(names):
    tf, ms = C.c_double(), C.c_double()
    _lib.check(L.casa_measure_fp32_peak(h, v, C.byref(tf), C.byref(ms)))
    print("%-40s %7.2f TFLOP/s-equiv  %.3f ms" % (names[v], tf.value, ms.value), flush=True)
