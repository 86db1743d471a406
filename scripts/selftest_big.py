"""casa_selftest_filter at 2^32 adversarial units per (threshold, spread): the filtered predicate's verdict, wherever
the chunk-level or the unit-level bound declares it certain, against exact_inlier() on the device.
usage (on a B200): python scripts/selftest_big.py [log2 units, default 32]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import _lib  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = _lib.lib()
h = _lib.handle(0)
print("thr      spread    tested        mismatches  uncertain     exact inliers   seconds")
bad = 0
for thr in (0.5, 0.9, 0.97, 0.99, 0.999, 0.9999):
    for spread in (2e-6, 2e-5, 1e-3, 0.3, -2e-5):
        res = (C.c_uint64 * 4)()
        t0 = time.time()
        _lib.check(L.casa_selftest_filter(h, 1 << lg, 20261017, thr, spread, res))
        print("%-8g %-9g %-13d %-11d %-13d %-15d %.1f" % (thr, spread, res[0], res[1], res[2], res[3], time.time() - t0), flush=True)
        bad += res[1]
print("total mismatches:", bad)
sys.exit(1 if bad else 0)
