"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): one line per kernel launch and the total.
usage: python scripts/launch_list.py gpurun_out/launches.csv"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = 0.0
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v = v / 1000.0 if r[iu] in ("ns", "nsecond") else v
    tot += v
    print("%-70s %9.2f us" % (r[ik][:70], v))
print("%-70s %9.2f us" % ("total", tot))
