"""Print kernel name, grid and duration (us) from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
tot = 0.0
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("qq::", "")
    us = float(r[vi].replace(",", "")) / 1000.0
    tot += us
    print("%9.1f us %16s %s" % (us, r[gi], name))
print("%9.1f us total" % tot)
