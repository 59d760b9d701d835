"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the
last `--last N` launches (default: everything after the last fixed_base_mul launch)."""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]
start = 0
for i, (k, _) in enumerate(seq):
    if "fixed_base" in k or "scalars_generate" in k:
        start = i + 1
seq = seq[start:]
# split into MSMs (each starts at a digits kernel) and keep the largest one (the timed 2^20 workload;
# the bench also runs a 1024-point parity MSM at the end)
starts = [i for i, (k, _) in enumerate(seq) if "digits" in k] + [len(seq)]
groups = [seq[a:b] for a, b in zip(starts[:-1], starts[1:])]
# a group ends at its final kernel
def cut(g):
    for j, (k, _) in enumerate(g):
        if "msm_final" in k:
            return g[:j + 1]
    return g
groups = [cut(g) for g in groups]
seq = max(groups, key=lambda g: sum(v for k, v in g if "accumulate" in k))
tot = sum(v for _, v in seq)
agg = OrderedDict()
for k, v in seq:
    name = k.split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
print("%-60s %5s %12s %7s" % ("kernel", "n", "total_us", "share"))
for k, (c, v) in agg.items():
    print("%-60s %5d %12.1f %6.1f%%" % (k[:60], c, v / 1e3, 100 * v / tot))
print("%-60s %5d %12.1f" % ("TOTAL (one MSM)", len(seq), tot / 1e3))
