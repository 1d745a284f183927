"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python profiles/launch_table.py launches.csv [first_id]     (only launches with ID >= first_id)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
agg, tot = collections.OrderedDict(), 0.0
for r in data:
    if len(r) <= vi or int(r[ii]) < first:
        continue
    v = float(r[vi].replace(",", ""))
    a = agg.setdefault(r[ki][:90], [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1e3:10.1f} us {n:4d}x {100 * v / tot:5.1f}%  {k}")
print(f"{tot / 1e3:10.1f} us total ({sum(n for n, _ in agg.values())} launches)")
