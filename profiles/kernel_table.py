"""Per-kernel table from `ncu --metrics ... --csv` over tools/one_forward.py (second, warm forward only).
    python profiles/kernel_table.py metrics.csv > profiles/r01_kernel_table.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, ni, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    per.setdefault(int(r[ii]), {"name": r[ki]})[r[ni]] = float(r[vi].replace(",", "") or 0)
ids = sorted(per)
firsts = [i for i in ids if "dedup_insert" in per[i]["name"]]
start = firsts[1] if len(firsts) > 1 else 0
agg = collections.OrderedDict()
for i in ids:
    if i < start:
        continue
    k = per[i]
    name = k["name"].split("(")[0].replace("ghf::<unnamed>::", "").replace("void ", "")[:58]
    a = agg.setdefault(name, collections.Counter())
    a["n"] += 1
    a["ns"] += k.get("gpu__time_duration.sum", 0)
    a["rd"] += k.get("dram__bytes_read.sum", 0)
    a["wr"] += k.get("dram__bytes_write.sum", 0)
    a["tc"] += k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) * k.get("gpu__time_duration.sum", 0)
tot = sum(a["ns"] for a in agg.values())
print(f"{'kernel':58s} {'n':>3s} {'us':>9s} {'share':>6s} {'DRAM GB':>8s} {'GB/s':>7s} {'tensor%':>7s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    gb = (a["rd"] + a["wr"]) / 1e9
    print(f"{name:58s} {a['n']:3d} {a['ns'] / 1e3:9.1f} {100 * a['ns'] / tot:5.1f}% {gb:8.3f} "
          f"{(a['rd'] + a['wr']) / max(a['ns'], 1):7.0f} {a['tc'] / max(a['ns'], 1):7.1f}")
print(f"{'total':58s} {sum(a['n'] for a in agg.values()):3d} {tot / 1e3:9.1f}")
