"""`ncu -i X.ncu-rep --page raw --csv` -> one block of selected metrics per captured launch (the format of the
r02_ncu_*_summary.txt files).     python profiles/ncu_kernel_blocks.py raw.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
name_col = hdr.index("Kernel Name")
for r in data:
    print(r[name_col][:110])
    for m in WANT:
        if m in hdr:
            i = hdr.index(m)
            print(f"    {m:84s} {r[i]:>16s} [{units[i]}]")
    print()
