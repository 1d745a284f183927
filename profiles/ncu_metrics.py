"""Print selected metrics from an `ncu --page raw --csv` dump (one column per captured launch)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'lts__t_bytes.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'l1tex__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
extra = sys.argv[2:]
for i, h in enumerate(hdr):
    if h in WANT or any(e in h for e in extra):
        print(f"{h:72s} {units[i]:14s}", [r[i] for r in data])
