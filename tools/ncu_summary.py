"""Summarise an ncu --page raw --csv dump: python tools/ncu_summary.py raw.csv [pattern ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or [
    "gpu__time_duration.sum", "sm__pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "sm__warps_active.avg.pct", "launch__registers_per_thread",
    "sm__throughput.avg.pct", "gpu__dram_throughput", "lts__throughput.avg.pct",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__average_warp", "smsp__warp_issue_stalled", "sm__inst_executed_pipe_uniform",
    "l1tex__throughput", "local_load", "local_store", "smsp__cycles_active.avg",
    "launch__grid_size", "launch__occupancy", "shared_mem", "tensor", "tmem", "utc",
]
for r in rows[2:]:
    print("==", r[4][:60], r[8], r[7])
    for h, u, v in zip(hdr, units, r):
        if any(p in h for p in pats):
            print(f"  {h} [{u}] = {v}")
