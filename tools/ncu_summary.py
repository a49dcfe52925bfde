"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py file.ncu-rep [substr ...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active", "pipe_tensor", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "l1tex__throughput.avg.pct",
        "lts__throughput.avg.pct", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_lsu", "smsp__average_warp", "smsp__warp_issue_stalled", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "achieved_occupancy", "sm__ctas_launched"]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
print("kernels:", [(r[hdr.index("Kernel Name")][:30], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]) for r in data])
for i, h in enumerate(hdr):
    if any(k in h for k in KEYS + extra):
        print("%-90s %-14s %s" % (h[-90:], units[i], [r[i] for r in data]))
