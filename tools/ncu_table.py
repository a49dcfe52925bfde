"""Compact per-launch table from an .ncu-rep (read offline): python tools/ncu_table.py file.ncu-rep > table.md"""
import csv
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
print("| kernel | grid | block | " + " | ".join(c[1] for c in COLS) + " |")
print("|---|---|---|" + "---|" * len(COLS))
for r in data:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("gdl::", "")
    vals = []
    for key, lab in COLS:
        i = idx.get(key)
        v = r[i] if i is not None else ""
        u = units[i] if i is not None else ""
        try:
            f = float(v.replace(",", ""))
            if lab in ("rdMB", "wrMB"):
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            if lab == "us":
                f *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
            v = "%.1f" % f if f < 1000 else "%.0f" % f
        except ValueError:
            pass
        vals.append(v)
    print("| %s | %s | %s | %s |" % (name[:48], r[idx["Grid Size"]], r[idx["Block Size"]], " | ".join(vals)))
