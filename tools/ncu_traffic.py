"""DRAM traffic and tensor-pipe activity of the implicit-GEMM launches of the step, from an ncu metrics CSV
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active...`
with `--csv --log-file`):  python tools/ncu_traffic.py in.csv out.json [batch] > table.md
The JSON is what bench.py's roofline.traffic reads (profiles/r1_conv_traffic.json)."""
import csv
import json
import re
import sys
from collections import OrderedDict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "%": 1.0}
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
launch = OrderedDict()
for r in rows:
    d = launch.setdefault(int(r[0]), {"name": re.sub(r"^void ", "", re.sub(r"\(.*", "", r[4])), "grid": r[8]})
    d[r[12]] = float(r[14].replace(",", "")) * UNIT.get(r[13], 1.0)
L = list(launch.values())
steps = max(1, sum(1 for d in L if "stem_fwd_kernel" in d["name"]) // 2)  # 2 stems per step
TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
agg = OrderedDict()
for d in L:
    a = agg.setdefault(d["name"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "tp_ns": 0.0})
    a["n"] += 1
    a["ns"] += d.get("gpu__time_duration.sum", 0.0)
    a["rd"] += d.get("dram__bytes_read.sum", 0.0)
    a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    a["tp_ns"] += d.get(TP, 0.0) * d.get("gpu__time_duration.sum", 0.0)
n = len(L)
rd, wr = sum(a["rd"] for a in agg.values()), sum(a["wr"] for a in agg.values())
ns = sum(a["ns"] for a in agg.values())
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (+ duration, tensor pipe) over every "
                 "implicit-GEMM launch of %d training step(s), batch %s" % (steps, sys.argv[3] if len(sys.argv) > 3 else "256"),
       "batch": int(sys.argv[3]) if len(sys.argv) > 3 else 256,
       "steps": steps, "launches": n, "launches_per_step": n / steps,
       "dram_read_bytes_per_step": rd / steps, "dram_write_bytes_per_step": wr / steps,
       "dram_bytes_per_launch": (rd + wr) / n,
       "tensor_pipe_active_pct_time_weighted": sum(a["tp_ns"] for a in agg.values()) / ns,
       "ncu_ms_per_step": ns / steps / 1e6}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print("| kernel | launches/step | ms/step (ncu) | DRAM read MB/step | DRAM write MB/step | tensor pipe active % (time-weighted) |")
print("|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    print("| `%s` | %.1f | %.3f | %.0f | %.0f | %.1f |" % (k[:70], a["n"] / steps, a["ns"] / steps / 1e6, a["rd"] / steps / 1e6,
                                                  a["wr"] / steps / 1e6, a["tp_ns"] / max(a["ns"], 1.0)))
print("\n%d launches over %d step(s): %.1f MB DRAM traffic per launch on average (read %.2f GB + write %.2f GB per step); "
      "tensor pipe active %.1f %% of the (ncu-serialised) conv time" % (n, steps, (rd + wr) / n / 1e6, rd / steps / 1e9,
                                                                       wr / steps / 1e9, out["tensor_pipe_active_pct_time_weighted"]))
