"""Collect the bench.py JSON lines of a tools/r2_dist.sh run into one strong-scaling table (profiles/r2_strong_scaling.json)."""
import glob
import json
import os
import re
import sys

out = {}
for f in sorted(glob.glob(os.path.join(sys.argv[1], "strong_*_n*.log"))):
    m = re.search(r"strong_(\w+)_G(\d+)_n(\d+)\.log", f)
    line = [ln for ln in open(f) if ln.startswith('{"metric"')]
    if not m or not line:
        continue
    d = json.loads(line[-1])
    key = "%s G=%s" % (m.group(1), m.group(2))
    out.setdefault(key, {})[int(m.group(3))] = {"samples_per_s": round(d["value"], 1), "ms_per_step": round(d["ms_per_step"], 3),
                                               "e2e_samples_per_s": round(d["e2e"]["value"], 1),
                                               "per_gpu_batch": d["config"]["global_batch"] // d["n_gpus"],
                                               "e2e_path": d["e2e"].get("path"),
                                               "e2e_host_frames_samples_per_s": round(d["e2e_host_frames"]["value"], 1)
                                               if "e2e_host_frames" in d else None,
                                               "clocks": d.get("clocks")}
for key, rows in out.items():
    if 1 in rows:
        for n, r in rows.items():
            r["speedup_vs_1gpu"] = round(r["samples_per_s"] / rows[1]["samples_per_s"], 3)
            r["efficiency"] = round(r["samples_per_s"] / rows[1]["samples_per_s"] / n, 4)
print(json.dumps({"what": "fixed global batch (strong scaling), bench.py --global-batch, device-timed, max over ranks",
                  "results": {k: {str(n): v for n, v in sorted(r.items())} for k, r in out.items()}}, indent=1))
