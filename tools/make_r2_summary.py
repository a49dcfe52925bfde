"""Assemble profiles/r2_ncu_summary.md from the ncu outputs of tools/profile_r2_final.sh (read here, no GPU):
    python tools/make_r2_summary.py        (expects gpurun_out/{launches_r2,conv_traffic_r2,hbm_kernels_r2}.csv, prof_final_r2.ncu-rep)"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
run = lambda *a: subprocess.run([sys.executable] + list(a), capture_output=True, text=True, cwd=ROOT).stdout

shutil.copy(os.path.join(G, "launches_r2.csv"), os.path.join(P, "r2_launches.csv"))
shutil.copy(os.path.join(G, "conv_traffic_r2.csv"), os.path.join(P, "r2_conv_traffic.csv"))
launch = run("tools/launch_summary.py", "profiles/r2_launches.csv")
traffic = run("tools/ncu_traffic.py", "profiles/r2_conv_traffic.csv", "profiles/r2_conv_traffic.json", "256")

lines = [l for l in open(os.path.join(G, "hbm_kernels_r2.csv")) if l.startswith('"')]
d = collections.OrderedDict()
for x in csv.DictReader(lines):
    k = (x["ID"], x["Kernel Name"].split("(")[0].replace("gdl::", "").replace("void ", ""))
    d.setdefault(k, {})[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
agg = collections.OrderedDict()
for (_, name), m in d.items():
    rd, wr, t = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0), m.get("gpu__time_duration.sum", 0)
    a = agg.setdefault(name, [0, 0.0, 0.0, []])
    a[0] += 1; a[1] += t; a[2] += rd + wr; a[3].append((t, rd + wr))
hbm = ["| kernel | launches | total ms (ncu) | DRAM GB read+write | GB/s over all launches | GB/s of the 3 largest launches |", "|---|---|---|---|---|---|"]
for name, (n, t, b, ls) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    big = sorted(ls, key=lambda x: -x[1])[:3]
    hbm.append("| `%s` | %d | %.3f | %.2f | %.0f | %s |" % (name[:60], n, t / 1e6, b / 1e9, b / t if t else 0,
                                                     ", ".join("%.0f" % (bb / tt) for tt, bb in big)))
fullp = os.path.join(P, "r2_full_audio_bwd.md")
if os.path.exists(os.path.join(G, "prof_final_r2.ncu-rep")):
    open(fullp, "w").write(run("tools/ncu_table.py", "gpurun_out/prof_final_r2.ncu-rep"))
full = open(fullp).read() if os.path.exists(fullp) else "(capture not present)"
visp = os.path.join(P, "r2_full_visual_bwd.md")
if os.path.exists(os.path.join(G, "prof_flat2_visual_r2.ncu-rep")):
    open(visp, "w").write(run("tools/ncu_table.py", "gpurun_out/prof_flat2_visual_r2.ncu-rep"))
vis = open(visp).read() if os.path.exists(visp) else "(capture not present)"

open(os.path.join(P, "r2_ncu_summary.md"), "w").write("""# Round 2 — ncu evidence (B200, batch 256 CREMA-D shape, eager step; `tools/profile_r2_final.sh` under gpurun)

All commands wrap `python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --no-device-pipeline --batch 256`.
Per-launch times under ncu are cold-cache and serialised: shares and per-kernel rates, not step time.  Measured peaks
(`MEASURED_PEAKS.json`): 6458 GB/s copy bandwidth, 1378.9 TF sustained bf16.

## 1. Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -s 1150 -c 900`; raw: `profiles/r2_launches.csv`)

900 consecutive launches = 2.4 training steps of 373 launches.  Against round 1 (`r1_ncu_summary.md`): no `bn_stats_kernel`
(the statistics come out of the conv / stem epilogues), `conv_flat2_kernel` (CTA pair, `UTCHMMA.2CTA`) in place of
`conv_flat_kernel` everywhere, `bn_relu_maxpool_fwd3_kernel` 312 us (fwd2: 507 us), `stem_wgrad_kernel` 267 us (393 us),
`stem_layout_kernel` 120 us (193 us).

%s
## 2. DRAM traffic and tensor-pipe activity of every implicit-GEMM launch of two steps
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:conv_flat…`;
raw: `profiles/r2_conv_traffic.csv`, summary `profiles/r2_conv_traffic.json` = `roofline.traffic` of bench.py)

Template arguments of `conv_flat2_kernel<BN, MT, WST, STATS, RES>`: STATS = 1 are the forward convolutions (BatchNorm
statistics in the epilogue), RES = 1 the 64 -> 64 channel layers with resident weights.

%s
## 3. HBM-bound kernels (`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput…`, 300 launches)

GB/s = DRAM bytes / duration per launch, cold cache.  The small layers (tens of microseconds) pull the all-launch
average down; the three largest launches of each kernel show what the kernel reaches when the tensor is large.

%s

## 4. `--set full` captures (24 launches of the audio encoder's backward; `gpurun_out/prof_final_r2.ncu-rep`, scratch)

%s
## 5. `--set full` capture of 20 CTA-pair convolution launches at the end of the audio and the start of the VISUAL backward
(`-k regex:conv_flat2_kernel -s 196 -c 20`).  The 137-158 us launches are the data gradients of the visual layer4 / layer3
(7x7 C512, 14x14 C256): **tensor pipe 86-92 %% active**; `<64, 2, 8, 0, 1>` are the 64-channel layers with resident weights
(audio 65x47 maps here: 39-51 %%, shared-memory-bandwidth bound).

%s
""" % (launch, traffic, "\n".join(hbm), full, vis))
print("wrote profiles/r2_ncu_summary.md")
