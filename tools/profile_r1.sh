#!/bin/bash
# ncu evidence for the round (run under gpurun, 1 GPU): launch list of one bench command + full captures
# of the dominant kernels.  Summaries are written by tools/launch_summary.py / tools/ncu_table.py into profiles/.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --batch ${PROF_BATCH:-256}"
# 4 warm-up + 1 timed + e2e eager steps; skip the first 3 steps' worth of launches, keep > one full step
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_flat_kernel -s 40 -c 8 -f -o gpurun_out/prof_flat $BENCH > gpurun_out/prof_flat.log 2>&1
$NCU -k regex:conv_wgrad_flat -s 20 -c 6 -f -o gpurun_out/prof_wflat $BENCH > gpurun_out/prof_wflat.log 2>&1
$NCU -k regex:"bn_bwd_reduce|bn_bwd_apply|bn_bwd_nores|bn_apply_kernel|bn_stats_kernel|bn_relu_maxpool" -s 60 -c 12 -f -o gpurun_out/prof_bn $BENCH > gpurun_out/prof_bn.log 2>&1
$NCU -k regex:"stem_fwd|stem_wgrad_kernel|sgd_momentum|grad_stats_kernel|dgl_head" -s 4 -c 8 -f -o gpurun_out/prof_misc $BENCH > gpurun_out/prof_misc.log 2>&1
ls -la gpurun_out/*.ncu-rep
