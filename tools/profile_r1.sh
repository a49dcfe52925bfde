#!/bin/bash
# ncu evidence for the round (run under gpurun, 1 GPU): launch list of one bench command + full captures
# of the dominant kernels.  Summaries are written by tools/ncu_summary.py into profiles/.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --batch ${PROF_BATCH:-256}"
# 4 eager steps = ~1900 kernels of ours + torch glue; skip the first 3 steps' worth, keep one full step
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo_kernel -s 60 -c 6 -f -o gpurun_out/prof_halo_fwd $BENCH > gpurun_out/prof_halo_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad_halo -s 30 -c 5 -f -o gpurun_out/prof_halo_wgrad $BENCH > gpurun_out/prof_halo_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_reduce|bn_bwd_apply|bn_apply_kernel|bn_stats_kernel" -s 120 -c 8 -f -o gpurun_out/prof_bn $BENCH > gpurun_out/prof_bn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stem_fwd|stem_wgrad_kernel|conv_igemm" -s 8 -c 6 -f -o gpurun_out/prof_misc $BENCH > gpurun_out/prof_misc.log 2>&1
ls -la gpurun_out/*.ncu-rep
