#!/bin/bash
# final-state refresh of the ncu evidence: launch list + visual-encoder conv_flat + stem tail captures
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --batch ${PROF_BATCH:-256}"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_flat_kernel -s 21 -c 10 -f -o gpurun_out/prof_flat_v $BENCH > gpurun_out/prof_flat_v.log 2>&1
$NCU -k regex:"bn_relu_maxpool|bn_bwd_nores" -s 2 -c 8 -f -o gpurun_out/prof_tail2 $BENCH > gpurun_out/prof_tail2.log 2>&1
ls -la gpurun_out/*.ncu-rep
