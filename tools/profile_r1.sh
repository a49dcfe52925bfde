#!/bin/bash
# ncu evidence for the round: launch list of one bench command + full captures of the conv kernels.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --batch ${PROF_BATCH:-256}"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 100 -c 4 -f -o gpurun_out/prof_igemm $BENCH > gpurun_out/prof_igemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 50 -c 3 -f -o gpurun_out/prof_wgrad $BENCH > gpurun_out/prof_wgrad.log 2>&1
ls -la gpurun_out/
