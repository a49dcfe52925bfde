#!/bin/bash
# Last call of the round: the whole GPU suite, smoke(), the default bench and the reference arm as the driver runs them,
# then the ncu launch list of the final library.
bash tools/r2_full.sh
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2full_reference.log 2>&1
echo "== reference arm exit $?"; tail -1 gpurun_out/r2full_reference.log | cut -c1-300
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --no-device-pipeline --batch 256"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1150 -c 900 --csv --log-file gpurun_out/launches_r2.csv $BENCH > gpurun_out/launches_bench_r2.log 2>&1
echo "== launch list exit $?"
