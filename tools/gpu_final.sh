#!/bin/bash
# what the driver runs at round end: smoke, default bench, reference arm
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1; echo "== smoke exit $?"; tail -4 gpurun_out/final_smoke.log
( time python bench.py ) > gpurun_out/final_bench.log 2>&1; echo "== bench exit $?"; tail -5 gpurun_out/final_bench.log | cut -c1-400
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/final_ref.log 2>&1; echo "== ref exit $?"; tail -5 gpurun_out/final_ref.log | cut -c1-600
( time python bench.py --dataset KineticSound --batch 64 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/final_ks.log 2>&1; echo "== ks exit $?"; tail -5 gpurun_out/final_ks.log | cut -c1-300
