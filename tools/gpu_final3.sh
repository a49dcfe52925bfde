#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider ) > gpurun_out/final_gpu_tests.log 2>&1; echo "== gpu tests exit $?"; tail -4 gpurun_out/final_gpu_tests.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1; echo "== smoke exit $?"; tail -4 gpurun_out/final_smoke.log
( time python bench.py ) > gpurun_out/final_bench.log 2>&1; echo "== bench exit $?"; tail -5 gpurun_out/final_bench.log | cut -c1-600
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/final_ref.log 2>&1; echo "== ref exit $?"; tail -5 gpurun_out/final_ref.log | cut -c1-400
