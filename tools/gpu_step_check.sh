#!/bin/bash
# Step-level regression + bench after a kernel change (1 GPU): step/compat parity tests, then bench with op dump.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_compat.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/step_tests_$TAG.log 2>&1
echo "== step tests exit $?"; tail -5 gpurun_out/step_tests_$TAG.log
GDL_DUMP_OPS=gpurun_out/ops_$TAG.json timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.log 2>&1
echo "== bench exit $?"; tail -2 gpurun_out/bench_$TAG.log | cut -c1-600
