#!/bin/bash
# full GPU regression: kernel tests, step/compat tests, bench with op dump
mkdir -p gpurun_out
TAG=${1:-full}
timeout 1200 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/kernel_tests_$TAG.log 2>&1
echo "== kernel tests exit $?"; tail -3 gpurun_out/kernel_tests_$TAG.log
bash tools/gpu_step_check.sh $TAG
