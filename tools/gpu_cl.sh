#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "conv_fwd or conv_dgrad" > gpurun_out/kernel_tests_final.log 2>&1
echo "== default conv tests exit $?"; tail -2 gpurun_out/kernel_tests_final.log
GDL_FLAT_CLUSTER=1 timeout 25 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -x -k "test_conv_fwd and case14" > gpurun_out/kernel_tests_cluster.log 2>&1
echo "== cluster conv test exit $?"; tail -12 gpurun_out/kernel_tests_cluster.log | cut -c1-300
