#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; ( "$@" ) > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -3 gpurun_out/$name.log; }
K="test_conv_fwd or test_conv_dgrad or test_conv_wgrad"
run t_def   env timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" --no-header -p no:cacheprovider
run b_s2    env timeout 600 python tools/conv_bench.py --tag s2 --only s2 --ops fwd --out gpurun_out/conv_bench3.json
grep fwd gpurun_out/b_s2.log
