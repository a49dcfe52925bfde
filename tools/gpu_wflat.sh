#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; ( "$@" ) > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -4 gpurun_out/$name.log; }
K="test_conv_wgrad"
run tw_auto  env timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" --no-header -p no:cacheprovider
run tw_all   env GDL_WFLAT=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" --no-header -p no:cacheprovider
run bw_auto  env timeout 600 python tools/conv_bench.py --tag wauto --ops wgrad --out gpurun_out/wgrad_bench.json
run bw_all   env GDL_WFLAT=2 timeout 600 python tools/conv_bench.py --tag wall --ops wgrad --out gpurun_out/wgrad_bench.json
