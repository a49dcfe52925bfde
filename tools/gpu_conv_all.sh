#!/bin/bash
# conv parity (fwd/dgrad/wgrad) under the default policy + the "always flat" wgrad policy, then per-layer timings.
mkdir -p gpurun_out
run() { name=$1; shift; ( "$@" ) > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -3 gpurun_out/$name.log; }
K="test_conv_fwd or test_conv_dgrad or test_conv_wgrad"
run t_def   env timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" --no-header -p no:cacheprovider
run t_wall  env GDL_WFLAT=2 GDL_FLAT_MT=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" --no-header -p no:cacheprovider
rm -f gpurun_out/conv_bench2.json
run b_def   env timeout 600 python tools/conv_bench.py --tag def --out gpurun_out/conv_bench2.json
run b_wall  env GDL_WFLAT=2 timeout 600 python tools/conv_bench.py --tag wall --ops wgrad --out gpurun_out/conv_bench2.json
