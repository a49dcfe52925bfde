#!/bin/bash
# conv_flat bring-up: parity under the three dispatch policies, then per-layer timings per policy.
mkdir -p gpurun_out
run() { # name cmd...
  name=$1; shift
  ( "$@" ) > gpurun_out/$name.log 2>&1
  echo "== $name exit $?"; tail -4 gpurun_out/$name.log
}
K="test_conv_fwd or test_conv_dgrad"
run t_auto   env timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" --no-header -p no:cacheprovider
run t_flat2  env GDL_FLAT=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" --no-header -p no:cacheprovider
run t_flat2m env GDL_FLAT=2 GDL_FLAT_MT=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$K" --no-header -p no:cacheprovider
rm -f gpurun_out/conv_bench.json
run b_flat0  env GDL_FLAT=0 timeout 600 python tools/conv_bench.py --tag flat0
run b_auto   env timeout 600 python tools/conv_bench.py --tag auto --ops fwd,dgrad
run b_mt1    env GDL_FLAT=2 GDL_FLAT_MT=1 timeout 600 python tools/conv_bench.py --tag flat2mt1 --ops fwd,dgrad
run b_mt2    env GDL_FLAT=2 GDL_FLAT_MT=2 timeout 600 python tools/conv_bench.py --tag flat2mt2 --ops fwd,dgrad
