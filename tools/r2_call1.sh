#!/bin/bash
# Round 2, call 1: at-size parity tests, large wgrad cases, CTA-pair (multicast) conv variant: validation + A/B.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
free -g > gpurun_out/r2c1_mem.txt; nproc >> gpurun_out/r2c1_mem.txt
timeout 1200 python -m pytest tests/test_gpu_parity_at_size.py -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/r2c1_parity.log 2>&1
echo "== at-size parity exit $?"; grep -E "^(forced|B=256|   worst)|passed|failed|Error|assert" gpurun_out/r2c1_parity.log | head -60
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "conv_wgrad" > gpurun_out/r2c1_wgrad.log 2>&1
echo "== wgrad tests exit $?"; tail -3 gpurun_out/r2c1_wgrad.log
GDL_FLAT_CLUSTER=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "conv_fwd or conv_dgrad" > gpurun_out/r2c1_cluster_tests.log 2>&1
echo "== cluster kernel tests exit $?"; tail -3 gpurun_out/r2c1_cluster_tests.log
for c in 0 1; do
  GDL_FLAT_CLUSTER=$c timeout 200 python tools/conv_bench.py --ops fwd,dgrad --out gpurun_out/r2c1_conv_bench.json --tag cluster$c > gpurun_out/r2c1_conv_cluster$c.log 2>&1
  echo "== conv_bench cluster=$c exit $?"; tail -1 gpurun_out/r2c1_conv_cluster$c.log
done
