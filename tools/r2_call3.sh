#!/bin/bash
# Round 2, call 3: CTA-pair (cta_group::2) conv kernel: parity + A/B against the single-CTA kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -x -k "conv_fwd or conv_dgrad" > gpurun_out/r2c3_pair_tests.log 2>&1
echo "== pair kernel tests exit $?"; tail -25 gpurun_out/r2c3_pair_tests.log | cut -c1-300
for c in 0 1; do
  GDL_FLAT_PAIR=$c timeout 200 python tools/conv_bench.py --ops fwd,dgrad --out gpurun_out/r2c3_conv_bench.json --tag pair$c > gpurun_out/r2c3_conv_pair$c.log 2>&1
  echo "== conv_bench pair=$c exit $?"; tail -1 gpurun_out/r2c3_conv_pair$c.log
done
timeout 900 python -m pytest tests/test_gpu_check_mode.py tests/test_gpu_parity_at_size.py -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/r2c3_parity.log 2>&1
echo "== parity + check-mode exit $?"; grep -E "passed|failed" gpurun_out/r2c3_parity.log; grep -E "B=256|bf16 emulation" gpurun_out/r2c3_parity.log | cut -c1-500
