#!/bin/bash
# Round 2, call 2: FP32 check mode tests + the margin-aware B=256 arg-max criterion.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_check_mode.py -q -m gpu --no-header -p no:cacheprovider -s --durations=10 > gpurun_out/r2c2_check.log 2>&1
echo "== check mode exit $?"; grep -E "check-mode|passed|failed|Error|assert " gpurun_out/r2c2_check.log | cut -c1-600 | head -40
timeout 600 python -m pytest tests/test_gpu_parity_at_size.py -q -m gpu --no-header -p no:cacheprovider -s -k bench_size > gpurun_out/r2c2_parity.log 2>&1
echo "== at-size parity exit $?"; grep -E "B=256|worst per|passed|failed|Error|assert " gpurun_out/r2c2_parity.log | cut -c1-600 | head -40
