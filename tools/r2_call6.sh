#!/bin/bash
# Round 2, call 6: BN-fold / STFT / fused gated head / check-mode thresholds; bench with fusion variants.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compat.py tests/test_gpu_datapipe.py tests/test_gpu_check_mode.py -q -m gpu --no-header -p no:cacheprovider -s > gpurun_out/r2c6_tests_a.log 2>&1
echo "== compat + datapipe + check-mode exit $?"; grep -E "passed|failed|^FAILED|eval logits" gpurun_out/r2c6_tests_a.log | tail -12 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2c6_tests_b.log 2>&1
echo "== kernels + step exit $?"; grep -E "passed|failed|^FAILED" gpurun_out/r2c6_tests_b.log | tail -12 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity_at_size.py -q -m gpu --no-header -p no:cacheprovider -k "concat-2 or gated" > gpurun_out/r2c6_tests_c.log 2>&1
echo "== at-size (concat-2, gated) exit $?"; grep -E "passed|failed|^FAILED" gpurun_out/r2c6_tests_c.log | tail -5 | cut -c1-300
for f in gated sum film; do
  timeout 300 python bench.py --fusion $f --steps 10 --warmup 3 --no-cpu --no-device-pipeline > gpurun_out/r2c6_bench_$f.log 2>&1
  echo "== bench $f exit $?"; grep '^{"metric"' gpurun_out/r2c6_bench_$f.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('   ms/step %.3f value %.0f e2e %.0f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['launches_per_step']))
"
done
timeout 600 python bench.py --impl cudnn_sidebar --steps 5 --warmup 3 > gpurun_out/r2c6_sidebar.log 2>&1
echo "== cudnn sidebar exit $?"; tail -2 gpurun_out/r2c6_sidebar.log | cut -c1-600
