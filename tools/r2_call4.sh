#!/bin/bash
# Round 2, call 4: CTA-pair kernel after the MEMBAR fix (A/B), tests, first full bench of the round.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -x -k "conv_fwd or conv_dgrad" > gpurun_out/r2c4_pair_tests.log 2>&1
echo "== pair kernel tests exit $?"; tail -3 gpurun_out/r2c4_pair_tests.log | cut -c1-300
for c in 0 1; do
  GDL_FLAT_PAIR=$c timeout 200 python tools/conv_bench.py --ops fwd,dgrad --out gpurun_out/r2c4_conv_bench.json --tag pair$c > gpurun_out/r2c4_conv_pair$c.log 2>&1
  echo "== conv_bench pair=$c exit $?"; tail -1 gpurun_out/r2c4_conv_pair$c.log
done
for c in 0 1; do
  GDL_FLAT_PAIR=$c GDL_DUMP_OPS=gpurun_out/r2c4_ops_pair$c.json timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu --no-device-pipeline > gpurun_out/r2c4_bench_pair$c.log 2>&1
  echo "== bench pair=$c exit $?"; grep '^{"metric"' gpurun_out/r2c4_bench_pair$c.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f frac %.3f fwd %.2f dgrad %.2f wgrad %.2f' % (d['ms_per_step'], d['value'], d['roofline']['frac'], kb['conv_fwd']['ms'], kb['conv_dgrad']['ms'], kb['conv_wgrad']['ms']))
"
done
timeout 900 python -m pytest tests/test_gpu_check_mode.py tests/test_gpu_compat.py tests/test_gpu_step.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2c4_tests.log 2>&1
echo "== check-mode + compat + step tests exit $?"; tail -15 gpurun_out/r2c4_tests.log | cut -c1-300
