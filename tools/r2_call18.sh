#!/bin/bash
# Round 2, call 18: direction of the BatchNorm apply pass after a convolution with fused statistics (GDL_APPLY_SWEEP): A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2c18_tests.log 2>&1
echo "== step tests exit $?"; grep -E "passed|failed|^FAILED" gpurun_out/r2c18_tests.log | tail -3
run() {  # label env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-device-pipeline > gpurun_out/r2c18_bench_$label.log 2>&1
  echo "== bench $label exit $?"; grep '^{"metric"' gpurun_out/r2c18_bench_$label.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f e2e %.0f frac %.3f clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz']))
print('   ' + ' '.join('%s=%.2f' % (k, v['ms']) for k, v in kb.items() if k.startswith('conv') or k.startswith('bn_')))
"
}
run as1
run as0 GDL_APPLY_SWEEP=0
run as1b
run as0b GDL_APPLY_SWEEP=0
