#!/bin/bash
# Round 2, call 16: CTA-pair weight-gradient kernel: parity, A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "conv_wgrad or gemm_tn or film" > gpurun_out/r2c16_tests_a.log 2>&1
echo "== wgrad kernel tests exit $?"; grep -E "passed|failed|^FAILED|Error|^E |timeout|trap" gpurun_out/r2c16_tests_a.log | tail -12 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity_at_size.py -q -m gpu --no-header -p no:cacheprovider -k "not bench_size" > gpurun_out/r2c16_tests_b.log 2>&1
echo "== step + forced-forward exit $?"; grep -E "passed|failed|^FAILED" gpurun_out/r2c16_tests_b.log | tail -8 | cut -c1-300
run() {  # label env...
  local label=$1; shift
  env "$@" GDL_DUMP_OPS=gpurun_out/r2c16_ops_$label.json timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-device-pipeline > gpurun_out/r2c16_bench_$label.log 2>&1
  echo "== bench $label exit $?"; grep '^{"metric"' gpurun_out/r2c16_bench_$label.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f e2e %.0f frac %.3f clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz']))
print('   ' + ' '.join('%s=%.2f' % (k, v['ms']) for k, v in kb.items() if k.startswith('conv')))
"
}
run pair1
run pair0 GDL_WGRAD_PAIR=0
run pair1b
run pair0b GDL_WGRAD_PAIR=0
