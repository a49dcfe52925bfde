#!/bin/bash
# Round 2, call 7: BatchNorm statistics from the conv / stem epilogues (shared-memory transpose): parity, then A/B on the bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -x -k "fused_bn_statistics or stem_space or conv_fwd or conv_dgrad" > gpurun_out/r2c7_tests_a.log 2>&1
echo "== kernels exit $?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/r2c7_tests_a.log | tail -8 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_compat.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2c7_tests_b.log 2>&1
echo "== step + compat exit $?"; grep -E "passed|failed|^FAILED" gpurun_out/r2c7_tests_b.log | tail -8 | cut -c1-300
run() {  # label env...
  local label=$1; shift
  env "$@" GDL_DUMP_OPS=gpurun_out/r2c7_ops_$label.json timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-device-pipeline > gpurun_out/r2c7_bench_$label.log 2>&1
  echo "== bench $label exit $?"; grep '^{"metric"' gpurun_out/r2c7_bench_$label.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f e2e %.0f frac %.3f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['launches_per_step']))
print('   ' + ' '.join('%s=%.2f' % (k, v['ms']) for k, v in kb.items()))
"
}
run old GDL_FUSED_STATS_MIN_K=100000000 GDL_STEM_STATS=0
run all GDL_FUSED_STATS_MIN_K=0 GDL_STEM_STATS=1
run k576 GDL_FUSED_STATS_MIN_K=576 GDL_STEM_STATS=1
run k1152 GDL_FUSED_STATS_MIN_K=1152 GDL_STEM_STATS=1
run old2 GDL_FUSED_STATS_MIN_K=100000000 GDL_STEM_STATS=0
