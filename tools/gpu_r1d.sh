#!/bin/bash
# stem-tail forward v2 + balanced stride-2 wgrad splits: tests and A/B on one GPU
mkdir -p gpurun_out
TAG=${1:-d1}
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/kernel_tests_$TAG.log 2>&1
echo "== kernel tests exit $?"; tail -3 gpurun_out/kernel_tests_$TAG.log
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_compat.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/step_tests_$TAG.log 2>&1
echo "== step tests exit $?"; tail -3 gpurun_out/step_tests_$TAG.log
run() {  # name, env...
  local name=$1; shift
  env "$@" GDL_DUMP_OPS=gpurun_out/ops_${TAG}_$name.json timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu > gpurun_out/bench_${TAG}_$name.log 2>&1
  echo "== $name exit $?"; tail -1 gpurun_out/bench_${TAG}_$name.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    kb = d['kernel_breakdown']
    print('   ms/step %.3f  value %.0f  e2e %.0f  conv frac %.3f  wgrad %.2f fwd %.2f dgrad %.2f bn_bwd %.2f stats %.2f apply %.2f tailf %.2f tailb %.2f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], kb['conv_wgrad']['ms'], kb['conv_fwd']['ms'], kb['conv_dgrad']['ms'], kb['bn_bwd']['ms'], kb['bn_stats']['ms'], kb['bn_apply']['ms'], kb['stem_tail_fwd']['ms'], kb['stem_tail_bwd']['ms']))
except Exception as e:
    print('   parse failed', e)
"
}
run default A=1
run tail1 GDL_STEM_TAIL=1
run default2 A=1
