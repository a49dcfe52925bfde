#!/bin/bash
# opt-in resident-weights conv_flat (C64 layers): parity through the step tests, then A/B
mkdir -p gpurun_out
export GDL_FLAT_RESIDENT=1
timeout 100 python -m pytest tests/test_gpu_step.py -q -m gpu --no-header -p no:cacheprovider -x -k "not two_gpu and not smoke" > gpurun_out/step_tests_res.log 2>&1
echo "== step tests (resident) exit $?"; tail -3 gpurun_out/step_tests_res.log
for r in 1 0; do
  GDL_FLAT_RESIDENT=$r GDL_DUMP_OPS=gpurun_out/ops_res$r.json timeout 60 python bench.py --steps 12 --warmup 4 --no-cpu --no-device-pipeline > gpurun_out/bench_res$r.log 2>&1
  echo "== bench resident=$r exit $?"; grep '^{"metric"' gpurun_out/bench_res$r.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f frac %.3f fwd %.2f dgrad %.2f' % (d['ms_per_step'], d['value'], d['roofline']['frac'], kb['conv_fwd']['ms'], kb['conv_dgrad']['ms']))
"
done
