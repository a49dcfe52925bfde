#!/bin/bash
# Round 2, call 14: programmatic dependent launch for the training-step kernels: parity, determinism, A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_compat.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/r2c14_tests_a.log 2>&1
echo "== kernels + step + compat exit $?"; grep -E "passed|failed|^FAILED|Error|^E " gpurun_out/r2c14_tests_a.log | tail -8 | cut -c1-300
run() {  # label args-and-env
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-device-pipeline $EXTRA > gpurun_out/r2c14_bench_$label.log 2>&1
  echo "== bench $label exit $?"; grep '^{"metric"' gpurun_out/r2c14_bench_$label.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('   ms/step %.3f value %.0f e2e %.0f frac %.3f launches %d clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['launches_per_step'], d['clocks']['sm_mhz']))
"
}
run pdl1
run pdl0 GDL_PDL=0
run pdl1b
run pdl0b GDL_PDL=0
EXTRA="--dataset KineticSound --batch 64"
run ks64_pdl1
run ks64_pdl0 GDL_PDL=0
