#!/bin/bash
# Round 2, call 15: stream priorities of the two encoder streams (A/B inside one call).
mkdir -p gpurun_out
for rep in 1 2; do
for pr in 0 1 2; do
  GDL_STREAM_PRIO=$pr timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-device-pipeline --no-roofline > gpurun_out/r2c15_bench_p$pr.log 2>&1
  echo "== bench prio $pr exit $?"; grep '^{"metric"' gpurun_out/r2c15_bench_p$pr.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('   ms/step %.3f value %.0f e2e %.0f clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))
"
done
done
