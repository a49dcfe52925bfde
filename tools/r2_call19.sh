#!/bin/bash
# Round 2, call 19: 64 samples per GPU (KineticSound shape): CTA-pair kernels on / off, resident pair on / off.
mkdir -p gpurun_out
run() {
  local label=$1; shift
  env "$@" timeout 200 python bench.py --dataset KineticSound --batch 64 --steps 30 --warmup 5 --no-cpu --no-device-pipeline --no-roofline > gpurun_out/r2c19_$label.log 2>&1
  grep '^{"metric"' gpurun_out/r2c19_$label.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$label: ms/step %.3f value %.0f' % (d['ms_per_step'], d['value']))
"
}
run default
run pair0 GDL_FLAT_PAIR=0
run res0 GDL_FLAT_PAIR64RES=0
run default2
run pair0b GDL_FLAT_PAIR=0
