#!/bin/bash
# Round 2, call 5: 64-channel CTA-pair tiles (A/B), wave-count tile selection, per-GPU batch sweep (strong-scaling proxy).
mkdir -p gpurun_out
for m in 2 1; do
  GDL_FLAT_PAIR64=$m timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -x -k "conv_fwd or conv_dgrad or film or gemm" > gpurun_out/r2c5_tests_pair64_$m.log 2>&1
  echo "== kernel tests PAIR64=$m exit $?"; tail -3 gpurun_out/r2c5_tests_pair64_$m.log | cut -c1-300
done
for m in 0 2 1; do
  GDL_FLAT_PAIR64=$m timeout 200 python tools/conv_bench.py --ops fwd,dgrad --out gpurun_out/r2c5_conv_bench.json --tag pair64_$m > gpurun_out/r2c5_conv_pair64_$m.log 2>&1
  echo "== conv_bench PAIR64=$m exit $?"; tail -1 gpurun_out/r2c5_conv_pair64_$m.log
done
for cfg in "CREMAD 256" "CREMAD 64" "VGGSound 128" "VGGSound 256" "KineticSound 64" "VGGSound 512"; do
  set -- $cfg
  GDL_DUMP_OPS=gpurun_out/r2c5_ops_$1_$2.json timeout 300 python bench.py --dataset $1 --batch $2 --steps 10 --warmup 3 --no-cpu --no-device-pipeline > gpurun_out/r2c5_bench_$1_$2.log 2>&1
  echo "== bench $1 B=$2 exit $?"; grep '^{"metric"' gpurun_out/r2c5_bench_$1_$2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
tot = sum(v['ms'] for v in kb.values())
print('   ms/step %.3f value %.0f e2e %.0f conv-frac %.3f serialised %.2f ms launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], tot, d['launches_per_step']))
print('   ' + ' '.join('%s=%.2f' % (k, v['ms']) for k, v in kb.items()))
"
done
