#!/bin/bash
# Round-2 first step: validate and measure the CTA-pair (multicast weight tiles) variant of conv_flat.
# Every run is wrapped in `timeout`: a protocol bug in the pair handshake traps (bounded mbarrier waits) rather than hangs.
mkdir -p gpurun_out
export GDL_FLAT_CLUSTER=1
timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -x -k "conv_fwd or conv_dgrad" > gpurun_out/kernel_tests_cluster.log 2>&1
echo "== kernel tests (cluster) exit $?"; tail -3 gpurun_out/kernel_tests_cluster.log
timeout 300 python -m pytest tests/test_gpu_step.py tests/test_gpu_compat.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/step_tests_cluster.log 2>&1
echo "== step tests (cluster) exit $?"; tail -3 gpurun_out/step_tests_cluster.log
for c in 1 0 1 0; do
  GDL_FLAT_CLUSTER=$c GDL_DUMP_OPS=gpurun_out/ops_cluster$c.json timeout 120 python bench.py --steps 15 --warmup 4 --no-cpu --no-device-pipeline > gpurun_out/bench_cluster$c.log 2>&1
  echo "== bench cluster=$c exit $?"; grep '^{"metric"' gpurun_out/bench_cluster$c.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f frac %.3f fwd %.2f dgrad %.2f' % (d['ms_per_step'], d['value'], d['roofline']['frac'], kb['conv_fwd']['ms'], kb['conv_dgrad']['ms']))
"
done
