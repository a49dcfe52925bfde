#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_datapipe.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/datapipe_tests.log 2>&1
echo "== datapipe tests exit $?"; tail -15 gpurun_out/datapipe_tests.log
timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu --no-roofline > gpurun_out/bench_dp.log 2>&1
echo "== bench exit $?"; tail -1 gpurun_out/bench_dp.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('   ms/step %.3f value %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value'])); print('   device pipeline', d.get('e2e_device_pipeline'))
"
