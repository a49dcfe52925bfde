#!/bin/bash
# 2 GPUs, exactly as the driver launches the scaling bench (weak scaling, default flags) + the 2-GPU pytest + dist check.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu --no-header -p no:cacheprovider -k "two_gpu or data_parallel or dist" > gpurun_out/r2n2_tests.log 2>&1
echo "== 2-GPU pytest exit $?"; tail -2 gpurun_out/r2n2_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_step_check.py > gpurun_out/r2n2_dist.log 2>&1
echo "== dist_step_check exit $?"; grep -E "vs the|DIST_STEP" gpurun_out/r2n2_dist.log | cut -c1-260 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2_bench.log 2>&1
echo "== bench --gpus 2 exit $?"; grep '^{"metric"' gpurun_out/r2n2_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('   ms/step %.3f value %.0f e2e %.0f (%s) host-frames %s frac %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['path'][:40], d.get('e2e_host_frames', {}).get('value'), d['roofline']['frac']))
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2n2_ref.log 2>&1
echo "== reference arm under torchrun exit $?"; tail -1 gpurun_out/r2n2_ref.log | cut -c1-200
