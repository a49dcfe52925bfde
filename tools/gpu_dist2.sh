#!/bin/bash
# 2-GPU: data-parallel parity (bucketed/overlapped all-reduce on and off) + N=2 bench A/B
mkdir -p gpurun_out
TAG=${1:-n2}
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "stem_tail or maxpool or wgrad or bn_" > gpurun_out/kernel_tests_$TAG.log 2>&1
echo "== kernel tests exit $?"; tail -3 gpurun_out/kernel_tests_$TAG.log
GDL_DUMP_OPS=gpurun_out/ops_${TAG}_1gpu.json timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu > gpurun_out/bench_${TAG}_1gpu.log 2>&1
echo "== bench N=1 exit $?"; tail -1 gpurun_out/bench_${TAG}_1gpu.log | cut -c1-300
for ov in 1 0; do
  GDL_AR_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$ov tests/dist_step_check.py > gpurun_out/dist_check_${TAG}_ov$ov.log 2>&1
  echo "== dist check overlap=$ov exit $?"; tail -4 gpurun_out/dist_check_${TAG}_ov$ov.log | cut -c1-300
done
for ov in 1 0 1 0; do
  GDL_AR_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$ov bench.py --gpus 2 --steps 15 --warmup 4 --no-cpu --no-roofline > gpurun_out/bench_${TAG}_ov$ov.log 2>&1
  echo "== bench N=2 overlap=$ov exit $?"; tail -1 gpurun_out/bench_${TAG}_ov$ov.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('   ms/step %.3f value %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
except Exception as e:
    print('   parse failed', e)
"
done
