#!/bin/bash
# N=8 weak-scaling A/B of the overlapped two-bucket all-reduce (8x GPU-minutes: keep it short)
mkdir -p gpurun_out
for ov in 1 0; do
  GDL_AR_OVERLAP=$ov timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$ov bench.py --gpus 8 --steps 12 --warmup 4 --no-cpu --no-roofline > gpurun_out/bench_n8_ov$ov.log 2>&1
  echo "== bench N=8 overlap=$ov exit $?"; tail -1 gpurun_out/bench_n8_ov$ov.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('   ms/step %.3f value %.0f e2e %.0f clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']))
except Exception as e:
    print('   parse failed', e)
"
done
