#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "== bench N=$N exit $?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-1500
