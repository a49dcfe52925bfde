#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; ( "$@" ) > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -4 gpurun_out/$name.log; }
run t_ew env timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "bn or maxpool or stem_tail or gap" --no-header -p no:cacheprovider
bash tools/gpu_step_check.sh ${1:-ew1}
