#!/bin/bash
# Runs each GPU test group in its own process (a trapped kernel poisons the CUDA context),
# each under a timeout, and collects the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "$@"; do
  name=$(echo "$grp" | tr -c 'A-Za-z0-9_' '_')
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$grp" --no-header -p no:cacheprovider \
      > gpurun_out/probe_${name}.log 2>&1
  echo "== $grp exit $?" | tee -a gpurun_out/probe_summary.txt
  tail -5 gpurun_out/probe_${name}.log | tee -a gpurun_out/probe_summary.txt
done
