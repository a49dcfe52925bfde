#!/bin/bash
# ncu --set full captures of the HBM-bound kernels that sit furthest below the measured copy bandwidth.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --no-device-pipeline --batch 256"
NCU="ncu --set full --clock-control none --import-source on"
K='regex:bn_bwd_nores_apply_kernel|bn_bwd_nores_reduce_kernel|bn_relu_maxpool_fwd2_kernel|bn_relu_maxpool_bwd_apply_kernel|bn_relu_maxpool_bwd_reduce_kernel|stem_layout_kernel|bn_bwd_apply_kernel|stem_wgrad_kernel'
# the visual encoder's launches come second: skip the audio ones (-s) and take a window that holds one of each
timeout 900 $NCU -k "$K" -s 60 -c 40 -f -o gpurun_out/prof_hbm_r2 $BENCH > gpurun_out/prof_hbm_r2.log 2>&1
echo "== hbm capture exit $?"; tail -3 gpurun_out/prof_hbm_r2.log
ls -la gpurun_out/prof_hbm_r2.ncu-rep
