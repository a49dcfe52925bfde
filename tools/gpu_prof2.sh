#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
GDL_WFLAT=2 $NCU -k regex:conv_wgrad_flat -c 2 -f -o gpurun_out/p2_wflat_l3 python tools/conv_bench.py --only v.l3.s1 --ops wgrad --reps 1 --out /tmp/x.json > gpurun_out/p2_a.log 2>&1
GDL_WFLAT=0 $NCU -k regex:conv3x3_wgrad_halo -c 2 -f -o gpurun_out/p2_whalo_l3 python tools/conv_bench.py --only v.l3.s1 --ops wgrad --reps 1 --out /tmp/x.json > gpurun_out/p2_b.log 2>&1
$NCU -k regex:conv_flat_kernel -c 2 -f -o gpurun_out/p2_flat_l4 python tools/conv_bench.py --only v.l4.s1 --ops fwd --reps 1 --out /tmp/x.json > gpurun_out/p2_c.log 2>&1
$NCU -k regex:conv_flat_kernel -c 2 -f -o gpurun_out/p2_flat_l1 python tools/conv_bench.py --only v.l1.s1 --ops fwd --reps 1 --out /tmp/x.json > gpurun_out/p2_d.log 2>&1
ls -la gpurun_out/p2_*
