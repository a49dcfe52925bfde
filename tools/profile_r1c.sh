#!/bin/bash
# Final ncu evidence of the round (1 GPU, under gpurun): launch list, DRAM traffic of every implicit-GEMM launch of
# the step (metrics-only pass), full captures of the kernels changed this session.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --batch ${PROF_BATCH:-256}"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/launches_r1c.csv $BENCH > gpurun_out/launches_bench_r1c.log 2>&1
echo "== launch list exit $?"
CONV='regex:conv_flat_kernel|conv_wgrad_flat_kernel|stem_fwd_kernel|stem_wgrad_kernel|conv_igemm_kernel'
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k "$CONV" -s 236 -c 236 --csv --log-file gpurun_out/conv_traffic_r1c.csv $BENCH > gpurun_out/conv_traffic_bench_r1c.log 2>&1
echo "== conv traffic exit $?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:conv_wgrad_flat_kernel -s 60 -c 6 -f -o gpurun_out/prof_wflat_r1c $BENCH > gpurun_out/prof_wflat_r1c.log 2>&1
echo "== wgrad capture exit $?"
timeout 400 $NCU -k regex:"bn_relu_maxpool_fwd2|bn_relu_maxpool_bwd_apply|wgrad_reduce_t" -s 12 -c 8 -f -o gpurun_out/prof_tail_r1c $BENCH > gpurun_out/prof_tail_r1c.log 2>&1
echo "== tail capture exit $?"
ls -la gpurun_out/*r1c*
