#!/bin/bash
# Round-2 final evidence (1 GPU, under gpurun): parity log, ncu launch list, DRAM-traffic + tensor-pipe metrics over every
# implicit-GEMM launch of two steps, HBM kernels' dram throughput, --set full captures of the CTA-pair kernels and the
# weight-gradient kernel, compute-sanitizer memcheck + racecheck over the tiny kernel cases.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_check_mode.py tests/test_gpu_compat.py -q -m gpu --no-header -p no:cacheprovider -s \
    > gpurun_out/r2_parity_raw.log 2>&1
echo "== parity tests exit $?"; grep -E "passed|failed|^FAILED" gpurun_out/r2_parity_raw.log | tail -3
grep -E "^forced|^B=256|^check-mode|^eval logits|^tail-batch|passed|failed" gpurun_out/r2_parity_raw.log | cut -c1-600 > gpurun_out/r2_parity.log
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --no-device-pipeline --batch ${PROF_BATCH:-256}"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1150 -c 900 --csv --log-file gpurun_out/launches_r2.csv $BENCH > gpurun_out/launches_bench_r2.log 2>&1
echo "== launch list exit $?"
CONV='regex:conv_flat_kernel|conv_flat2_kernel|conv_wgrad_flat_kernel|stem_fwd_kernel|stem_wgrad_kernel|conv_igemm_kernel'
timeout 700 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k "$CONV" -s 236 -c 236 --csv --log-file gpurun_out/conv_traffic_r2.csv $BENCH > gpurun_out/conv_traffic_bench_r2.log 2>&1
echo "== conv traffic exit $?"
HBM='regex:bn_apply_kernel|bn_bwd|bn_relu_maxpool|stem_layout_kernel|sgd_momentum_kernel|pack_weights|wgrad_reduce'
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k "$HBM" -s 300 -c 300 --csv --log-file gpurun_out/hbm_kernels_r2.csv $BENCH > gpurun_out/hbm_bench_r2.log 2>&1
echo "== hbm kernels exit $?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k 'regex:conv_flat2_kernel|conv_wgrad_flat_kernel|stem_wgrad_kernel|bn_relu_maxpool_fwd3_kernel|bn_bwd_nores_apply_kernel' -s 60 -c 24 -f -o gpurun_out/prof_final_r2 $BENCH > gpurun_out/prof_final_r2.log 2>&1
echo "== full captures exit $?"
SAN_K='(test_conv_fwd or test_conv_dgrad or test_conv_wgrad or test_bn_fwd_bwd or test_dgl_head or test_stem) and not bias_act and not case12 and not case13 and not case15 and not case17 and not case18 and not case19 and not generic and not 224 and not 257'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -x -k "$SAN_K" > gpurun_out/sanitizer_${tool}_r2.log 2>&1
  echo "== compute-sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_r2.log | tail -3
done
# BASELINE config 2: the other three heads at batch 256, and the stock-PyTorch (cuDNN / cuBLAS) side bar
for f in concat sum gated film; do
  timeout 300 python bench.py --fusion $f --steps 10 --warmup 3 --no-cpu --no-device-pipeline --no-roofline > gpurun_out/r2_variant_$f.log 2>&1
  echo "== bench $f exit $?"
done
timeout 600 python bench.py --impl cudnn_sidebar --steps 5 --warmup 3 > gpurun_out/r2_variant_sidebar.log 2>&1
echo "== cudnn sidebar exit $?"
python - <<'PY'
import json
out = {"what": "bench.py --fusion F --steps 10 --warmup 3 (batch 256, CREMA-D shape, one B200, same box) and --impl cudnn_sidebar"}
for f in ("concat", "sum", "gated", "film", "sidebar"):
    try:
        line = [l for l in open("gpurun_out/r2_variant_%s.log" % f) if l.startswith("{")][-1]
        d = json.loads(line)
        out[f] = {k: d[k] for k in ("value", "ms_per_step", "unit") if k in d}
        if "e2e" in d: out[f]["e2e"] = d["e2e"]["value"]
        if "launches_per_step" in d: out[f]["launches_per_step"] = d["launches_per_step"]
        if "modes" in d: out[f]["modes"] = d["modes"]
    except Exception as e:
        out[f] = {"error": str(e)}
json.dump(out, open("gpurun_out/r2_variants.json", "w"), indent=1)
print(json.dumps(out)[:900])
PY
ls -la gpurun_out/*r2* | head -30
