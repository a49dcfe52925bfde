#!/bin/bash
# Weak scaling (256 samples per GPU, the driver's launch line) on one 8-GPU box: N=8 alone, then N=4 / 2 / 1 side by side
# on disjoint GPUs.
OUT=gpurun_out/r2weak; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 $TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 --no-roofline > $OUT/weak_n8.log 2>&1
echo "== N=8 exit $?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 500 $TR --nproc-per-node 4 --master-port 29572 bench.py --gpus 4 --steps 20 --warmup 5 --no-roofline > $OUT/weak_n4.log 2>&1 &
CUDA_VISIBLE_DEVICES=4,5 timeout 500 $TR --nproc-per-node 2 --master-port 29573 bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline > $OUT/weak_n2.log 2>&1 &
CUDA_VISIBLE_DEVICES=6 timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-roofline --no-cpu > $OUT/weak_n1.log 2>&1 &
wait
python - <<'PY'
import json
out = {"what": "weak scaling, 256 samples per GPU, CREMA-D shape, ConcatFusion_DGL; bench.py launched as the driver does; N=4/2/1 ran side by side on disjoint GPUs of the same box", "results": {}}
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open("gpurun_out/r2weak/weak_n%d.log" % n) if l.startswith('{"metric"')][-1])
        r = {"samples_per_s": round(d["value"], 1), "ms_per_step": round(d["ms_per_step"], 3), "e2e_samples_per_s": round(d["e2e"]["value"], 1),
             "e2e_path": d["e2e"].get("path"), "clocks": d.get("clocks")}
        for k in ("e2e_host_frames", "e2e_device_pipeline"):
            if k in d: r[k + "_samples_per_s"] = round(d[k]["value"], 1)
        if n == 1: base = d["value"]
        if base: r["efficiency"] = round(d["value"] / base / n, 4)
        out["results"][str(n)] = r
        print("   N=%d: %.0f samples/s, %.3f ms/step, e2e %.0f, efficiency %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("efficiency")))
    except Exception as e:
        out["results"][str(n)] = {"error": str(e)}; print("   N=%d: %s" % (n, e))
json.dump(out, open("gpurun_out/r2weak/weak_scaling.json", "w"), indent=1)
PY
