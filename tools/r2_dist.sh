#!/bin/bash
# Round 2: multi-GPU correctness + the reference's entry point end to end + the fixed-global-batch (strong scaling) sweep.
#   gpurun --gpus N -- 'bash tools/r2_dist.sh N'      (N = 2 to validate the script, 8 for the record)
N=${1:-2}
OUT=gpurun_out/r2dist_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/gpus.txt; nproc >> $OUT/gpus.txt; free -g | head -2 >> $OUT/gpus.txt

if [ -z "$R2_DIST_ONLY_SCALING" ]; then
# 1. N-rank DGLStep == the oracle's N-shard simulation of nn.DataParallel (bf16 product path, then FP32 check mode)
timeout 400 $TR --nproc-per-node $N --master-port 29541 tests/dist_step_check.py > $OUT/dist_step_check.log 2>&1
echo "== dist_step_check N=$N exit $?"; grep -E "vs the|dist step|DIST_STEP|Error|assert" $OUT/dist_step_check.log | cut -c1-300 | tail -12

# 2. main_dgl.py under torchrun: train 2 epochs (DataParallel-chunk shards, valid each epoch, best-model save),
#    resume for a third epoch, then the evaluation-only path on the saved checkpoint
CK=$OUT/ckpt; rm -rf $CK
COMMON="--dataset CREMAD --fusion_method concat --fps 2 --alpha 4 --batch_size $((8*N)) --audio_path synthetic_exact --synthetic_len $((64*N)) --learning_rate 0.002"
( cd $OUT && timeout 600 $TR --nproc-per-node $N --master-port 29542 ../../main_dgl.py --train --ckpt_path ckpt --epochs 2 $COMMON > main_dgl_train.log 2>&1 )
echo "== main_dgl train N=$N exit $?"; grep -E "^Epoch|^Loss|Acc|saved|Error|Traceback" $OUT/main_dgl_train.log | cut -c1-200 | tail -12
BEST=$(ls $CK/*.pth 2>/dev/null | head -1)
if [ -n "$BEST" ]; then
  ( cd $OUT && timeout 600 $TR --nproc-per-node $N --master-port 29543 ../../main_dgl.py --train --ckpt_path ckpt --epochs 3 --resume "ckpt/$(basename $BEST)" $COMMON > main_dgl_resume.log 2>&1 )
  echo "== main_dgl resume exit $?"; grep -E "Resumed|^Epoch|^Loss|Error|Traceback" $OUT/main_dgl_resume.log | cut -c1-200 | tail -6
  ( cd $OUT && timeout 300 python ../../main_dgl.py --ckpt_path "ckpt/$(basename $BEST)" $COMMON > main_dgl_eval.log 2>&1 )
  echo "== main_dgl eval-only exit $?"; grep -E "loaded|Accuracy|Error|Traceback" $OUT/main_dgl_eval.log | cut -c1-200 | tail -4
  ls -la $CK | tail -5 > $OUT/ckpt_listing.txt; rm -f $CK/*.pth
fi

fi
# 3. fixed global batch (BASELINE configs 3-4): N ranks, then the smaller rank counts side by side on disjoint GPUs
bench() {  # dataset G ranks gpus port
  CUDA_VISIBLE_DEVICES=$4 timeout 400 $TR --nproc-per-node $3 --master-port $5 bench.py --gpus $3 --dataset $1 --global-batch $2 \
      --steps 10 --warmup 3 --no-cpu --no-roofline > $OUT/strong_$1_G$2_n$3.log 2>&1
}
ALL=$(seq -s, 0 $((N-1)))
for cfg in "VGGSound 1024" "KineticSound 512"; do
  set -- $cfg
  bench $1 $2 $N $ALL 29550
  if [ $N -eq 8 ]; then
    bench $1 $2 4 0,1,2,3 29551 &
    bench $1 $2 2 4,5 29552 &
    CUDA_VISIBLE_DEVICES=6 timeout 400 python bench.py --gpus 1 --dataset $1 --global-batch $2 --steps 10 --warmup 3 --no-cpu --no-device-pipeline --no-roofline > $OUT/strong_$1_G$2_n1.log 2>&1 &
    wait
  else
    CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --gpus 1 --dataset $1 --global-batch $2 --steps 10 --warmup 3 --no-cpu --no-device-pipeline --no-roofline > $OUT/strong_$1_G$2_n1.log 2>&1
  fi
  for f in $OUT/strong_$1_G$2_n*.log; do
    grep '^{"metric"' $f | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('   %s G=%d N=%d: %.0f samples/s, %.3f ms/step, e2e %.0f%s (%s)' % ('$1', $2, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], (' [host frames %.0f]' % d['e2e_host_frames']['value']) if 'e2e_host_frames' in d else '', d['scaling']))
" || echo "   $f: no result"
  done
done
python tools/strong_scaling_summary.py $OUT > $OUT/strong_scaling.json 2>/dev/null; cat $OUT/strong_scaling.json | head -40
