#!/bin/bash
# Soak: the reference entry point for 6 epochs x 16 steps of batch 64 (class-structured synthetic data, device-side
# crop / resize and log-STFT), valid() after every epoch, best-model save, then resume + evaluation of the saved file.
mkdir -p gpurun_out/r2soak && cd gpurun_out/r2soak && rm -rf ckpt
COMMON="--dataset CREMAD --fusion_method concat --fps 3 --alpha 4 --batch_size 64 --audio_path synthetic_device --synthetic_len 1024 --learning_rate 0.002"
timeout 900 python ../../main_dgl.py --train --ckpt_path ckpt --epochs 6 $COMMON > train.log 2>&1
echo "== train exit $?"; grep -E "^Loss|^Epoch|saved|Traceback|Error|nan" train.log | cut -c1-160 | tail -14
BEST=$(ls -t ckpt/*.pth 2>/dev/null | head -1)
if [ -n "$BEST" ]; then
  timeout 300 python ../../main_dgl.py --ckpt_path "$BEST" $COMMON > eval.log 2>&1
  echo "== eval exit $?"; grep -E "loaded|Accuracy|Traceback" eval.log | tail -3
  ls -la ckpt | tail -4 > ckpt_listing.txt; rm -f ckpt/*.pth
fi
