#!/bin/bash
# Closing call: the whole GPU suite + smoke + default bench + reference arm on the final library, and one --set full
# capture of the CTA-pair convolution launches of the VISUAL encoder (the large layers).
bash tools/r2_full.sh
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2full_reference.log 2>&1
echo "== reference arm exit $?"
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-roofline --no-cpu --no-device-pipeline --batch 256"
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:conv_flat2_kernel' -s 196 -c 40 -f -o gpurun_out/prof_flat2_visual_r2 $BENCH > gpurun_out/prof_flat2_visual_r2.log 2>&1
echo "== visual pair-kernel capture exit $?"
