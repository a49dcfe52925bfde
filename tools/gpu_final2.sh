#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_dp.sh
bash tools/profile_r1c.sh
