#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ew_bench.py 2>&1 | tail -4
