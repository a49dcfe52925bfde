#!/bin/bash
mkdir -p gpurun_out
for r in 0 1 2 3; do echo roles $r; GDL_STEM_ROLES=$r timeout 300 python tools/stem_bench.py 2>&1 | tail -2; done
echo grid296; GDL_STEM_ROLES=0 GDL_STEM_GRID=296 timeout 300 python tools/stem_bench.py 2>&1 | tail -2
echo grid296r3; GDL_STEM_ROLES=3 GDL_STEM_GRID=296 timeout 300 python tools/stem_bench.py 2>&1 | tail -2
