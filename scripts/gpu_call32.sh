#!/bin/bash
# plain passes of 2D P2 with 5 / 6 CTAs per SM (register caps 96 / 80) and single-stage pipelines
mkdir -p gpurun_out
L=gpurun_out/r2c32_plain_occ.log
timeout 300 python scripts/tune_stages.py --dim 2 --degree 2 2>&1 | grep -v "^Creat\|^Number" | head -5 > $L
cat $L
