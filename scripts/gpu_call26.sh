#!/bin/bash
# how does the 3D P3 kernel scale with resident warps? 1 CTA per SM (3 warps) vs 2 (6 warps)
mkdir -p gpurun_out
L=gpurun_out/r2c26_warps.log
: > $L
t() { SG_ONLY_DEFAULT=1 timeout 300 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
SG_GRID_PER_SM=1 t --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --tag "1 CTA/SM"
t --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --tag "2 CTA/SM"
SG_GRID_PER_SM=1 t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --tag "1 CTA/SM"
t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --tag "2 CTA/SM"
cat $L
