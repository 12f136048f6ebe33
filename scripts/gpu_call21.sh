#!/bin/bash
# AXPY operands straight from L2 (more CTAs per SM) for 2D P1/P2 and 3D P1; all 3D P2 variants on lattice tiles
mkdir -p gpurun_out
L=gpurun_out/r2c21_axs.log
: > $L
t() { timeout 300 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
t --dim 2 --degree 2
t --dim 2 --degree 1
t --dim 3 --degree 1 --nx 128 --ny 32 --nz 32
t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32
cat $L
