#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c24_ns1_3d.log
: > $L
t() { timeout 300 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
t --dim 3 --degree 3 --nx 64 --ny 32 --nz 16
t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32
SG_ONLY_DEFAULT=1 t --dim 2 --degree 3 --nx 1000 --ny 400
cat $L
