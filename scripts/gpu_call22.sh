#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c22_ns1.log
: > $L
t() { timeout 300 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
t --dim 2 --degree 2
SG_ONLY_DEFAULT=1 t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32
cat $L
