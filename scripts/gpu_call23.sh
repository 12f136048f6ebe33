#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c23_ns1.log
: > $L
t() { timeout 300 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
t --dim 2 --degree 1
t --dim 2 --degree 3 --nx 1000 --ny 400
t --dim 2 --degree 4 --nx 800 --ny 300
cat $L
