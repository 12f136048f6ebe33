#!/bin/bash
# facet staging (cp.async) for the 2D all-rows kernels: parity + timings
mkdir -p gpurun_out
L=gpurun_out/r2c28_staging2d.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "facet_staging" 2>&1 | tail -6 >> $L
t() { timeout 300 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" | head -2 >> $L; }
t --dim 2 --degree 2
t --dim 2 --degree 1
t --dim 2 --degree 3 --nx 1000 --ny 400
cat $L
