#!/bin/bash
# facet staging (cp.async of out-of-tile facet neighbours): parity, then per-pass timings against the unstaged variants
mkdir -p gpurun_out
L=gpurun_out/r2c18_staging.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "facet_staging" 2>&1 | tail -15 >> $L
t() { timeout 200 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
SG_ONLY_STG=1 t --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --tag "box"
SG_ONLY_STG=1 t --dim 3 --degree 3 --cube 26 --tag "cube26"
SG_ONLY_STG=1 t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --tag "box"
cat $L
