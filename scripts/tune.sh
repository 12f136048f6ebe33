#!/bin/bash
# sweep kernel variants on the GPU box (development aid)
run() { env "$@" python scripts/perf_probe.py $ARGS --reps 1 | tail -1; }
ARGS="--degree 2"
run SG_TILE=64 SG_PF=1; run SG_TILE=64 SG_PF=0; run SG_TILE=128 SG_PF=1
ARGS="--degree 1"
run SG_TILE=128 SG_PF=1; run SG_TILE=128 SG_PF=0; run SG_TILE=64
ARGS="--dim 3 --degree 1 --nx 128 --ny 32 --nz 32"
run SG_TILE=64 SG_PF=1; run SG_TILE=64 SG_PF=0; run SG_TILE=32 SG_SPLIT=3
ARGS="--degree 3 --nx 1000 --ny 400"; run SG_PF=1; run SG_PF=0
ARGS="--degree 4 --nx 800 --ny 300"; run SG_PF=1; run SG_PF=0
ARGS="--dim 3 --degree 2 --nx 64 --ny 32 --nz 32"; run SG_PF=1; run SG_PF=0
ARGS="--dim 3 --degree 3 --nx 64 --ny 32 --nz 16"; run SG_PF=1; run SG_PF=0
