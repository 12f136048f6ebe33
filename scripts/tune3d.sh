#!/bin/bash
# sweep the 3D kernel variants on the GPU box (development aid)
run() { env "$@" python scripts/perf_probe.py $ARGS --reps 1 | tail -1; }
ARGS="--dim 3 --degree 1 --nx 128 --ny 32 --nz 32"
run SG_TILE=64 SG_SPLIT=1 SG_MINB=4; run SG_TILE=64 SG_SPLIT=1 SG_MINB=8; run SG_TILE=64 SG_SPLIT=1 SG_MINB=6
run SG_TILE=128 SG_SPLIT=1 SG_MINB=4; run SG_TILE=32 SG_SPLIT=1 SG_MINB=8; run SG_TILE=32 SG_SPLIT=3 SG_MINB=4
run SG_TILE=64 SG_SPLIT=3 SG_MINB=4
ARGS="--dim 3 --degree 2 --nx 64 --ny 32 --nz 32"
run SG_TILE=32 SG_MINB=3; run SG_TILE=32 SG_MINB=4; run SG_TILE=32 SG_MINB=5; run SG_TILE=64 SG_MINB=2
ARGS="--dim 3 --degree 3 --nx 64 --ny 32 --nz 16"
run SG_MINB=2; run SG_MINB=3; run SG_MINB=4
