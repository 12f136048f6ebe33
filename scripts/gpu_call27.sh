#!/bin/bash
# warp-task F-type kernel for 3D P3 (8 warps per SM): parity, 2-rank exchange, timings
mkdir -p gpurun_out
L=gpurun_out/r2c27_warp_tasks.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "warp_task or 1-1 or 1-2" 2>&1 | tail -8 >> $L
timeout 600 python -m pytest tests/test_gpu_peer.py -q -k "env12 or env11" 2>&1 | tail -8 >> $L
t() { timeout 300 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
t --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --tag box
t --dim 3 --degree 3 --cube 26 --tag cube26
cat $L
