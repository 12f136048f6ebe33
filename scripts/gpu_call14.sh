#!/bin/bash
# L1 prefetch of out-of-tile neighbour rows: on/off x shared-memory carve-out, per element (per-pass times)
mkdir -p gpurun_out
for cfg in "2 2 1532 484 1" "2 1 1532 484 1" "3 1 128 32 32" "3 2 64 32 32" "3 3 64 32 16" "2 3 1000 400 1"; do
  set -- $cfg
  for mode in "SG_PREFETCH=0" "SG_PREFETCH=1" "SG_PREFETCH=1 SG_CARVEOUT=75" "SG_PREFETCH=0 SG_CARVEOUT=75"; do
    env $mode SG_ONLY_DEFAULT=1 timeout 300 python scripts/tune_stages.py --dim $1 --degree $2 --nx $3 --ny $4 --nz $5 --tag "[$mode]" 2>&1 | head -1 >> gpurun_out/r2c14_prefetch.log
  done
done
