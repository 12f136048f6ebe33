#!/bin/bash
set -x
mkdir -p gpurun_out
for cfg in "2 1 1532 484 1" "2 2 1532 484 1" "2 3 1000 400 1" "2 4 800 300 1" "3 1 128 32 32" "3 2 64 32 32" "3 3 64 32 16"; do
  set -- $cfg
  timeout 400 python scripts/tune_stages.py --dim $1 --degree $2 --nx $3 --ny $4 --nz $5 >> gpurun_out/r2c5_tune.log 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 --extras none > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
SG_TILE=64 timeout 600 python bench.py --steps 20 --warmup 5 --extras none --no-cpu > gpurun_out/r2c5_bench_t64.json 2> gpurun_out/r2c5_bench_t64.err
