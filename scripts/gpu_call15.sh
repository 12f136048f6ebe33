#!/bin/bash
# auto carve-out (per kernel) vs uniform settings: step time per element + headline bench
mkdir -p gpurun_out
for cfg in "2 2 1532 484 1" "2 1 1532 484 1" "2 3 1000 400 1" "2 4 800 300 1" "3 1 128 32 32" "3 2 64 32 32" "3 3 64 32 16"; do
  set -- $cfg
  for mode in "SG_CARVEOUT=0" "SG_CARVEOUT=100" "SG_CARVEOUT=86" "SG_CARVEOUT=72"; do
    env $mode SG_ONLY_DEFAULT=1 timeout 300 python scripts/tune_stages.py --dim $1 --degree $2 --nx $3 --ny $4 --nz $5 --tag "[$mode]" 2>&1 | head -1 >> gpurun_out/r2c15_carveout.log
  done
done
for mode in "SG_CARVEOUT=0" "SG_CARVEOUT=100"; do
  env $mode timeout 600 python bench.py --steps 50 --warmup 5 --extras none --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench [$mode]', d['value']/1e9, d['ms_per_step'], [round(s['ms']*1e3,1) for s in d['stages']])" >> gpurun_out/r2c15_carveout.log
  env $mode timeout 600 python bench.py --workload box3d --steps 50 --warmup 5 --extras none --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('box3d [$mode]', d['value']/1e9, d['ms_per_step'], [round(s['ms']*1e3,1) for s in d['stages']])" >> gpurun_out/r2c15_carveout.log
done
