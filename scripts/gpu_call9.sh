#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 1 1532 484 1" "2 2 1532 484 1" "2 3 1000 400 1" "2 4 800 300 1" "3 1 128 32 32" "3 2 64 32 32" "3 3 64 32 16"; do
  set -- $cfg
  for mode in "" "SG_PDL_LATE=1" "SG_NO_PDL=1"; do
    echo -n "mode=[$mode] " >> gpurun_out/r2c9_pdl.log
    env $mode timeout 300 python scripts/perf_probe.py --dim $1 --degree $2 --nx $3 --ny $4 --nz $5 --steps 40 --reps 3 2>&1 | sort -t' ' -k9 | tail -1 >> gpurun_out/r2c9_pdl.log
  done
done
for mode in "" "SG_PDL_LATE=1" "SG_NO_PDL=1"; do
  env $mode timeout 600 python bench.py --steps 50 --warmup 5 --extras none --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench mode=[$mode]', d['value']/1e9, d['ms_per_step'])" >> gpurun_out/r2c9_pdl.log
done
