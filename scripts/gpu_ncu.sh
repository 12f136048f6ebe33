#!/bin/bash
# ncu evidence for profiles/: --set full of the six passes (2D P2 headline workload, 3D P1 / P2 / P3) + launch list
set -x
mkdir -p gpurun_out
scripts/ncu_full.sh r2_full_2d_p2 python bench.py --steps 2 --warmup 3 --extras none --no-cpu
scripts/ncu_full.sh r2_full_3d_p1 python scripts/perf_probe.py --dim 3 --degree 1 --nx 128 --ny 32 --nz 32 --steps 2 --reps 1
scripts/ncu_full.sh r2_full_3d_p2 python scripts/perf_probe.py --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --steps 2 --reps 1
KEEP_REP=1 scripts/ncu_full.sh r2_full_3d_p3 python scripts/perf_probe.py --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --steps 2 --reps 1
scripts/ncu_full.sh r2_full_2d_p4 python scripts/perf_probe.py --dim 2 --degree 4 --nx 800 --ny 300 --steps 2 --reps 1
scripts/ncu_full.sh r2_full_2d_p3 python scripts/perf_probe.py --dim 2 --degree 3 --nx 1000 --ny 400 --steps 2 --reps 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 5 --extras none --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1
