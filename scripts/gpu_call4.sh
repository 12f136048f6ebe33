#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c4_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
SG_NO_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --extras none --no-cpu > gpurun_out/r2c4_bench_nopdl.json 2> gpurun_out/r2c4_bench_nopdl.err
timeout 300 python scripts/tune_stages.py --dim 2 --degree 2 --nx 1532 --ny 484 > gpurun_out/r2c4_tune_2d_p2.log 2>&1
timeout 300 python scripts/tune_stages.py --dim 3 --degree 1 --nx 128 --ny 32 --nz 32 > gpurun_out/r2c4_tune_3d_p1.log 2>&1
timeout 300 python scripts/tune_stages.py --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 > gpurun_out/r2c4_tune_3d_p2.log 2>&1
timeout 300 python scripts/tune_stages.py --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 > gpurun_out/r2c4_tune_3d_p3.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2c4_bench_ref.json 2> gpurun_out/r2c4_bench_ref.err
