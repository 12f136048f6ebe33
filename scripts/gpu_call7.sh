#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c7_pytest.log
for cfg in "2 3 1000 400 1" "2 4 800 300 1" "3 1 128 32 32" "3 2 64 32 32" "3 3 64 32 16"; do
  set -- $cfg
  timeout 400 python scripts/tune_stages.py --dim $1 --degree $2 --nx $3 --ny $4 --nz $5 >> gpurun_out/r2c7_tune.log 2>&1
done
SG_NO_PDL=1 timeout 400 python scripts/tune_stages.py --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --tag nopdl >> gpurun_out/r2c7_tune.log 2>&1
timeout 600 python bench.py --workload box3d --steps 50 --warmup 5 --extras none --no-cpu > gpurun_out/r2c7_box3d.json 2> gpurun_out/r2c7_box3d.err
SG_NO_PDL=1 timeout 600 python bench.py --workload box3d --steps 50 --warmup 5 --extras none --no-cpu > gpurun_out/r2c7_box3d_nopdl.json 2> gpurun_out/r2c7_box3d_nopdl.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
