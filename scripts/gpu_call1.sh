#!/bin/bash
# first GPU call of round 2: tests, bench with timers, element probes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -30 > gpurun_out/r2c1_pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c1_pytest_fullsize.log
SG_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
for cfg in "2 1 1532 484 1" "2 2 1532 484 1" "2 3 1000 400 1" "2 4 800 300 1" "3 1 128 32 32" "3 2 64 32 32" "3 3 64 32 16"; do
  set -- $cfg
  timeout 300 python scripts/perf_probe.py --dim $1 --degree $2 --nx $3 --ny $4 --nz $5 --steps 20 --reps 2 2>&1 | tail -1 >> gpurun_out/r2c1_probe.log
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1
