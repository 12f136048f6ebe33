#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2c29_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c29_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c29_bench.json 2> gpurun_out/r2c29_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c29_bench_ref.json 2> gpurun_out/r2c29_bench_ref.err
bash scripts/gpu_sanitize.sh
