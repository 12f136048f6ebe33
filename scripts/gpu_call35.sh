#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/r2c35_smoke.log
timeout 100 python -m pytest tests/test_gpu_vectors.py tests/test_gpu_parity.py -m gpu -q -x -k "vector or 2-2 or 3-1 or 3-3" 2>&1 | tail -2 >> gpurun_out/r2c35_smoke.log
cat gpurun_out/r2c35_smoke.log
