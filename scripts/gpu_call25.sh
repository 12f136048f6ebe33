#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c25_pytest.log
cat gpurun_out/r2c25_pytest.log
