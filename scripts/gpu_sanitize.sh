#!/bin/bash
# compute-sanitizer evidence (profiles/r02_sanitizer_*.log)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py single > gpurun_out/r2_sanitizer_${tool}_single.log 2>&1
  echo "exit code: $?" >> gpurun_out/r2_sanitizer_${tool}_single.log
done
timeout 900 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 python scripts/sanitize_case.py peers > gpurun_out/r2_sanitizer_memcheck_peers.log 2>&1
echo "exit code: $?" >> gpurun_out/r2_sanitizer_memcheck_peers.log
