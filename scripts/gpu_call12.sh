#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2c12_n8.json 2> gpurun_out/r2c12_n8.err
timeout 400 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu --extras none > gpurun_out/r2c12_n1.json 2> gpurun_out/r2c12_n1.err
