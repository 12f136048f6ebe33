#!/bin/bash
# 8 GPUs: weak (headline workload) + strong + box3d through bench.py's extras, then N = 1 on the same box
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c8_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1200 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2c8_n8.json 2> gpurun_out/r2c8_n8.err
timeout 900 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/r2c8_n4.json 2> gpurun_out/r2c8_n4.err
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu > gpurun_out/r2c8_n1.json 2> gpurun_out/r2c8_n1.err
