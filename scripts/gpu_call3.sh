#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2c3_n2_fused.json 2> gpurun_out/r2c3_n2_fused.err
SG_PEER_SCHED_SPLIT=1 timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2c3_n2_split.json 2> gpurun_out/r2c3_n2_split.err
