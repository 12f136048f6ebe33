#!/bin/bash
# N = 2: peer tests across NVLink, weak / strong / box3d with the fused and the two-stream exchange schedules
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c2_topo.txt 2>&1
SG_TEST_SPREAD=1 timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c2_pytest_peer.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2c2_n2_fused.json 2> gpurun_out/r2c2_n2_fused.err
SG_PEER_SCHED_SPLIT=1 timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2c2_n2_split.json 2> gpurun_out/r2c2_n2_split.err
NCU_LOG=gpurun_out/r2c2_n2_launches.csv NCU_COUNT=200 timeout 600 $TR --master-port 29513 --no-python scripts/rank0_ncu.sh bench.py --gpus 2 --steps 3 --warmup 3 --extras none --no-cpu > gpurun_out/r2c2_n2_ncu.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --extras none --no-cpu > gpurun_out/r2c2_n1.json 2> gpurun_out/r2c2_n1.err
