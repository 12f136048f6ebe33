#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G', round(d['ms_per_step']*1e3,1), 'us/step', [round(s['ms']*1e3,1) for s in d['stages']], d['config']['cells_per_gpu'])"; }
SG_TEST_SPREAD=1 timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2c11_pytest_peer.log
port=29600
for tile in 64 128; do
  port=$((port+1))
  SG_TILE=$tile timeout 300 $TR --master-port $port bench.py --gpus 2 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 tile=$tile" >> gpurun_out/r2c11_small.log
done
port=$((port+1))
SG_PDL_EARLY=1 timeout 300 $TR --master-port $port bench.py --gpus 2 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 tile=128 pdl-early" >> gpurun_out/r2c11_small.log
SG_PDL_EARLY=1 timeout 300 python bench.py --gpus 1 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n1 tile=128 pdl-early" >> gpurun_out/r2c11_small.log
timeout 300 python bench.py --gpus 1 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n1 tile=128" >> gpurun_out/r2c11_small.log
port=$((port+1))
timeout 300 $TR --master-port $port bench.py --gpus 2 --scale 0.5 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 scale0.5" >> gpurun_out/r2c11_small.log
port=$((port+1))
timeout 600 $TR --master-port $port bench.py --gpus 2 --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2c11_n2.json
