#!/bin/bash
# small per-GPU problem (the size of one GPU's share in the N = 8 strong-scaling run): kernel variants and exchange cost
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G', round(d['ms_per_step']*1e3,1), 'us/step', [round(s['ms']*1e3,1) for s in d['stages']], d['config']['cells_per_gpu'])"; }
for tile in 32 64 128 256; do
  SG_TILE=$tile timeout 300 python bench.py --gpus 1 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n1 tile=$tile" >> gpurun_out/r2c10_small.log
  SG_TILE=$tile timeout 300 $TR --master-port 295$tile bench.py --gpus 2 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 tile=$tile" >> gpurun_out/r2c10_small.log
done
SG_TILE=64 SG_NO_PDL=1 timeout 300 python bench.py --gpus 1 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n1 tile=64 nopdl" >> gpurun_out/r2c10_small.log
SG_TILE=64 SG_NO_PDL=1 timeout 300 $TR --master-port 29577 bench.py --gpus 2 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 tile=64 nopdl" >> gpurun_out/r2c10_small.log
SG_TILE=64 SG_PEER_SCHED_SPLIT=1 timeout 300 $TR --master-port 29578 bench.py --gpus 2 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 tile=64 two-stream" >> gpurun_out/r2c10_small.log
for tile in 64 128; do
  SG_TILE=$tile timeout 300 $TR --master-port 296$tile bench.py --gpus 2 --scale 0.5 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 scale0.5 tile=$tile" >> gpurun_out/r2c10_small.log
  SG_TILE=$tile timeout 300 python bench.py --gpus 1 --scale 0.5 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n1 scale0.5 tile=$tile" >> gpurun_out/r2c10_small.log
done
