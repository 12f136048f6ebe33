#!/bin/bash
# what each part of the in-kernel exchange costs at the N = 8 strong-scaling size (timings only: results are wrong with SG_EXCHANGE_DEBUG)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G', round(d['ms_per_step']*1e3,1), 'us/step')"; }
port=29700
for dbg in 0 1 2 4 6 7; do
  port=$((port+1))
  SG_EXCHANGE_DEBUG=$dbg timeout 300 $TR --master-port $port bench.py --gpus 2 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 dbg=$dbg" >> gpurun_out/r2c13_chain.log
done
