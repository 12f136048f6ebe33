#!/bin/bash
# N = 2 with the final build: peer tests one rank per GPU, the driver's bench command (weak + extras: strong, box3d), small problem
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G', round(d['ms_per_step']*1e3,1), 'us/step', [round(s['ms']*1e3,1) for s in d['stages']], d['config']['cells_per_gpu'])"; }
SG_TEST_SPREAD=1 timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2c30_pytest_peer.log
timeout 600 $TR --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/r2c30_n2.err | tail -1 > gpurun_out/r2c30_n2.json
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --extras none --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2c30_n1.json
: > gpurun_out/r2c30_small.log
timeout 300 $TR --master-port 29712 bench.py --gpus 2 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n2 small" >> gpurun_out/r2c30_small.log
timeout 300 python bench.py --gpus 1 --scale 0.354 --steps 300 --warmup 20 --extras none --no-cpu 2>/dev/null | summ "n1 small" >> gpurun_out/r2c30_small.log
cat gpurun_out/r2c30_pytest_peer.log gpurun_out/r2c30_small.log
python -c "
import json
for f in ('r2c30_n1','r2c30_n2'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f, round(d['value']/1e9,2), d['ms_per_step'], {k:(round(v['value']/1e9,2), v.get('ms_per_step')) for k,v in d.get('extra',{}).items() if isinstance(v,dict) and 'value' in v})
"
