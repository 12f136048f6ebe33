#!/bin/bash
# N = 8 with the final build: the driver's bench command (weak + extras: strong, box3d) and N = 1 on the same box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29721 bench.py --gpus 8 --steps 50 --warmup 5 2>gpurun_out/r2c31_n8.err | tail -1 > gpurun_out/r2c31_n8.json
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --extras none --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2c31_n1.json
python -c "
import json
for f in ('r2c31_n1','r2c31_n8'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f, round(d['value']/1e9,2), d['ms_per_step'], d['clocks'], {k:(round(v['value']/1e9,2), v.get('ms_per_step')) for k,v in d.get('extra',{}).items() if isinstance(v,dict) and 'value' in v})
"
