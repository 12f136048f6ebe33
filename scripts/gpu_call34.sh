#!/bin/bash
# final bench line of the round at HEAD (the driver's command)
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c34_bench.json 2> gpurun_out/r2c34_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2c34_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value']/1e9,2), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e9,2), [round(s['ms']*1e3,1) for s in d['stages']], 'frac', d['roofline']['frac'], d['roofline']['moved_frac'], 'box3d', round(d['extra']['box3d']['value']/1e9,2), [(e['dim'],e['degree'],round(e['value']/1e9,1)) for e in d['extra']['elements']], d['clocks'])
"
