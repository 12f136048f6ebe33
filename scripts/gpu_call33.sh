#!/bin/bash
# cells ordered inside their tiles so that out-of-tile gathers share L2 sectors: A/B per element, then the GPU suite and the bench
mkdir -p gpurun_out
L=gpurun_out/r2c33_tile_order.log
: > $L
t() { SG_ONLY_DEFAULT=1 timeout 200 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
for o in 0 1; do
  export SG_TILE_ORDER=$o
  t --dim 2 --degree 2 --tag "order=$o"
  t --dim 2 --degree 3 --nx 1000 --ny 400 --tag "order=$o"
  t --dim 3 --degree 1 --nx 128 --ny 32 --nz 32 --tag "order=$o"
  t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --tag "order=$o"
  t --dim 3 --degree 3 --cube 26 --tag "order=$o cube26"
done
unset SG_TILE_ORDER
cat $L
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2c33_pytest.log
cat gpurun_out/r2c33_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c33_bench.json 2> gpurun_out/r2c33_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2c33_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value']/1e9,2), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e9,2), [round(s['ms']*1e3,1) for s in d['stages']], 'box3d', round(d['extra']['box3d']['value']/1e9,2), [(e['dim'],e['degree'],round(e['value']/1e9,1)) for e in d['extra']['elements']], d['config'].get('setup_s'))
"
