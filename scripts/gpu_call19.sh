#!/bin/bash
# lattice-aligned Hilbert tiles + L1-prefetch code removed: parity (single + peers), per-element timings, bench
mkdir -p gpurun_out
L=gpurun_out/r2c19_lattice.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer.py tests/test_gmsh_mesh.py -q -x 2>&1 | tail -5 >> $L
t() { SG_ONLY_DEFAULT=1 timeout 200 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
t --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --tag "box"
t --dim 3 --degree 3 --cube 26 --tag "cube26"
t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --tag "box"
t --dim 3 --degree 1 --nx 128 --ny 32 --nz 32 --tag "box"
t --dim 2 --degree 1
t --dim 2 --degree 2
t --dim 2 --degree 3 --nx 1000 --ny 400
t --dim 2 --degree 4 --nx 800 --ny 300
timeout 600 python bench.py --steps 20 --warmup 5 --extras none > gpurun_out/r2c19_bench.json 2> gpurun_out/r2c19_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2c19_bench.json').read().strip().splitlines()[-1])
print('bench', d['value']/1e9, d['ms_per_step'], 'e2e', d['e2e']['value']/1e9, [round(s['ms']*1e3,1) for s in d['stages']], d['config'].get('setup_s'))
" >> $L
cat $L
