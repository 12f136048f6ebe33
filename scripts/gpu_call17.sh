#!/bin/bash
# upper bound of what staging the out-of-tile facet neighbours could give (--fake-intile: all gathers hit shared memory),
# on the tune box and on the box3d workload's cube; TILE 64 variant of 3D P3
mkdir -p gpurun_out
L=gpurun_out/r2c17_intile_bound.log
: > $L
t() { timeout 200 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
t --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --tag "box real"
t --dim 3 --degree 3 --nx 64 --ny 32 --nz 16 --fake-intile --tag "box fake-intile"
t --dim 3 --degree 3 --cube 26 --tag "cube26 real"
t --dim 3 --degree 3 --cube 26 --fake-intile --tag "cube26 fake-intile"
SG_ONLY_DEFAULT=1 t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --tag "box real"
SG_ONLY_DEFAULT=1 t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --fake-intile --tag "box fake-intile"
SG_ONLY_DEFAULT=1 t --dim 3 --degree 1 --nx 128 --ny 32 --nz 32 --tag "box real"
SG_ONLY_DEFAULT=1 t --dim 3 --degree 1 --nx 128 --ny 32 --nz 32 --fake-intile --tag "box fake-intile"
SG_ONLY_DEFAULT=1 t --dim 2 --degree 4 --nx 800 --ny 300 --tag "real"
SG_ONLY_DEFAULT=1 t --dim 2 --degree 4 --nx 800 --ny 300 --fake-intile --tag "fake-intile"
SG_ONLY_DEFAULT=1 t --dim 2 --degree 2 --tag "real"
SG_ONLY_DEFAULT=1 t --dim 2 --degree 2 --fake-intile --tag "fake-intile"
cat $L
