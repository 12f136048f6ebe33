#!/bin/bash
# A/B: qall (all-rows F-type facet term) folded weights vs weighted jump first; libs built with -DSG_QALL_DELTA=<min D*ND>
mkdir -p gpurun_out
L=gpurun_out/r2c20_qall.log
: > $L
t() { SG_ONLY_DEFAULT=1 timeout 200 python scripts/tune_stages.py "$@" 2>&1 | grep -v "^Creat\|^Number" >> $L; }
for lib in libseigen_b200.so libseigen_b200_delta30.so libseigen_b200_delta0.so; do
  export SG_LIB=$PWD/seigen_b200/$lib
  t --dim 3 --degree 2 --nx 64 --ny 32 --nz 32 --tag "$lib"
  if [ $lib != libseigen_b200_delta30.so ]; then
    t --dim 3 --degree 1 --nx 128 --ny 32 --nz 32 --tag "$lib"
    t --dim 2 --degree 2 --tag "$lib"
    t --dim 2 --degree 3 --nx 1000 --ny 400 --tag "$lib"
  fi
  t --dim 2 --degree 4 --nx 800 --ny 300 --tag "$lib"
done
cat $L
