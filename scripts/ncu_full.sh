#!/bin/bash
# ncu --set full capture of the first six stage launches of a command, summarised on the box (development aid):
#   scripts/ncu_full.sh <name> <command...>   ->  gpurun_out/<name>.md (+ gpurun_out/<name>.ncu-rep if KEEP_REP=1)
name=$1; shift
ncu --set full --clock-control none --import-source on -k regex:stage_ -c 6 -f -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1
python scripts/summarize_ncu.py full /tmp/$name.ncu-rep gpurun_out/$name.md >> gpurun_out/$name.log 2>&1
if [ -n "$KEEP_REP" ]; then cp /tmp/$name.ncu-rep gpurun_out/; fi
