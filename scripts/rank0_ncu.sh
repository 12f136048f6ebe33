#!/bin/bash
# torchrun --no-python wrapper: rank 0 runs under ncu (launch list only: one pass per kernel, no replay), the others free
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-300} --csv --log-file "$NCU_LOG" python "$@"
else
  exec python "$@"
fi
