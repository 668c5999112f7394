#!/bin/bash
# ncu launch list (durations) of scripts/profile_overlap.py <mode>
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
for m in "$@"; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$m.csv python scripts/profile_overlap.py $m 2 > gpurun_out/launches_$m.log 2>&1
tail -1 gpurun_out/launches_$m.log
done
