#!/bin/bash
# ncu --set full of ONE kernel (regex $1) of the second repetition; report -> gpurun_out/$2.ncu-rep
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s 1 -c 1 -f -o gpurun_out/$2 python scripts/profile_overlap.py ${3:-2236} 2 > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log
