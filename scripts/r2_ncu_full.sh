#!/bin/bash
# ncu --set full (+ source) of the three big kernels, second repetition
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_traverse|k_sat|k_clip_survivors' -s 3 -c 3 -f -o gpurun_out/${1:-r2_full} python scripts/profile_overlap.py 2236 2 > gpurun_out/${1:-r2_full}.log 2>&1
tail -2 gpurun_out/${1:-r2_full}.log
