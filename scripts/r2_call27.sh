#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python scripts/e2e_timeline.py 2236 "0.15,0.5,0.85,1;0.2,0.85,1;0.15,0.8,1;0.25,0.9,1;0.1,0.9,1;0.2,0.6,0.9,1" > gpurun_out/c27_timeline.txt 2>&1
grep "chunks=\|weights done\|regrid done\|job" gpurun_out/c27_timeline.txt
