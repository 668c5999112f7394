#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python scripts/e2e_timeline.py 2236 "0.1,0.4,1;0.3,1;0.2,1;0.15,0.55,1;0.2,0.6,1;1" > gpurun_out/c11_timeline.txt 2>&1
grep -v "copy stream" gpurun_out/c11_timeline.txt
