#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python scripts/e2e_timeline.py 2236 "0.2,0.85,1;0.1,0.8,1;0.1,0.5,0.88,1;0.08,0.4,0.85,1;0.15,0.88,1;0.12,0.55,0.9,1" > gpurun_out/c32_timeline.txt 2>&1
grep "chunks=\|weights done\|regrid done" gpurun_out/c32_timeline.txt | head -30
