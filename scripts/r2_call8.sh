#!/bin/bash
# round 2, GPU visit 8: end-to-end timeline at N = 1 (copy-stream milestones included)
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python scripts/e2e_timeline.py 2236 "0.1,0.4,1;4;0.05,0.2,0.45,0.7,1;8" > gpurun_out/c8_timeline.txt 2>&1
cat gpurun_out/c8_timeline.txt
