#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
V="main c192x3 c128x5 c128x6 c256x3 sat4"
bash scripts/ab.sh 2236 $V
python scripts/ab_summary.py $(for v in $V; do echo gpurun_out/ab_$v.csv; done) > gpurun_out/c16_ab.txt; grep -v "k_move\|k_counts\|k_traverse" gpurun_out/c16_ab.txt
