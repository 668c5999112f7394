#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
export XG_GRAPH_DEBUG=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_deferred.py -m gpu -x -q -s > gpurun_out/c30_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c30_pytest.log; grep "\[xg\]" gpurun_out/c30_pytest.log | sort | uniq -c | head; tail -5 gpurun_out/c30_pytest.log
unset XG_GRAPH_DEBUG
timeout 600 python scripts/e2e_timeline.py 2236 "0.2,0.85,1;0.15,0.5,0.85,1" > gpurun_out/c30_timeline.txt 2>&1
grep "chunks=\|weights done\|regrid done\|job\|Error\|error" gpurun_out/c30_timeline.txt | head -30
