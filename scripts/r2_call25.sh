#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_deferred.py tests/test_gpu_regridder.py -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c25_pytest.log; tail -25 gpurun_out/c25_pytest.log
timeout 600 python scripts/e2e_timeline.py 2236 "0.1,0.4,1;0.1,0.4,0.8,1;0.1,0.35,0.65,0.9,1;0.15,0.5,0.85,1" > gpurun_out/c25_timeline.txt 2>&1
grep -v "copy stream\|values on device\|apply done" gpurun_out/c25_timeline.txt
