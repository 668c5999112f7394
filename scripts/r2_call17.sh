#!/bin/bash
# round 2, GPU visit 17: Voronoi kernels: tests, stage times of config 3
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_voronoi.py -m gpu -x -q > gpurun_out/c17_vor.log 2>&1; echo "rc=$?" >> gpurun_out/c17_vor.log; tail -40 gpurun_out/c17_vor.log
timeout 600 python -m pytest tests/test_gpu_regridder.py tests/test_gpu_reference_vectors.py -m gpu -x -q > gpurun_out/c17_regr.log 2>&1; echo "rc=$?" >> gpurun_out/c17_regr.log; tail -15 gpurun_out/c17_regr.log
timeout 600 python scripts/bary_timeline.py > gpurun_out/c17_bary.txt 2>&1; cat gpurun_out/c17_bary.txt
