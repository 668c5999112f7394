#!/bin/bash
# k_locate with per-lane hit masks: parity + stage times of the config-3 build
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_voronoi.py tests/test_gpu_reference_vectors.py tests/test_gpu_regridder.py -m gpu -x -q > gpurun_out/c37_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c37_pytest.log; tail -4 gpurun_out/c37_pytest.log
timeout 100 python scripts/bary_timeline.py > gpurun_out/c37_bary.txt 2>&1; cat gpurun_out/c37_bary.txt
