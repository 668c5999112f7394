#!/bin/bash
# xg_locate_any for the outside mask: full suite + stage times of the config-3 build
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/c38_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c38_pytest.log; tail -12 gpurun_out/c38_pytest.log
timeout 60 python scripts/bary_timeline.py > gpurun_out/c38_bary.txt 2>&1; cat gpurun_out/c38_bary.txt
