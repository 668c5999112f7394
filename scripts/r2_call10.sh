#!/bin/bash
# round 2, GPU visit 10: deferred overlap job (xg_overlap_run): tests, timeline, bench
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_deferred.py -m gpu -x -q > gpurun_out/c10_deferred.log 2>&1; echo "rc=$?" >> gpurun_out/c10_deferred.log; tail -30 gpurun_out/c10_deferred.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c10_pytest.log; tail -6 gpurun_out/c10_pytest.log
timeout 600 python scripts/e2e_timeline.py 2236 "0.1,0.4,1;4;8" > gpurun_out/c10_timeline.txt 2>&1
cat gpurun_out/c10_timeline.txt
timeout 300 python bench.py --steps 10 --skip-cpu-baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; python -c "
import json;b=json.load(open('gpurun_out/c10_bench.json'));print(b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'])"; tail -3 gpurun_out/c10_bench.err
