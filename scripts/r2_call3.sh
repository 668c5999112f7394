#!/bin/bash
# round 2, GPU visit 3: next-node prefetch in the walk (A/B), bit-parallel Hilbert key, device-side
# clip box / lazy index finalize / stream-ordered destroy; full suite + bench
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c3_pytest.log; tail -5 gpurun_out/c3_pytest.log
bash scripts/ab.sh 2236 main nopf
python scripts/ab_summary.py gpurun_out/ab_main.csv gpurun_out/ab_nopf.csv > gpurun_out/c3_ab.txt; cat gpurun_out/c3_ab.txt
timeout 300 python bench.py --steps 10 --skip-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; python -c "
import json;b=json.load(open('gpurun_out/c3_bench.json'));print(b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'])"; tail -3 gpurun_out/c3_bench.err
