#!/bin/bash
# round 2, GPU visit 5: reserved arena windows (persistent walk, k_sat), fp32 SAT pre-decision,
# late slot atomic in the clip kernel: full suite + A/B of each change + bench
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c5_pytest.log; tail -12 gpurun_out/c5_pytest.log
bash scripts/ab.sh 2236 main sat64 atom0 win256
python scripts/ab_summary.py gpurun_out/ab_main.csv gpurun_out/ab_sat64.csv gpurun_out/ab_atom0.csv gpurun_out/ab_win256.csv > gpurun_out/c5_ab.txt; cat gpurun_out/c5_ab.txt
timeout 300 python bench.py --steps 10 --skip-cpu-baseline > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; python -c "
import json;b=json.load(open('gpurun_out/c5_bench.json'));print(b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'])"; tail -3 gpurun_out/c5_bench.err
