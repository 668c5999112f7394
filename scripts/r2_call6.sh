#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c6_pytest.log; tail -12 gpurun_out/c6_pytest.log
bash scripts/ab.sh 2236 main
python scripts/ab_summary.py gpurun_out/ab_main.csv > gpurun_out/c6_ab.txt; cat gpurun_out/c6_ab.txt
timeout 300 python bench.py --steps 10 --skip-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; python -c "
import json;b=json.load(open('gpurun_out/c6_bench.json'));print(b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'])"; tail -3 gpurun_out/c6_bench.err
timeout 300 python bench.py --steps 5 --skip-cpu-baseline --workload shuffled > gpurun_out/c6_bench_shuf.json 2> gpurun_out/c6_bench_shuf.err; python -c "
import json;b=json.load(open('gpurun_out/c6_bench_shuf.json'));print('shuffled',b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'])"; tail -3 gpurun_out/c6_bench_shuf.err
