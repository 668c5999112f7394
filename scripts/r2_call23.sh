#!/bin/bash
# round 2, GPU visit 23: full suite, smoke, configs 2/3 with oracle sample checks, bench line
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c23_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c23_pytest.log; tail -4 gpurun_out/c23_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python scripts/bench_configs.py 2 3 > gpurun_out/c23_cfg23.jsonl 2> gpurun_out/c23_cfg23.err; cut -c1-900 gpurun_out/c23_cfg23.jsonl; tail -3 gpurun_out/c23_cfg23.err
timeout 600 python bench.py > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err; tail -2 gpurun_out/c23_bench.err
python -c "
import json;b=json.load(open('gpurun_out/c23_bench.json'));print(b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'],b['cpu_baseline'], b['gpu_launches'], b['clocks'])"
