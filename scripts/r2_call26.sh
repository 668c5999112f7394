#!/bin/bash
# round 2, GPU visit 26: final build: full suite, ncu metric list, bench line
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c26_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c26_pytest.log; tail -4 gpurun_out/c26_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
   --clock-control none --csv --log-file gpurun_out/c26_metrics.csv python scripts/profile_overlap.py 2236 2 > gpurun_out/c26_metrics.log 2>&1
tail -1 gpurun_out/c26_metrics.log
python scripts/kernel_metrics.py gpurun_out/c26_metrics.csv 2236 2 > gpurun_out/c26_kernel_metrics.json
python -c "
import json;d=json.load(open('gpurun_out/c26_kernel_metrics.json'));print(d['step_ms_under_ncu']);[print(k,round(v['ms'],3),round(v['dram_bytes']/1e9,2),round(v['fp64_pipe_pct'],1),round(v['issue_pct'],1),round(v['inst_executed']/1e6,1)) for k,v in d['kernels'].items()]"
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err; tail -2 gpurun_out/c26_bench.err
python -c "
import json;b=json.load(open('gpurun_out/c26_bench.json'));print(b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'], b['gpu_launches'], b['roofline']['frac'], b['roofline']['frac_measured_dram'])"
