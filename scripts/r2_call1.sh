#!/bin/bash
# round 2, first GPU visit: validate the XG_REC_PREFETCH build, A/B it, and capture the
# per-kernel DRAM bytes / fp64 pipe / issue metrics the measured rooflines are computed from
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/c1_smi.txt
XG_LIB_PATH=$PWD/xugrid_b200/_lib/libxgregrid_rec.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c1_pytest_rec.log 2>&1; echo "rc=$?" >> gpurun_out/c1_pytest_rec.log; tail -3 gpurun_out/c1_pytest_rec.log
bash scripts/ab.sh 2236 main rec
python scripts/ab_summary.py gpurun_out/ab_main.csv gpurun_out/ab_rec.csv > gpurun_out/c1_ab.txt; cat gpurun_out/c1_ab.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/c1_metrics.csv python scripts/profile_overlap.py 2236 2 > gpurun_out/c1_metrics.log 2>&1; tail -1 gpurun_out/c1_metrics.log
timeout 300 python bench.py --steps 5 --skip-cpu-baseline > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 1500 gpurun_out/c1_bench.json
