#!/bin/bash
# round 2, GPU visit 2: full GPU suite on the new build (Hilbert order, mask-accumulating walk,
# compact 16-byte provisional rows, raster->raster path, reference vectors), A/B against the
# Morton-order and round-1-rows builds, bench line, per-kernel ncu metrics
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c2_pytest.log; tail -15 gpurun_out/c2_pytest.log
bash scripts/ab.sh 2236 main morton rows0
python scripts/ab_summary.py gpurun_out/ab_main.csv gpurun_out/ab_morton.csv gpurun_out/ab_rows0.csv > gpurun_out/c2_ab.txt; cat gpurun_out/c2_ab.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/c2_metrics.csv python scripts/profile_overlap.py 2236 2 > gpurun_out/c2_metrics.log 2>&1; tail -1 gpurun_out/c2_metrics.log
timeout 300 python bench.py --steps 5 --skip-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; tail -c 2500 gpurun_out/c2_bench.json; tail -5 gpurun_out/c2_bench.err
