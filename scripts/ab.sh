#!/bin/bash
# A/B of tuning builds: ncu launch list (time + executed instructions) per library variant.
# usage: ab.sh <n_side> <variant> [<variant> ...]   ("main" = the default library)
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
n=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then unset XG_LIB_PATH; else export XG_LIB_PATH=$PWD/xugrid_b200/_lib/libxgregrid_$v.so; fi
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none \
     -k 'regex:k_traverse|k_sat|k_clip|k_move|k_counts' --csv --log-file gpurun_out/ab_$v.csv python scripts/profile_overlap.py $n 2 > gpurun_out/ab_$v.log 2>&1
  tail -1 gpurun_out/ab_$v.log
done
