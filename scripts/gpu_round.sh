#!/bin/bash
# One GPU-box visit.  Usage: gpu_round.sh [tests] [bench] [ref] [launches] [ncu]
set -x
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
what="${*:-tests bench ref launches ncu}"
for w in $what; do case $w in
tests) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log;;
bench) timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json;;
ref) timeout 300 python bench.py --impl reference --steps 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json;;
launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_overlap.py 2236 2 > gpurun_out/launches.log 2>&1; tail -2 gpurun_out/launches.log;;
ncu) timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_traverse|k_sat|k_clip_survivors|k_counts_to_rows|k_move|k_pack|k_face_keys|k_csr_apply' -s 12 -c 12 -f -o gpurun_out/prof_main python scripts/profile_overlap.py 2236 2 > gpurun_out/prof_main.log 2>&1; tail -2 gpurun_out/prof_main.log;;
esac; done
ls -la gpurun_out
