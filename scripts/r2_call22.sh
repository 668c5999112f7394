#!/bin/bash
# round 2, GPU visit 22 (8 GPUs): bench N = 8 (lattice, shuffled), end-to-end timeline, configurations 4 and 5 as named
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29508 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c22_bench_n8.json 2> gpurun_out/c22_bench_n8.err
timeout 300 $TR --master-port 29509 bench.py --gpus 8 --steps 10 --warmup 3 --workload shuffled > gpurun_out/c22_bench_shuf_n8.json 2> gpurun_out/c22_bench_shuf_n8.err
for f in c22_bench_n8 c22_bench_shuf_n8; do
python -c "
import json;b=json.load(open('gpurun_out/$f.json'));print('$f', b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'],b.get('multi_gpu_check'),b['detail'])"; tail -2 gpurun_out/$f.err
done
timeout 300 $TR --master-port 29518 scripts/e2e_timeline_multi.py > gpurun_out/c22_timeline_n8.txt 2> gpurun_out/c22_timeline_n8.err
cat gpurun_out/c22_timeline_n8.txt; tail -2 gpurun_out/c22_timeline_n8.err
timeout 900 $TR --master-port 29533 scripts/bench_configs_multi.py 4 5 > gpurun_out/c22_cfg45_n8.jsonl 2> gpurun_out/c22_cfg45_n8.err; cat gpurun_out/c22_cfg45_n8.jsonl; tail -4 gpurun_out/c22_cfg45_n8.err
