#!/bin/bash
# round 2, GPU visit 9 (8 GPUs): bench at N = 2, 4, 8 (lattice), N = 8 shuffled, timelines at N = 2 and 8
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
for n in 8 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n scripts/e2e_timeline_multi.py > gpurun_out/c9_timeline_n$n.txt 2> gpurun_out/c9_timeline_n$n.err
cat gpurun_out/c9_timeline_n$n.txt; tail -2 gpurun_out/c9_timeline_n$n.err
done
for n in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/c9_bench_n$n.json 2> gpurun_out/c9_bench_n$n.err
python -c "
import json;b=json.load(open('gpurun_out/c9_bench_n$n.json'));print($n, b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'],b.get('multi_gpu_check'))"; tail -2 gpurun_out/c9_bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --workload shuffled > gpurun_out/c9_bench_shuf_n8.json 2> gpurun_out/c9_bench_shuf_n8.err
python -c "
import json;b=json.load(open('gpurun_out/c9_bench_shuf_n8.json'));print('shuffled 8', b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'],b.get('multi_gpu_check'))"; tail -2 gpurun_out/c9_bench_shuf_n8.err
