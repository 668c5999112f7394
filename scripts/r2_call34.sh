#!/bin/bash
# round 2, GPU visit 34 (8 GPUs): final build at N = 8 (and 4): multi-GPU test, bench lines, timeline
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/c34_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c34_pytest.log; tail -4 gpurun_out/c34_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c34_bench_n8.json 2> gpurun_out/c34_bench_n8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/c34_bench_n4.json 2> gpurun_out/c34_bench_n4.err
timeout 300 $TR --nproc-per-node 8 --master-port 29509 bench.py --gpus 8 --steps 10 --warmup 3 --workload shuffled > gpurun_out/c34_bench_shuf_n8.json 2> gpurun_out/c34_bench_shuf_n8.err
for f in c34_bench_n8 c34_bench_n4 c34_bench_shuf_n8; do
python -c "
import json;b=json.load(open('gpurun_out/$f.json'));print('$f', b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['e2e']['h2d_bytes_per_step'],b.get('multi_gpu_check'))"; tail -2 gpurun_out/$f.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29518 scripts/e2e_timeline_multi.py > gpurun_out/c34_timeline_n8.txt 2> gpurun_out/c34_timeline_n8.err
grep -v "^===.*resident" gpurun_out/c34_timeline_n8.txt | head -12; tail -2 gpurun_out/c34_timeline_n8.err
