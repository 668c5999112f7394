#!/bin/bash
# round 2, GPU visit 28 (2 GPUs): split ranges + drain for shards: tests, bench N = 1 / 2, timeline N = 2
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c28_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c28_pytest.log; tail -15 gpurun_out/c28_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/e2e_timeline_multi.py > gpurun_out/c28_timeline_n2.txt 2> gpurun_out/c28_timeline_n2.err
grep -v "^===.*resident" gpurun_out/c28_timeline_n2.txt | head -14; tail -2 gpurun_out/c28_timeline_n2.err
timeout 300 python bench.py --steps 10 --skip-cpu-baseline > gpurun_out/c28_bench_n1.json 2> gpurun_out/c28_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c28_bench_n2.json 2> gpurun_out/c28_bench_n2.err
for n in 1 2; do
python -c "
import json;b=json.load(open('gpurun_out/c28_bench_n$n.json'));print($n, b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b.get('multi_gpu_check'))"; tail -2 gpurun_out/c28_bench_n$n.err
done
