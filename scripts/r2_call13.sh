#!/bin/bash
# round 2, GPU visit 13 (8 GPUs): deferred jobs + shared host result: timelines at N = 8, 2, 1; bench at N = 1, 2, 4, 8
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_deferred.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest.log; tail -4 gpurun_out/c13_pytest.log
for n in 8 2 1; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n scripts/e2e_timeline_multi.py > gpurun_out/c13_timeline_n$n.txt 2> gpurun_out/c13_timeline_n$n.err
cat gpurun_out/c13_timeline_n$n.txt; tail -2 gpurun_out/c13_timeline_n$n.err
done
timeout 300 python bench.py --steps 10 --skip-cpu-baseline > gpurun_out/c13_bench_n1.json 2> gpurun_out/c13_bench_n1.err
for n in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/c13_bench_n$n.json 2> gpurun_out/c13_bench_n$n.err
done
for n in 1 2 4 8; do
python -c "
import json;b=json.load(open('gpurun_out/c13_bench_n$n.json'));print($n, b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'],b.get('multi_gpu_check'))"; tail -2 gpurun_out/c13_bench_n$n.err
done
