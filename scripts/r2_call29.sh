#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_deferred.py tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c29_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c29_pytest.log; tail -4 gpurun_out/c29_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/e2e_timeline_multi.py > gpurun_out/c29_timeline_n2.txt 2> gpurun_out/c29_timeline_n2.err
grep -v "^===.*resident" gpurun_out/c29_timeline_n2.txt | head -12; tail -2 gpurun_out/c29_timeline_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c29_bench_n2.json 2> gpurun_out/c29_bench_n2.err
python -c "
import json;b=json.load(open('gpurun_out/c29_bench_n2.json'));print(2, b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b.get('multi_gpu_check'))"; tail -2 gpurun_out/c29_bench_n2.err
