#!/bin/bash
# round 2, GPU visit 33 (2 GPUs): 24-bit connectivity transport: tests, timeline, bench N = 1 / 2
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c33_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c33_pytest.log; tail -15 gpurun_out/c33_pytest.log
timeout 600 python scripts/e2e_timeline.py 2236 "0.2,0.85,1" > gpurun_out/c33_timeline.txt 2>&1
grep -v "job" gpurun_out/c33_timeline.txt | head -20
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/c33_bench_n1.json 2> gpurun_out/c33_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c33_bench_n2.json 2> gpurun_out/c33_bench_n2.err
for n in 1 2; do
python -c "
import json;b=json.load(open('gpurun_out/c33_bench_n$n.json'));print($n, b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['e2e']['h2d_bytes_per_step'],b['e2e']['one_off_staging_ms'],b.get('multi_gpu_check'))"; tail -2 gpurun_out/c33_bench_n$n.err
done
