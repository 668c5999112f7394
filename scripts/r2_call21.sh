#!/bin/bash
# round 2, GPU visit 21 (2 GPUs): fused apply + peer stores: multi-GPU test, bench N = 2 with / without
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_regridder.py -m gpu -x -q > gpurun_out/c21_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c21_pytest.log; tail -25 gpurun_out/c21_pytest.log
for mode in peers nccl; do
if [ $mode = nccl ]; then export XG_BENCH_NCCL_GATHER=1; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c21_bench_n2_$mode.json 2> gpurun_out/c21_bench_n2_$mode.err
python -c "
import json;b=json.load(open('gpurun_out/c21_bench_n2_$mode.json'));print('$mode', b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b['phase_ms'],b.get('multi_gpu_check'),b['detail']['output_assembly'])"; tail -3 gpurun_out/c21_bench_n2_$mode.err
done
