#!/bin/bash
# round 2, GPU visit 19 (2 GPUs): full suite, configs 4/5 sharded (small), config 3
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c19_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c19_pytest.log; tail -4 gpurun_out/c19_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/bench_configs_multi.py 4 5 --small > gpurun_out/c19_cfg_small.jsonl 2> gpurun_out/c19_cfg_small.err; cat gpurun_out/c19_cfg_small.jsonl; tail -5 gpurun_out/c19_cfg_small.err
timeout 600 python scripts/bench_configs.py 2 3 > gpurun_out/c19_cfg23.jsonl 2> gpurun_out/c19_cfg23.err; cat gpurun_out/c19_cfg23.jsonl | cut -c1-700; tail -3 gpurun_out/c19_cfg23.err
