#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py 3 > gpurun_out/c24_cfg3.jsonl 2> gpurun_out/c24_cfg3.err; cut -c1-900 gpurun_out/c24_cfg3.jsonl; tail -3 gpurun_out/c24_cfg3.err
