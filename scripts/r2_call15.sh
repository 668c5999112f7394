#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 300 python scripts/job_overhead.py > gpurun_out/c15_jobs.txt 2>&1; cat gpurun_out/c15_jobs.txt
