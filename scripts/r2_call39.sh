#!/bin/bash
# last GPU visit of the round: the full GPU suite and smoke() on the final tree
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/c39_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c39_pytest.log; tail -6 gpurun_out/c39_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
