#!/bin/bash
# round 2, last GPU visit: the full GPU suite and smoke() on the final tree
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/c35_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c35_pytest.log; tail -4 gpurun_out/c35_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
