#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_regridder.py -m gpu -x -q > gpurun_out/c20_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c20_pytest.log; tail -4 gpurun_out/c20_pytest.log
timeout 300 python scripts/profile_shard.py 8 4 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c20_shard_launches.csv python scripts/profile_shard.py 8 3 > gpurun_out/c20.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/c20_shard_launches.csv') if not l.startswith('=='))]
n=len(rows)//3
tot=0
for r in rows[-n:]:
    v=float(r['Metric Value'].replace(',',''))
    ms=v*{'ns':1e-6,'us':1e-3,'ms':1.0}.get(r['Metric Unit'],1.0)
    tot+=ms
    print(f"{ms:8.4f}  {r['Kernel Name'][:80]}")
print("total", tot, "launches", n)
PY
