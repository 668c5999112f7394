#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c12_e2e_launches.csv python scripts/profile_e2e.py 2236 3 > gpurun_out/c12.log 2>&1
tail -2 gpurun_out/c12.log
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/c12_e2e_launches.csv') if not l.startswith('=='))]
n=len(rows)//3
last=rows[-n:]
for r in last:
    v=float(r['Metric Value'].replace(',',''))
    u=r['Metric Unit']
    ms=v*{'ns':1e-6,'us':1e-3,'ms':1.0,'nsecond':1e-6,'usecond':1e-3,'msecond':1.0}.get(u,1.0)
    print(f"{ms:8.4f}  {r['Kernel Name'][:90]}")
PY
