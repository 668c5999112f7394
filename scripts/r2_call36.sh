#!/bin/bash
# ncu launch list (time + DRAM bytes) of the BarycentricInterpolator build at config-3 size
export PATH=/usr/local/cuda/bin:$PATH
export PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 60 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c36_bary_launches.csv python scripts/bary_timeline.py > gpurun_out/c36.log 2>&1
tail -3 gpurun_out/c36.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/c36_bary_launches.csv') if not l.startswith('=='))]
per=collections.OrderedDict()
for r in rows:
    v=float(r['Metric Value'].replace(',',''))
    u=r['Metric Unit']
    sc={'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1.0,'ns':1e-6,'us':1e-3,'ms':1.0}.get(u,1.0)
    per.setdefault(r['ID'],{'name':r['Kernel Name']})[r['Metric Name']]=v*sc
ids=list(per); half=ids[len(ids)//2:]
tot=0
for i in half:
    d=per[i]; ms=d.get('gpu__time_duration.sum',0); b=d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0); tot+=ms
    if ms>0.05: print(f"{ms:8.3f} ms {b/1e9:7.3f} GB {b/1e9/(ms*1e-3)/1e3 if ms else 0:6.2f} TB/s  {d['name'][:70]}")
print('total', tot, 'launches', len(half))
PY
