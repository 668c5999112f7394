"""Summarise gpurun_out/ab_<variant>.csv files written by scripts/ab.sh (last repetition)."""
import csv, sys, collections
for path in sys.argv[1:]:
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r['ID'], {'name': r['Kernel Name'].split('(')[0][-40:]})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    ids = list(per)
    half = ids[len(ids) // 2:]
    print(path)
    tot = 0
    for i in half:
        d = per[i]
        t = d.get('gpu__time_duration.sum', 0) / 1e6
        tot += t
        print(f"  {t:8.3f} ms  inst {d.get('smsp__inst_executed.sum', 0)/1e6:9.1f} M  lanes {d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):5.1f}  issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  {d['name']}")
    print(f"  {tot:8.3f} ms total")
