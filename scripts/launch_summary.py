"""Summarise an ncu launch list CSV (gpu__time_duration.sum): last repetition only."""
import csv, sys
path = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lines = [l for l in open(path) if not l.startswith('==')]
rows = [(r['Kernel Name'].split('(')[0][-60:], float(r['Metric Value'].replace(',', '')), r['Metric Unit']) for r in csv.DictReader(lines)]
n = len(rows) // reps
tot = 0.0
for name, v, u in rows[-n:]:
    ms = v / 1e6 if u.startswith('ns') else (v / 1e3 if u.startswith('us') else v)
    tot += ms
    print(f"{ms:9.3f} ms  {name}")
print(f"{tot:9.3f} ms  TOTAL (cold-cache, serialised)")
