"""Top stall-sample instructions of one kernel of an .ncu-rep (SASS view with stall reasons).
    python scripts/src_hot.py <rep> <kernel regex> [n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
# first row: kernel name; second: header
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print(f"kernel {rows[0][1][:80]}  samples {tot}  inst {tot_inst}")
stall_cols = [h for h in hdr if h.startswith("stall_")]
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = body[i]
    s = int(r[ix["# Samples"]] or 0)
    reasons = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    rs = " ".join(f"{c[6:]}={v}" for v, c in reasons if v)
    print(f"{i:5d} {100.0 * s / tot:5.1f}%  {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} {rs}")
