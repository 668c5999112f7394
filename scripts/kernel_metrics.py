"""ncu metrics CSV (scripts/r2_metrics.sh: gpu__time_duration, dram__bytes_read/write, fp64 pipe,
issue slots, executed instructions per launch) -> profiles/r2_kernel_metrics.json, the per-kernel
numbers bench.py puts next to its live timings.  Last repetition of the run only.

    python scripts/kernel_metrics.py gpurun_out/metrics.csv 2236 [reps] > profiles/r2_kernel_metrics.json
"""
import collections
import csv
import json
import sys

path, n_side = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
             "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(u, 1.0)
    per.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = v * scale
ids = list(per)
last = ids[len(ids) - len(ids) // reps:]
short = {"k_traverse": "k_traverse", "k_sat": "k_sat", "k_clip_survivors": "k_clip_survivors",
         "k_move_rows": "k_move_rows", "k_pack": "k_pack", "k_face_keys": "k_face_keys", "k_csr_apply": "k_csr_apply",
         "k_reduce_level": "k_reduce_level", "k_counts_to_rows": "k_counts_to_rows", "DeviceRadixSort": "cub_radix_sort",
         "DeviceScan": "cub_scan"}
kernels = collections.OrderedDict()
total_ms = 0.0
for i in last:
    d = per[i]
    key = next((v for k, v in short.items() if k in d["name"]), None)
    ms = d.get("gpu__time_duration.sum", 0.0)
    total_ms += ms
    if key is None:
        continue
    k = kernels.setdefault(key, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0,
                                 "inst_executed": 0.0, "_fp64": 0.0, "_issue": 0.0})
    k["launches"] += 1
    k["ms"] += ms
    k["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
    k["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
    k["dram_bytes"] = k["dram_read"] + k["dram_write"]
    k["inst_executed"] += d.get("smsp__inst_executed.sum", 0.0)
    k["_fp64"] += ms * d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0.0)
    k["_issue"] += ms * d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
for k in kernels.values():
    k["fp64_pipe_pct"] = k.pop("_fp64") / k["ms"] if k["ms"] else 0.0
    k["issue_pct"] = k.pop("_issue") / k["ms"] if k["ms"] else 0.0
    k["share_of_step"] = k["ms"] / total_ms
json.dump({"n_side": n_side, "step_ms_under_ncu": total_ms,
           "source": f"{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                     "sm__inst_executed_pipe_fp64...,smsp__issue_active...,smsp__inst_executed.sum "
                     "--clock-control none over scripts/profile_overlap.py (last repetition); times are "
                     "cold-cache and serialised: use the shares, and bench.py's live times for absolutes",
           "kernels": kernels}, sys.stdout, indent=1)
