"""Per-phase device time of deferred overlap jobs of different sizes on resident data."""
import sys
import torch
import xugrid_b200 as xb
from xugrid_b200 import synthetic

n_side = 2236
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
dev = torch.device("cuda", 0)
dsxy, dsf = torch.from_numpy(sxy).to(dev), torch.from_numpy(sf).to(dev).to(torch.int32)
dtxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf).to(dev).to(torch.int32)
tree = xb.CellTree2d(dsxy, dsf)
n_t = tf.shape[0]
indptr = torch.empty(n_t + 1, dtype=torch.int64, device=dev)
indices = torch.empty(16 * n_t, dtype=torch.int32, device=dev)
area = torch.empty(16 * n_t, dtype=torch.float64, device=dev)
for label, ranges in (("1M", [(0, 1000385)]), ("3 chained", [(0, 1000385), (1000385, 4001541), (4001541, n_t)]),
                      ("whole", [(0, n_t)])):
    for rep in range(4):
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        jobs = [tree.intersect_faces_into(dtxy, dtf, indptr[b: e + 1], indices, area, t_begin=b, t_end=e, chained=c > 0)
                for c, (b, e) in enumerate(ranges)]
        ev1.record()
        torch.cuda.synchronize()
        res = [j.result() for j in jobs]
    print(f"--- {label}: total {ev0.elapsed_time(ev1):.3f} ms")
    for j in jobs:
        print("   " + "  ".join(f"{k} {v:.3f}" for k, v in j.phase_ms.items()) + f"  sum {sum(j.phase_ms.values()):.3f}")
