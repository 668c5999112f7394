"""One rank's resident step of an N-rank job on ONE GPU (for an ncu launch list): rank 0's
target range of the 10 M -> 10 M workload, culled index, overlap job, apply."""
import sys
import torch
import xugrid_b200 as xb
from xugrid_b200 import sparse, synthetic
from xugrid_b200.distributed import shard_range

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_side = 2236
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
dev = torch.device("cuda", 0)
dsxy, dsf = torch.from_numpy(sxy).to(dev), torch.from_numpy(sf).to(dev)
dtxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf).to(dev)
vals = torch.rand(1, sf.shape[0], dtype=torch.float64, device=dev)
b, e = shard_range(tf.shape[0], 0, world)
rows = e - b
for it in range(reps):
    torch.cuda.synchronize()
    ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ev0.record()
    tree = xb.CellTree2d(dsxy, dsf, device=dev, clip_targets=(dtxy, dtf, b, e))
    ev1.record()
    indptr = torch.empty(rows + 1, dtype=torch.int64, device=dev)
    indices = torch.empty(16 * rows, dtype=torch.int32, device=dev)
    area = torch.empty(16 * rows, dtype=torch.float64, device=dev)
    job = tree.intersect_faces_into(dtxy, dtf, indptr, indices, area, t_begin=b, t_end=e)
    A = sparse.MatrixCSR(area, indices, indptr, rows, sf.shape[0], -1)
    out = sparse.csr_apply(0, 0.0, vals, A, rows)
    ev2.record()
    torch.cuda.synchronize()
print("rank 0 of", world, "rows", rows, "indexed", tree.n_indexed, "index ms", ev0.elapsed_time(ev1), "job+apply ms", ev1.elapsed_time(ev2),
      job.result(), job.phase_ms)
