"""One overlap build + apply on synthetic jittered-lattice triangles (for ncu)."""
import sys
import torch
import xugrid_b200 as xb
from xugrid_b200 import sparse, synthetic

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
dsxy, dsf = torch.from_numpy(sxy).cuda(), torch.from_numpy(sf).cuda()
dtxy, dtf = torch.from_numpy(txy).cuda(), torch.from_numpy(tf).cuda()
vals = torch.rand(1, sf.shape[0], dtype=torch.float64, device="cuda")
for it in range(reps):
    tree = xb.CellTree2d(dsxy, dsf)
    indptr, indices, area = tree.intersect_faces_csr(dtxy, dtf)
    A = sparse.MatrixCSR(area, indices, indptr, tf.shape[0], sf.shape[0], int(area.numel()))
    out = sparse.csr_apply(0, 0.0, vals, A, A.n)
    torch.cuda.synchronize()
print("nnz", area.numel(), tree.last_stats)
