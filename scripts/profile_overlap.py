"""One overlap build + apply on synthetic jittered-lattice triangles (for ncu)."""
import sys
import torch
import xugrid_b200 as xb
from xugrid_b200 import sparse, synthetic

mode = sys.argv[1] if len(sys.argv) > 1 else "2236"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if mode == "cfg4":    # mixed quad/tri source -> tri target
    sxy, sf = synthetic.jittered_lattice(2582, 2582, 30, "mixed")
    txy, tf = synthetic.jittered_lattice(2236, 2236, 32, "tri")
elif mode == "cfg5":  # 50M-face source, 2M-face Delaunay target (44 entries per row)
    sxy, sf = synthetic.jittered_lattice(5000, 5000, 40, "tri")
    txy, tf = synthetic.delaunay(1_000_000, 41)
else:
    n_side = int(mode)
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
