"""Per-step wall/device times of the resident step (diagnostics)."""
import sys, time
import torch
import xugrid_b200 as xb
from xugrid_b200 import sparse, synthetic

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
d = [torch.from_numpy(a).cuda() for a in (sxy, sf, txy, tf)]
vals = torch.rand(1, sf.shape[0], dtype=torch.float64, device="cuda")
tree = None
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tree = xb.CellTree2d(d[0], d[1])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    indptr, indices, area = tree.intersect_faces_csr(d[2], d[3])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    A = sparse.MatrixCSR(area, indices, indptr, tf.shape[0], sf.shape[0], int(area.numel()))
    out = sparse.csr_apply(0, 0.0, vals, A, A.n)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"step {it}: index {1e3*(t1-t0):7.2f} overlap {1e3*(t2-t1):7.2f} apply {1e3*(t3-t2):6.2f} ms | phases "
          + " ".join(f"{k}={v:.2f}" for k, v in tree.last_phase_ms.items()), flush=True)
