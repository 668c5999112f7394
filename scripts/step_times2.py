"""Resident step without intermediate syncs (mimics bench.py) - diagnostics."""
import sys, time
import torch
import xugrid_b200 as xb
from xugrid_b200 import sparse, synthetic

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
d = [torch.from_numpy(a).cuda() for a in (sxy, sf, txy, tf)]
vals = torch.rand(1, sf.shape[0], dtype=torch.float64, device="cuda")
state = {}
def step():
    tree = xb.CellTree2d(d[0], d[1])
    indptr, indices, area = tree.intersect_faces_csr(d[2], d[3], t_begin=0, t_end=tf.shape[0])
    A = sparse.MatrixCSR(area, indices, indptr, tf.shape[0], sf.shape[0], int(area.numel()))
    out = sparse.csr_apply(0, 0.0, vals, A, A.n)
    if mode == "keep":
        state["tree"], state["A"] = tree, A
    return out, tree
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, tree = step()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"{mode} step {it}: {1e3*(t3-t0):7.2f} ms | " + " ".join(f"{k}={v:.2f}" for k, v in tree.last_phase_ms.items()), flush=True)
    del out, tree

# no host sync between steps (like bench.py)
for rep in range(3):
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for it in range(5):
        out, tree = step()
    ev1.record()
    torch.cuda.synchronize()
    print(f"{mode} nosync rep {rep}: {ev0.elapsed_time(ev1)/5:7.2f} ms/step | " + " ".join(f"{k}={v:.2f}" for k, v in tree.last_phase_ms.items()), flush=True)
