"""Fixed cost of an overlap job: time intersect_faces_csr over target ranges of growing size."""
import sys
import torch
import xugrid_b200 as xb
from xugrid_b200 import synthetic

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
dsxy, dsf = torch.from_numpy(sxy).cuda(), torch.from_numpy(sf).cuda().to(torch.int32)
dtxy, dtf = torch.from_numpy(txy).cuda(), torch.from_numpy(tf).cuda().to(torch.int32)
tree = xb.CellTree2d(dsxy, dsf)
n = tf.shape[0]
for frac in (0.01, 0.05, 0.1, 0.25, 0.5, 1.0):
    e = int(n * frac)
    for _ in range(3):
        tree.intersect_faces_csr(dtxy, dtf, t_begin=0, t_end=e)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tree.intersect_faces_csr(dtxy, dtf, t_begin=0, t_end=e)
    e1.record()
    torch.cuda.synchronize()
    print(f"targets {e:9d}  {e0.elapsed_time(e1) / 5:7.3f} ms  phases {tree.last_phase_ms}")
