"""A few end-to-end steps (public API, host buffers) for an ncu launch list."""
import sys
import numpy as np
import torch
import xugrid_b200 as xb
from xugrid_b200 import synthetic

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
values = torch.from_numpy(np.random.default_rng(12).random((1, sf.shape[0]))).pin_memory()
src = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
tgt = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
for rep in range(reps):
    src.release_device(); tgt.release_device()
    r = xb.OverlapRegridder(src, tgt, method="mean")
    out = r.regrid(values)
    torch.cuda.synchronize()
print("done", float(out[0, 0]))
