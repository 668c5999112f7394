import time, torch
import xugrid_b200 as xb
from xugrid_b200 import synthetic
from xugrid_b200.distributed import target_range_box, shard_range
n=2236
sxy, sf = synthetic.jittered_lattice(n, n, 10, "tri")
txy, tf = synthetic.jittered_lattice(n + 3, n - 2, 11, "tri")
d = [torch.from_numpy(a).cuda() for a in (sxy, sf, txy, tf)]
b, e = shard_range(tf.shape[0], 0, 2)
for it in range(6):
    torch.cuda.synchronize(); t0=time.perf_counter()
    box = target_range_box(d[2], d[3], b, e)
    torch.cuda.synchronize(); t1=time.perf_counter()
    tree = xb.CellTree2d(d[0], d[1], clip_box=box)
    torch.cuda.synchronize(); t2=time.perf_counter()
    ip, ix, ar = tree.intersect_faces_csr(d[2], d[3], t_begin=b, t_end=e)
    torch.cuda.synchronize(); t3=time.perf_counter()
    print(f"box {1e3*(t1-t0):.2f} index {1e3*(t2-t1):.2f} (indexed {tree.n_indexed}) overlap {1e3*(t3-t2):.2f}", tree.last_phase_ms)
