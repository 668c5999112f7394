"""End-to-end step at N ranks (torchrun): CUDA-event timestamps on rank 0's compute stream at
the milestones of bench.py's e2e_step (lattice workload: target_range shards, sharded upload),
plus the copy-stream milestones of the grids.  Also the resident step split in its calls.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/e2e_timeline_multi.py [n_side]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

import xugrid_b200 as xb
from xugrid_b200 import celltree, sparse, synthetic
from xugrid_b200 import distributed as xd

rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
n_t = tf.shape[0]
if world > 1:
    bound = None if os.environ.get("XG_BENCH_NO_BIND") else xd.bind_to_gpu_cpus(local_rank)
    dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        print("rank 0 bound to", None if bound is None else len(bound), "cores local to its GPU")
values = torch.from_numpy(np.random.default_rng(12).random((1, sf.shape[0]))).pin_memory()
h_out = xd.shared_host_array((1, n_t), torch.float64)
src = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf, device=dev)
tgt = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf, device=dev)
for g in (src, tgt):
    g._stage()
b, e = xd.shard_range(n_t, rank, world)

marks = []
def mark(name):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    marks.append((name, ev, time.perf_counter()))

def wrap(mod, fname, label):
    orig = getattr(mod, fname)
    def f(*a, **k):
        r = orig(*a, **k)
        mark(label)
        return r
    setattr(mod, fname, f)
    return orig

orig_init = celltree.CellTree2d.__init__
def init(self, *a, **k):
    orig_init(self, *a, **k); mark("index built")
celltree.CellTree2d.__init__ = init
orig_csr = celltree.CellTree2d.intersect_faces_csr
def csr(self, *a, **k):
    r = orig_csr(self, *a, **k); mark("overlap rows done"); return r
celltree.CellTree2d.intersect_faces_csr = csr
wrap(sparse, "csr_apply", "apply done")

def e2e_step():
    src.release_device(); tgt.release_device()
    mark("start")
    if world == 1:
        r = xb.OverlapRegridder(src, tgt, method="mean")
        mark("weights done")
        out = r.regrid(values)
        mark("regrid done (after D2H sync)")
        return
    r = xb.OverlapRegridder(src, tgt, method="mean", target_range=(b, e), sharded_upload=True)
    mark("weights enqueued")
    r.regrid(values, out=h_out[:, b:e])
    mark("own columns on the host (D2H done)")

for rep in range(5):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    marks.clear()
    t0 = time.perf_counter()
    e2e_step()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
if rank == 0:
    print(f"=== N={world} end to end, wall {wall:.2f} ms (rank 0)")
    rows = [(marks[0][1].elapsed_time(ev), f"{name}   [host {1e3 * (th - marks[0][2]):.2f}]") for name, ev, th in marks[1:]]
    for label, g in (("source", src), ("target", tgt)):
        rows += [(marks[0][1].elapsed_time(ev), f"[copy stream] {label} {name}") for name, ev in getattr(g, "_upload_events", [])]
    for t, name in sorted(rows):
        print(f"  {t:8.3f} ms  {name}")

# ---- resident step, per call
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(sxy=sxy, sf=sf, txy=txy, tf=tf).items()}
dv = values.to(dev)
def resident():
    mark("start")
    clip = (d["txy"], d["tf"], b, e) if world > 1 else None
    tree = xb.CellTree2d(d["sxy"], d["sf"], device=dev, clip_targets=clip)
    indptr, indices, area = tree.intersect_faces_csr(d["txy"], d["tf"], t_begin=b, t_end=e)
    A = sparse.MatrixCSR(area, indices, indptr, e - b, sf.shape[0], int(area.numel()))
    local = sparse.csr_apply(0, 0.0, dv, A, e - b)
    full = xd.all_gather_columns(local, n_t) if world > 1 else local
    mark("all-gather done")
    return tree
for rep in range(5):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    marks.clear()
    tree = resident()
    torch.cuda.synchronize()
if rank == 0:
    print(f"=== N={world} resident step (rank 0), indexed {tree.n_indexed} source faces, phases {tree.last_phase_ms}")
    for name, ev, th in marks[1:]:
        print(f"  {marks[0][1].elapsed_time(ev):8.3f} ms  {name}   [host {1e3 * (th - marks[0][2]):.2f}]")
if world > 1:
    dist.destroy_process_group()
