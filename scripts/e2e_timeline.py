"""Where the end-to-end step (public API, host buffers) spends its time: CUDA-event
timestamps on the compute stream at the milestones of OverlapRegridder + regrid."""
import sys
import numpy as np
import torch
import xugrid_b200 as xb
from xugrid_b200 import celltree, synthetic, regridder

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
# chunk specs separated by ';': an integer (equal chunks) or cumulative fractions "0.125,0.5,1"
def _spec(t):
    return int(t) if t.isdigit() else tuple(float(f) for f in t.split(","))
chunks = [_spec(t) for t in sys.argv[2].split(";")] if len(sys.argv) > 2 else [4]
sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
values = torch.from_numpy(np.random.default_rng(12).random((1, sf.shape[0]))).pin_memory()
src = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
tgt = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
for g in (src, tgt):
    g.prefetch()
torch.cuda.synchronize()

marks = []
def mark(name):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    marks.append((name, ev))

orig_init = celltree.CellTree2d.__init__
orig_csr = celltree.CellTree2d.intersect_faces_csr
def init(self, *a, **k):
    orig_init(self, *a, **k); mark("index built")
def csr(self, *a, **k):
    r = orig_csr(self, *a, **k); mark(f"overlap rows [{k.get('t_begin')},{k.get('t_end')})"); return r
from xugrid_b200 import sparse as _sp
orig_apply = _sp.csr_apply
def apply(*a, **k):
    mark("values on device (apply starts)"); r = orig_apply(*a, **k); mark("apply done"); return r
_sp.csr_apply = apply
celltree.CellTree2d.__init__ = init
celltree.CellTree2d.intersect_faces_csr = csr

for nchunk in chunks:
    regridder.BaseRegridder._TARGET_CHUNKS = nchunk
    for rep in range(4):
        src.release_device(); tgt.release_device()
        torch.cuda.synchronize()
        marks.clear()
        import time
        t0 = time.perf_counter()
        mark("start")
        r = xb.OverlapRegridder(src, tgt, method="mean")
        mark("weights done")
        pend = r.__dict__.get("_pending")
        out = r.regrid(values)
        mark("regrid done (after D2H sync)")
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
    print(f"--- chunks={nchunk}  wall {wall:.2f} ms")
    rows = [(marks[0][1].elapsed_time(ev), name) for name, ev in marks[1:]]
    for label, g in (("source", src), ("target", tgt)):
        rows += [(marks[0][1].elapsed_time(ev), f"[copy stream] {label} {name}") for name, ev in getattr(g, "_upload_events", [])]
    for t, name in sorted(rows):
        print(f"  {t:8.3f} ms  {name}")
    if pend is not None:
        pend.complete()
        print("  jobs launched as graphs:", [getattr(j, "launched_as_graph", None) for j in getattr(pend, "_all_jobs", [])])
        for j, ph in enumerate(pend.job_phase_ms):
            print(f"  job {j}: " + "  ".join(f"{k} {v:.3f}" for k, v in ph.items()) + f"  sum {sum(ph.values()):.3f}")
