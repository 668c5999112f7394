"""
The BASELINE.json configurations 2-5 on ONE B200 (resident inputs), each with
size-independent checks; one JSON line per configuration.  The headline bench is
bench.py -- this script documents that the other configurations run at their
named sizes and how long their stages take.

    python scripts/bench_configs.py [2 3 4 5] [--small]
"""
import json
import sys
import time

import numpy as np
import torch

import xugrid_b200 as xb
from xugrid_b200 import sparse, synthetic

small = "--small" in sys.argv
which = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2, 3, 4, 5]
dev = torch.device("cuda", 0)


def timed(fn, reps=2):
    out = fn()
    out = fn()  # two warm-ups: a result kept alive while the next is built needs two sets of cached blocks
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def to_dev(xy, f):
    return torch.from_numpy(xy).to(dev), torch.from_numpy(f.astype(np.int32)).to(dev)


def overlap(sxy, sf, txy, tf, relative=False):
    def build():
        tree = xb.CellTree2d(sxy, sf, device=dev)
        ip, ix, w = tree.intersect_faces_csr(txy, tf, relative=relative)
        return tree, sparse.MatrixCSR(w, ix, ip, int(tf.shape[0]), int(sf.shape[0]), int(w.numel()))
    return timed(build)


def emit(d):
    print(json.dumps(d), flush=True)


if 2 in which:
    n = 50_000 if small else 500_000
    t0 = time.time()
    sxy, sf = synthetic.delaunay(n, 10)
    txy, tf = synthetic.delaunay(n, 11)
    gen_s = time.time() - t0
    dsxy, dsf = to_dev(sxy, sf)
    dtxy, dtf = to_dev(txy, tf)
    (tree, A), build_ms = overlap(dsxy, dsf, dtxy, dtf)
    vals = torch.from_numpy(np.random.default_rng(12).random((1, len(sf)))).to(dev)
    out, apply_ms = timed(lambda: sparse.csr_apply(0, 0.0, vals, A, A.n))
    # checks: a target inside the source hull is fully covered (row sum == its area);
    # mean of a constant field is that constant
    t_area = xb.celltree.face_area_device(dtxy, dtf)
    rows = torch.repeat_interleave(torch.arange(A.n, device=dev), A.indptr[1:] - A.indptr[:-1])
    row_sum = torch.zeros(A.n, dtype=torch.float64, device=dev).index_add_(0, rows, A.data)
    inside = (row_sum - t_area).abs() <= 1e-9 * t_area
    ones = sparse.csr_apply(0, 0.0, torch.full_like(vals, 3.5), A, A.n)
    emit({"config": 2, "workload": f"OverlapRegridder mean: {len(sf)/1e6:.2f}M-face Delaunay -> {len(tf)/1e6:.2f}M-face Delaunay, 1 step",
          "build_ms": build_ms, "apply_ms": apply_ms, "target_faces_per_s": len(tf) / ((build_ms + apply_ms) * 1e-3),
          "nnz": A.nnz, "stats": tree.last_stats._asdict(),
          "checks": {"rows_fully_covered_frac": float(inside.double().mean()),
                     "row_sum_never_exceeds_area": bool((row_sum <= t_area * (1 + 1e-9)).all()),
                     "mean_of_constant_max_rel_err": float(((ones[~torch.isnan(ones)] - 3.5).abs() / 3.5).max())},
          "host_mesh_generation_s": gen_s})
    del tree, A, vals, out

if 3 in which:
    ns, nr = (400, 512) if small else (2236, 4096)
    sxy, sf = synthetic.jittered_lattice(ns, ns, 20, "tri")   # (ns+1)^2 nodes, 2 ns^2 triangles
    source = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf, device=dev)
    c = np.linspace(0.02, 0.98, nr + 1)
    mid = 0.5 * (c[:-1] + c[1:])
    target = xb.StructuredGrid2d(mid, mid)
    source.prefetch()
    torch.cuda.synchronize()
    t0 = time.time()
    reg, build_ms = timed(lambda: xb.BarycentricInterpolator(source, target), reps=1)
    vals = torch.from_numpy(np.random.default_rng(21).random((1, len(sf)))).to(dev)
    out, apply_ms = timed(lambda: reg.regrid_device(vals))
    # checks: weights of a row sum to 1; a linear field is reproduced at the target centroids
    A = reg._weights
    rows = torch.repeat_interleave(torch.arange(A.n, device=dev), A.indptr[1:] - A.indptr[:-1])
    row_sum = torch.zeros(A.n, dtype=torch.float64, device=dev).index_add_(0, rows, A.data)
    has = (A.indptr[1:] - A.indptr[:-1]) > 0
    cen = source.centroids_device
    lin = (2.0 * cen[:, 0] - 3.0 * cen[:, 1] + 0.25).reshape(1, -1)
    got = reg.regrid_device(lin).reshape(-1)
    yy, xx = torch.meshgrid(torch.from_numpy(mid).to(dev), torch.from_numpy(mid).to(dev), indexing="ij")
    want = (2.0 * xx - 3.0 * yy + 0.25).reshape(-1)
    ok = ~torch.isnan(got)
    emit({"config": 3, "workload": f"BarycentricInterpolator: {source.n_node/1e6:.2f}M-node jittered-lattice tri source ({len(sf)/1e6:.1f}M faces) -> {nr}x{nr} raster points",
          "build_ms": build_ms, "apply_ms": apply_ms, "points_per_s_build": nr * nr / (build_ms * 1e-3), "nnz": A.nnz,
          "checks": {"rows_with_weights_frac": float(has.double().mean()),
                     "max_abs_row_sum_minus_1": float((row_sum[has] - 1.0).abs().max()),
                     "linear_field_max_abs_err": float((got[ok] - want[ok]).abs().max())}})
    del reg, A, vals, out, source, target

if 4 in which:
    ns, nt, steps = (300, 260, 16) if small else (2582, 2236, 256)
    sxy, sf = synthetic.jittered_lattice(ns, ns, 30, "mixed")
    txy, tf = synthetic.jittered_lattice(nt, nt, 32, "tri")
    dsxy, dsf = to_dev(sxy, sf)
    dtxy, dtf = to_dev(txy, tf)
    (tree, A), build_ms = overlap(dsxy, dsf, dtxy, dtf)
    g = torch.Generator(device=dev).manual_seed(33)
    vals = torch.rand((steps, len(sf)), dtype=torch.float64, device=dev, generator=g)
    vals[torch.rand(vals.shape, device=dev, generator=g) < 0.01] = float("nan")
    out, apply_ms = timed(lambda: sparse.csr_apply(0, 0.0, vals, A, A.n), reps=1)
    one, one_ms = timed(lambda: sparse.csr_apply(0, 0.0, vals[:1], A, A.n))
    # checks: batched apply == step-by-step apply (bitwise); result within the range of the data
    same = bool(torch.equal(torch.nan_to_num(out[:1], nan=-1.0), torch.nan_to_num(one, nan=-1.0)))
    fin = out[~torch.isnan(out)]
    emit({"config": 4, "workload": f"OverlapRegridder mean: {len(sf)/1e6:.1f}M-face mixed quad/tri -> {len(tf)/1e6:.1f}M-face tri, {steps}-step apply",
          "build_ms": build_ms, "apply_ms_all_steps": apply_ms, "apply_ms_one_step": one_ms,
          "face_steps_per_s": len(tf) * steps / (apply_ms * 1e-3), "nnz": A.nnz, "stats": tree.last_stats._asdict(),
          "apply_bytes_model_per_step": A.nnz * 20 + A.n * 16,
          "checks": {"batched_equals_single_step": same, "values_within_data_range": bool((fin >= 0).all() and (fin <= 1).all())}})
    del tree, A, vals, out

if 5 in which:
    ns, npts = (700, 20_000) if small else (5000, 1_000_000)
    sxy, sf = synthetic.jittered_lattice(ns, ns, 40, "tri")
    t0 = time.time()
    txy, tf = synthetic.delaunay(npts, 41)
    gen_s = time.time() - t0
    dsxy, dsf = to_dev(sxy, sf)
    dtxy, dtf = to_dev(txy, tf)
    res = {}
    (tree, A), build_ms = overlap(dsxy, dsf, dtxy, dtf, relative=False)
    vals = torch.from_numpy(np.random.default_rng(42).random((1, len(sf)))).to(dev)
    for name, code in (("mean", 0), ("sum", 3), ("minimum", 4), ("maximum", 5)):
        out, ms = timed(lambda: sparse.csr_apply(code, 0.0, vals, A, A.n))
        res[name] = ms
    mn = sparse.csr_apply(4, 0.0, vals, A, A.n)
    me = sparse.csr_apply(0, 0.0, vals, A, A.n)
    mx = sparse.csr_apply(5, 0.0, vals, A, A.n)
    ok = ~torch.isnan(me)
    tol = 1e-12
    ordered = bool(((mn[ok] <= me[ok] * (1 + tol)) & (me[ok] <= mx[ok] * (1 + tol))).all())
    stats = tree.last_stats._asdict()
    nnz_abs = A.nnz
    del tree, A
    (tree, R), rel_build_ms = overlap(dsxy, dsf, dtxy, dtf, relative=True)
    out, cons_ms = timed(lambda: sparse.csr_apply(9, 0.0, vals, R, R.n))
    # check: relative weights of a source face sum to 1 over the targets when it is fully covered
    col_sum = torch.zeros(len(sf), dtype=torch.float64, device=dev).index_add_(0, R.indices.long(), R.data)
    emit({"config": 5, "workload": f"Relative/OverlapRegridder sweep: {len(sf)/1e6:.0f}M-face tri source -> {len(tf)/1e6:.2f}M-face Delaunay target",
          "build_ms": build_ms, "relative_build_ms": rel_build_ms, "apply_ms": res, "first_order_conservative_ms": cons_ms,
          "nnz": nnz_abs, "stats": stats,
          "checks": {"min_le_mean_le_max": ordered,
                     "source_faces_with_relative_weight_sum_1_frac": float(((col_sum - 1.0).abs() < 1e-9).double().mean()),
                     "max_relative_column_sum": float(col_sum.max())},
          "host_mesh_generation_s": gen_s})
