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

import oracle
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


N_SAMPLE = 2_000 if small else 20_000


def overlap_rows_equal_oracle(sxy, sf, txy, tf, A, seed):
    """pair set and weights of N_SAMPLE target rows of a device CSR against the CPU oracle"""
    import os

    oracle.set_num_threads(os.cpu_count() or 1)
    sample = np.sort(np.random.default_rng(seed).choice(len(tf), min(N_SAMPLE, len(tf)), replace=False))
    ti, si, w = oracle.CellTree2d(sxy, sf).intersect_faces(txy, np.ascontiguousarray(tf[sample]))
    order = np.lexsort((si, ti))
    ti, si, w = ti[order], si[order], w[order]
    o_ptr = oracle.coo_to_csr(ti, len(sample))
    ip = A.indptr.cpu().numpy()
    lo, hi = ip[sample], ip[sample + 1]
    same_counts = bool(np.array_equal(hi - lo, np.diff(o_ptr)))
    take = torch.from_numpy(np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)])).to(A.data.device)
    idx = A.indices[take].cpu().numpy().astype(np.int64)
    wd = A.data[take].cpu().numpy()
    return {"sample_targets": int(len(sample)), "pair_set_equal": same_counts and bool(np.array_equal(idx, si)),
            "weights_bit_identical": same_counts and bool(np.array_equal(wd.view(np.int64), w.view(np.int64)))}


def bary_rows_close_to_oracle(sxy, sf, mid, A, reg):
    """BarycentricInterpolator rows of N_SAMPLE raster points against the CPU statement of the same
    path: Voronoi topology by the tensor-op version on CPU tensors (pinned by the reference's golden
    vectors), point location / Wachspress weights by the oracle, corner redistribution and outside
    mask as in unstructured.py:146-201.  Columns exact, weights to 1e-10."""
    from xugrid_b200 import voronoi

    nr = len(mid)
    sample = np.sort(np.random.default_rng(22).choice(nr * nr, min(N_SAMPLE, nr * nr), replace=False))
    # the very points the regridder queried: the centroids of the raster cells as laid out on the
    # device (they differ from the midpoints in the last bit, which a Wachspress weight next to a
    # polygon edge magnifies to 1e-12)
    from xugrid_b200.regridder import _target_points

    pts = _target_points(reg._target, dev)[torch.from_numpy(sample).to(dev)].cpu().numpy()
    assert np.allclose(pts, np.column_stack([mid[sample % nr], mid[sample // nr]]), rtol=0, atol=1e-12)  # C order: row = y index
    cent = oracle.face_centroids(sxy, sf)
    topo = voronoi.voronoi_topology(torch.from_numpy(sf), torch.from_numpy(sxy), torch.from_numpy(cent))
    tree = oracle.CellTree2d(topo.vertices.numpy(), topo.faces.numpy(), -1)
    fi, w = tree.compute_barycentric_weights(pts)
    conn = topo.faces[torch.from_numpy(fi).clamp(min=0)]
    w = torch.from_numpy(w)
    voronoi.redistribute_corner_weights(topo, conn, w)
    outside = oracle.CellTree2d(sxy, sf, -1).locate_points(pts) == -1
    w[torch.from_numpy(outside)] = 0
    keep = w > 0
    o_cnt = keep.sum(1).numpy()
    o_idx = topo.vertex_face[conn.clamp(min=0)][keep].numpy()
    o_w = w[keep].numpy()
    ip = A.indptr.cpu().numpy()
    lo, hi = ip[sample], ip[sample + 1]
    same_counts = bool(np.array_equal(hi - lo, o_cnt))
    take = torch.from_numpy(np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)])).to(A.data.device)
    idx = A.indices[take].cpu().numpy().astype(np.int64)
    wd = A.data[take].cpu().numpy()
    return {"sample_points": int(len(sample)), "columns_equal": same_counts and bool(np.array_equal(idx, o_idx)),
            "weights_close_1e-10": same_counts and bool(np.allclose(wd, o_w, rtol=1e-10, atol=1e-15)),
            "max_abs_weight_diff": float(np.abs(wd - o_w).max()) if same_counts and len(o_w) else None}


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
          "oracle_check": overlap_rows_equal_oracle(sxy, sf, txy, tf, A, 13),
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
                     "linear_field_max_abs_err": float((got[ok] - want[ok]).abs().max())},
          "oracle_check": bary_rows_close_to_oracle(sxy, sf, mid, A, reg)})
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
