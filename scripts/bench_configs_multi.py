"""
BASELINE.json configurations 4 and 5 as they are named: sharded over the N GPUs of one box
(one process per GPU, torchrun), each with ORACLE EQUALITY on a sample of target faces --
pair set and weights bit for bit, regridded values bit for bit -- next to the stage times.

  config 4  OverlapRegridder mean: 10M-face mixed quad/tri -> 10M-face tri, 256-step apply;
            every rank owns a contiguous range of target rows (the lattice numbering is
            spatial), replicated source culled to the range, batched apply of its columns,
            one all-gather of the (256, n_target) output
  config 5  RelativeOverlapRegridder + reduction sweep (min / max / mean / sum on the
            absolute weights, first_order_conservative on the relative ones): 50M-face tri
            source -> 2M-face Delaunay target; qhull's face numbering means nothing in
            space, so the ranks share the targets along the space-filling curve
            (distributed.curve_shard) and the outputs are scattered back by face number

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/bench_configs_multi.py [4 5] [--small]
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

import oracle
import xugrid_b200 as xb
from xugrid_b200 import sparse, synthetic
from xugrid_b200 import distributed as xd

small = "--small" in sys.argv
which = [int(a) for a in sys.argv[1:] if a.isdigit()] or [4, 5]
rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N_SAMPLE = 2_000 if small else 20_000


def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)


def max_over_ranks(ms: float) -> float:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, reps=2, warm=2):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, max_over_ranks(e0.elapsed_time(e1) / reps)


def oracle_rows(sxy, sf, txy, tf, sample, relative=False):
    """CSR rows of the sampled target faces from the CPU oracle (all host threads)."""
    oracle.set_num_threads(os.cpu_count() or 1)
    tree = oracle.CellTree2d(sxy, sf)
    ti, si, w = tree.intersect_faces(txy, np.ascontiguousarray(tf[sample]))
    order = np.lexsort((si, ti))
    ti, si, w = ti[order], si[order], w[order]
    if relative:
        w = w / oracle.face_area(sxy, sf)[si]
    return oracle.coo_to_csr(ti, len(sample)), si, w


def rows_equal(indptr, indices, data, rows_of_sample, o_ptr, o_idx, o_w):
    """the sampled rows of a device CSR against the oracle's: same columns, same bits"""
    ip = indptr.cpu().numpy()
    lo, hi = ip[rows_of_sample], ip[rows_of_sample + 1]
    if not np.array_equal(hi - lo, np.diff(o_ptr)):
        return False, False
    take = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)]) if len(lo) else np.zeros(0, np.int64)
    take_t = torch.from_numpy(take).to(indices.device)
    idx = indices[take_t].cpu().numpy().astype(np.int64)
    w = data[take_t].cpu().numpy()
    return bool(np.array_equal(idx, o_idx)), bool(np.array_equal(w.view(np.int64), o_w.view(np.int64)))


if 4 in which:
    ns, nt, steps = (300, 260, 16) if small else (2582, 2236, 256)
    sxy, sf = synthetic.jittered_lattice(ns, ns, 30, "mixed")
    txy, tf = synthetic.jittered_lattice(nt, nt, 32, "tri")
    n_s, n_t = len(sf), len(tf)
    dsxy, dsf = torch.from_numpy(sxy).to(dev), torch.from_numpy(sf.astype(np.int32)).to(dev)
    dtxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf.astype(np.int32)).to(dev)
    b, e = xd.shard_range(n_t, rank, world)

    def build():
        clip = (dtxy, dtf, b, e) if world > 1 else None
        tree = xb.CellTree2d(dsxy, dsf, device=dev, clip_targets=clip)
        ip, ix, w = tree.intersect_faces_csr(dtxy, dtf, t_begin=b, t_end=e)
        return tree, sparse.MatrixCSR(w, ix, ip, e - b, n_s, int(w.numel()))

    (tree, A), build_ms = timed(build)
    g = torch.Generator(device=dev).manual_seed(33)  # same seed on every rank: replicated values
    vals = torch.rand((steps, n_s), dtype=torch.float64, device=dev, generator=g)
    vals[torch.rand(vals.shape, device=dev, generator=g) < 0.01] = float("nan")
    local, apply_ms = timed(lambda: sparse.csr_apply(0, 0.0, vals, A, e - b), reps=1)
    full, gather_ms = timed(lambda: xd.all_gather_columns(local, n_t), reps=1, warm=1)
    # ---- oracle equality on a sample of this job's target faces (rank 0 checks rows of ALL ranks:
    #      it reads them from the gathered output; the weights of its own range)
    check = None
    sample = np.sort(np.random.default_rng(34).choice(n_t, N_SAMPLE, replace=False))
    if rank == 0:
        t0 = time.time()
        o_ptr, o_idx, o_w = oracle_rows(sxy, sf, txy, tf, sample)
        hv = vals[:3].cpu().numpy()
        o_out = oracle.csr_apply("mean", hv, o_w, o_idx, o_ptr, len(sample))
        got = full[:3][:, torch.from_numpy(sample).to(dev)].cpu().numpy()
        mine = sample[sample < e] - b
        n_mine = len(mine)
        same_idx, same_w = rows_equal(A.indptr, A.indices, A.data, mine, o_ptr[: n_mine + 1], o_idx[: o_ptr[n_mine]],
                                      o_w[: o_ptr[n_mine]])
        check = {"sample_targets": int(len(sample)), "sample_rows_of_rank0": int(n_mine),
                 "pair_set_equal": same_idx, "weights_bit_identical": same_w,
                 "regridded_values_bit_identical_all_ranks": bool(np.array_equal(got, o_out, equal_nan=True)),
                 "oracle_s": time.time() - t0}
    nnz = torch.tensor([A.nnz], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nnz)
    emit({"config": 4, "n_gpus": world,
          "workload": f"OverlapRegridder mean: {n_s/1e6:.1f}M-face mixed quad/tri -> {n_t/1e6:.1f}M-face tri, {steps}-step apply, "
                      f"target rows sharded over {world} GPU(s)",
          "build_ms": build_ms, "apply_ms_all_steps": apply_ms, "all_gather_ms": gather_ms if world > 1 else 0.0,
          "face_steps_per_s": n_t * steps / ((apply_ms + (gather_ms if world > 1 else 0.0)) * 1e-3),
          "target_faces_per_s_build": n_t / (build_ms * 1e-3), "nnz": int(nnz.item()),
          "source_faces_indexed_on_rank0": tree.n_indexed, "oracle_check": check})
    del tree, A, vals, local, full
    torch.cuda.empty_cache()

if 5 in which:
    ns, npts = (700, 20_000) if small else (5000, 1_000_000)
    sxy, sf = synthetic.jittered_lattice(ns, ns, 40, "tri")
    txy, tf = synthetic.delaunay(npts, 41)
    n_s, n_t = len(sf), len(tf)
    dsxy, dsf = torch.from_numpy(sxy).to(dev), torch.from_numpy(sf.astype(np.int32)).to(dev)
    dtxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf.astype(np.int32)).to(dev)
    vals = torch.from_numpy(np.random.default_rng(42).random((1, n_s))).to(dev)
    ids = xd.curve_shard(dtxy, dtf, rank, world)
    all_ids, lens = xd.all_gather_ids(ids)

    def build(relative):
        def f():
            clip_ids = (dtxy, dtf, ids) if world > 1 else None
            tree = xb.CellTree2d(dsxy, dsf, device=dev, clip_target_ids=clip_ids)
            ip, ix, w = tree.intersect_faces_csr(dtxy, dtf, relative=relative, ids=ids)
            return tree, sparse.MatrixCSR(w, ix, ip, int(ids.numel()), n_s, int(w.numel()))
        return f

    (tree, A), build_ms = timed(build(False))
    res, outs = {}, {}
    for name, code in (("mean", 0), ("sum", 3), ("minimum", 4), ("maximum", 5)):
        local, ms = timed(lambda: sparse.csr_apply(code, 0.0, vals, A, A.n))
        res[name] = ms
        outs[name] = xd.all_gather_by_ids(local, all_ids, lens, n_t)
    stats = tree.last_stats._asdict()
    n_indexed = tree.n_indexed
    nnz = torch.tensor([A.nnz], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nnz)
    _, gather_ms = timed(lambda: xd.all_gather_by_ids(local, all_ids, lens, n_t), reps=1, warm=1)
    # rows of rank 0's share against the oracle (absolute weights)
    sample = np.sort(np.random.default_rng(43).choice(n_t, N_SAMPLE, replace=False))
    check = None
    if rank == 0:
        t0 = time.time()
        o_ptr, o_idx, o_w = oracle_rows(sxy, sf, txy, tf, sample)
        hv = vals.cpu().numpy()
        same_out = {}
        for name in ("mean", "sum", "minimum", "maximum"):
            o_out = oracle.csr_apply(name, hv, o_w, o_idx, o_ptr, len(sample))
            got = outs[name][:, torch.from_numpy(sample).to(dev)].cpu().numpy()
            same_out[name] = bool(np.array_equal(got, o_out, equal_nan=True))
        # rows of the sample that rank 0 owns: position of every sampled face in its id list
        my_ids = ids.cpu().numpy().astype(np.int64)
        pos = np.full(n_t, -1, dtype=np.int64)
        pos[my_ids] = np.arange(len(my_ids))
        owned = pos[sample] >= 0
        rows = pos[sample][owned]
        keep_entries = np.repeat(owned, np.diff(o_ptr))
        sub_ptr = np.concatenate([[0], np.cumsum(np.diff(o_ptr)[owned])])
        same_idx, same_w = rows_equal(A.indptr, A.indices, A.data, rows, sub_ptr, o_idx[keep_entries], o_w[keep_entries])
        check = {"sample_targets": int(len(sample)), "sample_rows_of_rank0": int(owned.sum()),
                 "pair_set_equal": same_idx, "weights_bit_identical": same_w,
                 "regridded_values_bit_identical_all_ranks": same_out, "oracle_s": time.time() - t0}
    del tree, A
    (tree, R), rel_build_ms = timed(build(True))
    local, cons_ms = timed(lambda: sparse.csr_apply(9, 0.0, vals, R, R.n))
    cons = xd.all_gather_by_ids(local, all_ids, lens, n_t)
    if rank == 0:
        # (relative weights = the absolute ones over the area of their source face: no second tree)
        o_w = o_w / oracle.face_area(sxy, sf)[o_idx]
        o_out = oracle.csr_apply("first_order_conservative", vals.cpu().numpy(), o_w, o_idx, o_ptr, len(sample))
        got = cons[:, torch.from_numpy(sample).to(dev)].cpu().numpy()
        check["first_order_conservative_close_1e-12"] = bool(np.allclose(got, o_out, rtol=1e-12, atol=0, equal_nan=True))
        check["first_order_conservative_bit_identical"] = bool(np.array_equal(got, o_out, equal_nan=True))
    emit({"config": 5, "n_gpus": world,
          "workload": f"Relative/OverlapRegridder sweep: {n_s/1e6:.0f}M-face tri source -> {n_t/1e6:.2f}M-face Delaunay target, "
                      f"targets shared along the space-filling curve over {world} GPU(s)",
          "build_ms": build_ms, "relative_build_ms": rel_build_ms, "apply_ms": res, "first_order_conservative_ms": cons_ms,
          "all_gather_by_ids_ms": gather_ms if world > 1 else 0.0, "nnz": int(nnz.item()), "stats_rank0": stats,
          "source_faces_indexed_on_rank0": n_indexed, "oracle_check": check})

if world > 1:
    dist.destroy_process_group()
