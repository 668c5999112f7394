#!/usr/bin/env python
"""
bench.py -- target faces/s of OverlapRegridder(method="mean") weight build +
apply on synthetic triangle meshes (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full pass of the hot path over the workload: spatial index of
the source mesh + overlap weights (CSR) for this rank's target range + one
apply of `mean` + (N > 1) the all-gather that reassembles the output.

 * value : whole-job target faces/s with the meshes already resident in HBM
 * e2e   : the same step through the public API with HOST (pinned) buffers:
           H2D of both meshes and the values, compute, D2H of the result
 * roofline    : the dominant kernel (k_clip_survivors), algorithmic bytes / its measured time
 * cpu_baseline: the CPU oracle (C/OpenMP restatement of the reference's path)
                 on a bounded sample of the same workload, N = 1 only
 * --impl reference: the oracle port on all host threads (the reference is
   pure Python + numba + numba_celltree and cannot be installed offline)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "target faces/s (weight build + apply) OverlapRegridder mean tri->tri"
UNIT = "target_faces/s"
CPU_SAMPLE_SIDE = 1414  # 4.0M -> 4.0M triangles for the CPU legs


def make_workload(n_side: int):
    """source: n x n jittered lattice, 2 triangles per cell; target: an
    independent (n+3) x (n-2) lattice over the same unit square."""
    from xugrid_b200 import synthetic

    sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
    txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
    values = np.random.default_rng(12).random((1, sf.shape[0]))
    return sxy, sf, txy, tf, values


def workload_name(n_side: int) -> str:
    ns, nt = 2 * n_side * n_side, 2 * (n_side + 3) * (n_side - 2)
    return (f"OverlapRegridder mean: {ns / 1e6:.1f}M-face jittered-lattice tri -> {nt / 1e6:.1f}M-face tri, "
            f"float64, int64 connectivity, 1 time step")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
                for nme, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


# ----------------------------------------------------------------------------- CPU legs
def cpu_step(sxy, sf, txy, tf, values):
    """The reference's path on the CPU oracle: celltree build, intersect_faces,
    to_csr, one mean apply (single-threaded over rows like regridder.py:50)."""
    import oracle

    t0 = time.perf_counter()
    tree = oracle.CellTree2d(sxy, sf)
    t1 = time.perf_counter()
    ti, si, w = tree.intersect_faces(txy, tf)
    t2 = time.perf_counter()
    indptr = oracle.coo_to_csr(ti, len(tf))
    out = oracle.csr_apply("mean", values, w, si, indptr, len(tf))
    t3 = time.perf_counter()
    return out, {"tree_build_s": t1 - t0, "intersect_s": t2 - t1, "apply_s": t3 - t2, "total_s": t3 - t0,
                 "nnz": int(len(w))}


def cpu_baseline(n_side: int) -> dict:
    import oracle

    side = min(n_side, CPU_SAMPLE_SIDE)
    sxy, sf, txy, tf, values = make_workload(side)
    _, t = cpu_step(sxy, sf, txy, tf, values)
    return {
        "value": len(tf) / t["total_s"], "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
        "sample": (f"full build+apply of the same generator at {2 * side * side / 1e6:.1f}M -> {len(tf) / 1e6:.1f}M "
                   f"triangles ({side}x{side} lattice), C/OpenMP oracle; serial tree build {t['tree_build_s']:.2f}s, "
                   f"intersect {t['intersect_s']:.2f}s, apply {t['apply_s']:.2f}s (apply is single-threaded for one "
                   f"time step, as in the reference)"),
    }


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle

    side = min(args.n_side, CPU_SAMPLE_SIDE)
    sxy, sf, txy, tf, values = make_workload(side)
    for _ in range(args.warmup):
        cpu_step(sxy, sf, txy, tf, values)
    t0 = time.perf_counter()
    detail = None
    for _ in range(args.steps):
        _, detail = cpu_step(sxy, sf, txy, tf, values)
    dt = (time.perf_counter() - t0) / args.steps
    value = len(tf) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n_side),
                   "note": "reference arm = CPU oracle port (xugrid needs xarray + numba_celltree, absent offline)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                         "sample": f"each step = full build+apply at {2 * side * side / 1e6:.1f}M -> {len(tf) / 1e6:.1f}M "
                                   f"triangles; last step: {json.dumps(detail)}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args) -> int:
    import torch
    import torch.distributed as dist

    import xugrid_b200 as xb
    from xugrid_b200 import sparse
    from xugrid_b200.distributed import all_gather_columns, shard_range, sharded_upload, target_range_box

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout must carry exactly one JSON line, but native libraries write there too (NCCL
    # prints its version banner with printf): keep the real stdout aside for the result and
    # point file descriptor 1 at stderr for everything else
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    sxy, sf, txy, tf, values = make_workload(args.n_side)
    n_s, n_t = sf.shape[0], tf.shape[0]
    b, e = shard_range(n_t, rank, world)
    # host buffers (pinned) for the end-to-end leg
    h = {k: torch.from_numpy(v).pin_memory() for k, v in
         dict(sxy=sxy, sf=sf, txy=txy, tf=tf, values=values).items()}
    h_out = torch.empty((1, n_t), dtype=torch.float64).pin_memory()
    d = {k: v.to(dev) for k, v in h.items()}
    d2h_bytes = h_out.numel() * h_out.element_size()

    state = {}

    def step(src):
        box = target_range_box(src["txy"], src["tf"], b, e) if world > 1 else None
        tree = xb.CellTree2d(src["sxy"], src["sf"], device=dev, clip_box=box)
        indptr, indices, area = tree.intersect_faces_csr(src["txy"], src["tf"], t_begin=b, t_end=e)
        A = sparse.MatrixCSR(area, indices, indptr, e - b, n_s, int(area.numel()))
        local = sparse.csr_apply(0, 0.0, src["values"], A, e - b)
        full = all_gather_columns(local, n_t) if world > 1 else local
        state["tree"], state["A"], state["apply_in"] = tree, A, src["values"]
        return full

    # end-to-end leg: the public API on host data.  The grids are host containers
    # (numpy + a pinned staging copy made once, outside the timed region); every
    # step drops their device mirrors, so every step pays the full H2D of both
    # meshes and of the values, and the D2H of the result.
    src_grid = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf, device=dev)
    tgt_grid = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf, device=dev)
    for g in (src_grid, tgt_grid):
        g.prefetch()  # first call: pinned staging copy (int32 connectivity), outside the timed region
    torch.cuda.synchronize()
    nbytes = lambda t: t.numel() * t.element_size()
    if world == 1:
        h2d_bytes = sum(nbytes(t) for t in (src_grid._xy_pin, src_grid._faces_pin, tgt_grid._xy_pin,
                                            tgt_grid._faces_pin, h["values"]))
    else:
        # per rank: 1/world of the replicated arrays (assembled over NVLink) + its own target rows
        h2d_bytes = (sum(nbytes(t) for t in (src_grid._xy_pin, src_grid._faces_pin, tgt_grid._xy_pin,
                                             h["values"])) // world
                     + nbytes(tgt_grid._faces_pin[b:e]))

    def e2e_step():
        src_grid.release_device()
        tgt_grid.release_device()
        if world == 1:
            r = xb.OverlapRegridder(src_grid, tgt_grid, method="mean")
            state["e2e_out"] = r.regrid(h["values"])  # pinned host tensor in, pinned host tensor out
        else:
            r = xb.OverlapRegridder(src_grid, tgt_grid, method="mean", target_range=(b, e), sharded_upload=True)
            local = r.regrid(sharded_upload(h["values"].reshape(-1), dev).reshape(1, -1))
            h_out.copy_(all_gather_columns(local, n_t), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("XG_BENCH_NO_SMI")) else None
    ms_dev = timed(lambda: step(d), args.steps, args.warmup)
    # per-phase device times of the last resident step (CUDA events inside the library)
    phases = dict(state["tree"].last_phase_ms)
    stats = state["tree"].last_stats
    # apply kernel alone
    A = state["A"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        sparse.csr_apply(0, 0.0, d["values"], A, e - b)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(10):
        sparse.csr_apply(0, 0.0, d["values"], A, e - b)
    ev1.record()
    torch.cuda.synchronize()
    apply_ms = ev0.elapsed_time(ev1) / 10
    ms_e2e = timed(e2e_step, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        k_t = k_s = 3
        rows = e - b
        # algorithmic bytes per launch of the three big kernels:
        #  k_traverse        one 1 KB node (32 boxes x 32 B) per node visit, one 32 B box per target,
        #                    8 B per candidate pair written
        #  k_sat             per candidate: pair ids (8 B) + both polygons (k*16 B each); 8 B per survivor written
        #  k_clip_survivors  per survivor: pair ids (8 B) + both polygons + provisional row start (8 B),
        #                    one slot counter (4 B), one provisional entry written (4 + 8 B)
        kern = {
            "k_traverse": (stats.node_visits * 1024 + rows * 32 + stats.candidates * 8, phases.get("traverse")),
            "k_sat<3,3>": (stats.candidates * (8 + 16 * k_t + 16 * k_s) + stats.clip_pairs * 8, phases.get("sat")),
            "k_clip_survivors<3,3>": (stats.clip_pairs * (8 + 16 * k_t + 16 * k_s + 8 + 4 + 12), phases.get("clip")),
        }
        kernels = {}
        for name, (nbytes_k, ms_k) in kern.items():
            ach = nbytes_k / (ms_k * 1e-3) / 1e9 if ms_k and ms_k > 0 else None
            kernels[name] = {"algorithmic_bytes": nbytes_k, "ms": ms_k, "achieved": ach,
                             "frac": (ach / peak) if ach else None}
        dominant = max(kernels, key=lambda k_: kernels[k_]["ms"] or 0.0)
        clip_bytes, clip_ms = kernels[dominant]["algorithmic_bytes"], kernels[dominant]["ms"]
        achieved = kernels[dominant]["achieved"]
        # SURVEY 8(d) gather model for the whole build of this rank's rows
        # (wide BVH: one node visit = one coalesced load of 32 boxes x 32 B, shared by the 32
        #  targets of a tile, so the per-lane leaf tests C add no memory traffic)
        b_build = (rows * (k_t * 4 + k_t * 16) + stats.node_visits * 32 * 32
                   + stats.clip_pairs * (k_s * 4 + k_s * 16) + stats.nnz * 12 + (rows + 1) * 8)
        build_ms = sum(phases.values())
        b_apply = stats.nnz * (4 + 8 + 8) + rows * (8 + 8)
        traffic = None
        prof = os.path.join(ROOT, "profiles", "r1_dominant_kernel.json")
        if os.path.exists(prof) and world == 1:
            pj = json.load(open(prof))
            if pj.get("n_side") == args.n_side:
                traffic = pj.get("dram_bytes_per_launch", {}).get(dominant.split("<")[0])
        n_levels = state["tree"].n_levels
        line = {
            "metric": METRIC, "value": n_t / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args.n_side), "n_source_faces": n_s, "n_target_faces": n_t,
                "nnz": stats.nnz if world == 1 else None, "sharding": f"target rows [{b},{e}) on rank 0 of {world}",
                "l2": "inputs (2 x 320 MB meshes + 1.5 GB packed records) exceed the 126 MB L2; no explicit flush",
                "cub_launches_per_step": 12,
            },
            "e2e": {"value": n_t / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": 6 + n_levels + 9 + 1,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes": clip_bytes, "ms": clip_ms,
                         "note": "dominant = slowest kernel of this run; all three big kernels under "
                                 "roofline_kernels (instruction-issue bound, see DESIGN.md section 5)"},
            "roofline_kernels": kernels,
            "roofline_build": {"model": "SURVEY 8(d) gather model", "bytes": b_build, "ms": build_ms,
                               "achieved": b_build / (build_ms * 1e-3) / 1e9 if build_ms else None,
                               "frac": b_build / (build_ms * 1e-3) / 1e9 / peak if build_ms else None,
                               "V": stats.node_visits, "C": stats.leaf_tests, "C_clip": stats.clip_pairs,
                               "candidates": stats.candidates},
            "roofline_apply": {"bytes": b_apply, "ms": apply_ms, "achieved": b_apply / (apply_ms * 1e-3) / 1e9,
                               "frac": b_apply / (apply_ms * 1e-3) / 1e9 / peak},
            "phase_ms": phases,
            "clocks": clocks,
        }
        if world == 1 and not args.skip_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.n_side)
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--n-side", type=int, default=2236, help="source lattice side; 2236 -> 10.0M triangles")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
