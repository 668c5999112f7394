"""world_size-2 gloo test of the sharded output assembly (runs on CPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from xugrid_b200 import synthetic
from xugrid_b200.distributed import all_gather_columns, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, n_extra, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank regrids only its own target range (the oracle stands in
        # for the GPU here: this test is about the host-side sharding logic)
        sxy, sf = synthetic.delaunay(400, 1)
        txy, tf = synthetic.jittered_lattice(9, 7, 2, "tri", bounds=(0.1, 0.9, 0.1, 0.9))
        assert len(tf) == n_total
        b, e = shard_range(n_total, rank, world)
        tree = oracle.CellTree2d(sxy, sf)
        ti, si, w = tree.intersect_faces(txy, tf[b:e])
        indptr = oracle.coo_to_csr(ti, e - b)
        vals = np.random.default_rng(3).random((n_extra, len(sf)))
        local = torch.from_numpy(oracle.csr_apply("mean", vals, w, si, indptr, e - b))
        full = all_gather_columns(local, n_total)
        if rank == 0:
            ti, si, w = tree.intersect_faces(txy, tf)
            indptr = oracle.coo_to_csr(ti, n_total)
            ref = oracle.csr_apply("mean", vals, w, si, indptr, n_total)
            q.put(bool(np.array_equal(full.numpy(), ref, equal_nan=True)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_regrid_matches_single():
    world, n_total, n_extra = 2, 126, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, n_extra, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def _upload_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xugrid_b200.distributed import sharded_upload

        ok = True
        for n in (0, 1, 7, 10):  # shorter than, not a multiple of, and a multiple of the world size
            host = torch.arange(n * 3, dtype=torch.int32).reshape(n, 3)
            full = sharded_upload(host, torch.device("cpu"))
            ok = ok and full.shape == host.shape and bool(torch.equal(full, host))
        xy = torch.from_numpy(np.random.default_rng(5).random((11, 2)))
        ok = ok and bool(torch.equal(sharded_upload(xy, torch.device("cpu")), xy))
        q.put(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_upload_replicates_a_host_array():
    """every rank uploads 1/world of a host array; the all-gather gives all of it to each"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_upload_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True and q.get(timeout=10) is True


def _csr_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xugrid_b200.distributed import all_gather_csr
        from xugrid_b200.sparse import MatrixCSR

        sxy, sf = synthetic.delaunay(300, 7)
        txy, tf = synthetic.jittered_lattice(7, 9, 8, "tri", bounds=(0.1, 0.9, 0.1, 0.9))
        n_total = len(tf)  # 126 rows over 3 ranks: the last shard is the same size, then try 4 below
        tree = oracle.CellTree2d(sxy, sf)
        ok = True
        for n_rows in (n_total, n_total - 5, 1):
            b, e = shard_range(n_rows, rank, world)
            ti, si, w = tree.intersect_faces(txy, tf[b:e]) if e > b else (np.zeros(0, np.int64),) * 2 + (np.zeros(0),)
            order = np.lexsort((si, ti))
            ti, si, w = ti[order], si[order], w[order]
            indptr = oracle.coo_to_csr(ti, e - b)
            local = MatrixCSR(torch.from_numpy(w), torch.from_numpy(si.astype(np.int32)), torch.from_numpy(indptr),
                              e - b, len(sf), len(w))
            full = all_gather_csr(local, n_rows)
            ti, si, w = tree.intersect_faces(txy, tf[:n_rows])
            order = np.lexsort((si, ti))
            ref_ptr = oracle.coo_to_csr(ti[order], n_rows)
            ok = ok and full.n == n_rows and full.nnz == len(w)
            ok = ok and np.array_equal(full.indptr.numpy(), ref_ptr)
            ok = ok and np.array_equal(full.indices.numpy(), si[order])
            ok = ok and np.array_equal(full.data.numpy(), w[order])
        q.put(ok)
    finally:
        dist.destroy_process_group()


def test_all_gather_csr_reassembles_sharded_weights():
    """three ranks (uneven and empty shards included) gather their CSR rows into the full matrix"""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_csr_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert all(q.get(timeout=10) is True for _ in range(world))


def _shared_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xugrid_b200.distributed import shared_host_array

        n = 37
        out = shared_host_array((2, n), torch.float64)
        b, e = shard_range(n, rank, world)
        # every rank writes only its own columns; after a barrier all of them read the whole array
        out[:, b:e] = torch.arange(b, e, dtype=torch.float64)[None, :] + 100.0 * torch.arange(2, dtype=torch.float64)[:, None]
        dist.barrier()
        expect = torch.arange(n, dtype=torch.float64)[None, :] + 100.0 * torch.arange(2, dtype=torch.float64)[:, None]
        second = shared_host_array(5, torch.int32)  # a second array gets its own segment
        second[rank] = rank + 1
        dist.barrier()
        ok = bool(torch.equal(out, expect)) and second[:world].tolist() == list(range(1, world + 1))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_shared_host_array_is_one_array_for_all_ranks():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=10) for _ in range(world))
    assert got == {0: True, 1: True, 2: True}
