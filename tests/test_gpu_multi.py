"""On-hardware multi-GPU check (needs >= 2 GPUs; skipped otherwise): one process per GPU
over NCCL, every rank builds and applies only its own target range, the all-gathered
output must be BIT-IDENTICAL to one GPU doing the whole target mesh -- the north star's
"results match ... at 1, 2, 4 and 8 GPUs".  (bench.py --gpus N asserts the same at the
10 M-face size, outside its timed region.)"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, order, q):
    import torch.distributed as dist

    import xugrid_b200 as xb
    from xugrid_b200 import synthetic
    from xugrid_b200.distributed import (all_gather_by_ids, all_gather_columns, all_gather_ids, curve_shard,
                                         shard_range)

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # (meshes this small would travel as int32: take the 24-bit transport of large meshes, which
    #  the PCIe-parallel upload shares between the ranks too)
    from xugrid_b200 import grid as grid_mod

    grid_mod.Ugrid2d._PACK_MIN_FACES = 100
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sxy, sf = synthetic.jittered_lattice(120, 110, 21, "mixed")
        txy, tf = synthetic.jittered_lattice(131, 97, 22, "tri")
        if order == "shuffled":
            tf = np.ascontiguousarray(tf[np.random.default_rng(5).permutation(len(tf))])
        n_t = len(tf)
        values = np.random.default_rng(23).random((3, len(sf)))
        values[1, ::11] = np.nan
        source = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
        target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
        b, e = shard_range(n_t, rank, world)
        ok = True
        for cls, method in ((xb.OverlapRegridder, "mean"), (xb.OverlapRegridder, "maximum"),
                            (xb.RelativeOverlapRegridder, "first_order_conservative")):
            r = cls(source, target, method=method, target_range=(b, e), sharded_upload=(method == "mean"))
            local = r.regrid(torch.from_numpy(values).cuda())
            full = all_gather_columns(local, n_t).contiguous()
            if rank == 0:
                alone = cls(source, target, method=method).regrid(torch.from_numpy(values).cuda())
                ok = ok and bool(torch.equal(full.view(torch.int64), alone.contiguous().view(torch.int64)))
            # and the weights of all shards collected on every rank
            r.gather_weights()
            again = r.regrid(torch.from_numpy(values).cuda())
            ok = ok and bool(torch.equal(again.contiguous().view(torch.int64), full.view(torch.int64)))
        # the apply fused with the all-gather (peer stores into every rank's copy of the output)
        # against apply + NCCL all-gather, twice (the second call overwrites what peers have read)
        from xugrid_b200.distributed import GatheredOutput

        gathered = GatheredOutput(3, n_t)
        dv = torch.from_numpy(values).cuda()
        for method in ("mean", "maximum"):
            r = xb.OverlapRegridder(source, target, method=method, target_range=(b, e))
            expect = all_gather_columns(r.regrid(dv), n_t).contiguous()
            code, p = r._method
            got = gathered.apply(code, p, dv, r._weights)
            ok = ok and bool(torch.equal(got.view(torch.int64), expect.view(torch.int64)))
        # shares along the space-filling curve (compact whatever the face numbering) + scatter by ids
        dxy, dtf = target.device_arrays()
        ids = curve_shard(dxy, dtf, rank, world)
        all_ids, lens = all_gather_ids(ids)
        r = xb.OverlapRegridder(source, target, method="mean", target_ids=ids)
        full = all_gather_by_ids(r.regrid(torch.from_numpy(values).cuda()), all_ids, lens, n_t)
        if rank == 0:
            alone = xb.OverlapRegridder(source, target, method="mean").regrid(torch.from_numpy(values).cuda())
            ok = ok and bool(torch.equal(full.contiguous().view(torch.int64), alone.contiguous().view(torch.int64)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("order", ["row_major", "shuffled"])
def test_sharded_regrid_is_bit_identical_to_one_gpu(order):
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, order, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(world))
    assert all(results.values()), results
