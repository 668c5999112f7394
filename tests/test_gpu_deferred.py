"""The overlap job enqueued without a host round trip (xg_overlap_run, what the regridders
use so that regrid() can queue behind their construction): same CSR as the waiting
path, bit for bit; chained target ranges fill one matrix; a job whose upper-bound arrays
are too small reports itself incomplete and the regridder falls back to the waiting path."""
import numpy as np
import pytest
import torch

from xugrid_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xugrid_b200

    assert torch.cuda.is_available()
    return xugrid_b200


def _meshes(kind_s="mixed", kind_t="tri"):
    sxy, sf = synthetic.jittered_lattice(61, 57, 31, kind_s)
    txy, tf = synthetic.jittered_lattice(66, 49, 32, kind_t)
    return sxy, sf, txy, tf


@pytest.mark.parametrize("relative", [False, True])
@pytest.mark.parametrize("n_ranges", [1, 3])
def test_run_equals_begin_finish(xb, relative, n_ranges):
    sxy, sf, txy, tf = _meshes()
    dev = torch.device("cuda", 0)
    tree = xb.CellTree2d(sxy, sf)
    dxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf).to(dev)
    n_t = len(tf)
    ip0, ix0, ar0 = tree.intersect_faces_csr(dxy, dtf, relative=relative)
    cap = 16 * n_t
    indptr = torch.full((n_t + 1,), -7, dtype=torch.int64, device=dev)
    indices = torch.empty(cap, dtype=torch.int32, device=dev)
    area = torch.empty(cap, dtype=torch.float64, device=dev)
    cuts = [n_t * c // n_ranges for c in range(n_ranges + 1)]
    jobs = [tree.intersect_faces_into(dxy, dtf, indptr[b: e + 1], indices, area, t_begin=b, t_end=e,
                                      relative=relative, chained=c > 0)
            for c, (b, e) in enumerate(zip(cuts[:-1], cuts[1:]))]
    end = 0
    for j in jobs:
        nnz, end, ok = j.result()
        assert ok
    assert end == int(ip0[-1])
    assert torch.equal(indptr, ip0)
    assert torch.equal(indices[:end], ix0)
    assert torch.equal(area[:end].view(torch.int64), ar0.view(torch.int64))
    assert jobs[-1].stats.nnz == int(ip0[cuts[-1]] - ip0[cuts[-2]])


def test_run_with_face_list(xb):
    sxy, sf, txy, tf = _meshes("tri", "mixed")
    dev = torch.device("cuda", 0)
    tree = xb.CellTree2d(sxy, sf)
    dxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf).to(dev)
    ids = torch.from_numpy(np.random.default_rng(3).permutation(len(tf))[:1500].astype(np.int32)).to(dev)
    ip0, ix0, ar0 = tree.intersect_faces_csr(dxy, dtf, ids=ids)
    indptr = torch.empty(ids.numel() + 1, dtype=torch.int64, device=dev)
    indices = torch.empty(16 * ids.numel(), dtype=torch.int32, device=dev)
    area = torch.empty(16 * ids.numel(), dtype=torch.float64, device=dev)
    nnz, end, ok = tree.intersect_faces_into(dxy, dtf, indptr, indices, area, ids=ids).result()
    assert ok and nnz == end == int(ip0[-1])
    assert torch.equal(indptr, ip0) and torch.equal(indices[:end], ix0)
    assert torch.equal(area[:end].view(torch.int64), ar0.view(torch.int64))


def test_run_reports_arrays_too_small(xb):
    sxy, sf, txy, tf = _meshes()
    dev = torch.device("cuda", 0)
    tree = xb.CellTree2d(sxy, sf)
    dxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf).to(dev)
    n_t = len(tf)
    nnz_true = int(tree.intersect_faces_csr(dxy, dtf)[0][-1])
    indptr = torch.empty(n_t + 1, dtype=torch.int64, device=dev)
    cap = nnz_true // 2
    guard = 64
    indices = torch.full((cap + guard,), -5, dtype=torch.int32, device=dev)
    area = torch.full((cap + guard,), -5.0, dtype=torch.float64, device=dev)
    nnz, end, ok = tree.intersect_faces_into(dxy, dtf, indptr, indices[:cap], area[:cap]).result()
    assert not ok
    assert nnz == nnz_true  # the row pointer is still exact
    # nothing was written behind the capacity the library was given
    assert bool((indices[cap:] == -5).all()) and bool((area[cap:] == -5.0).all())


def test_run_empty_jobs(xb):
    sxy, sf, txy, tf = _meshes()
    dev = torch.device("cuda", 0)
    tree = xb.CellTree2d(sxy, sf)
    dxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf).to(dev)
    indptr = torch.full((1,), 41, dtype=torch.int64, device=dev)
    indices = torch.empty(8, dtype=torch.int32, device=dev)
    area = torch.empty(8, dtype=torch.float64, device=dev)
    nnz, end, ok = tree.intersect_faces_into(dxy, dtf, indptr, indices, area, t_begin=5, t_end=5, chained=True).result()
    assert ok and nnz == 0 and int(indptr[0]) == 41
    # targets far away from the source: all rows empty, pointers = the chained offset
    far = torch.from_numpy(txy + 100.0).to(dev)
    indptr = torch.full((len(tf) + 1,), 9, dtype=torch.int64, device=dev)
    indices = torch.empty(16 * len(tf), dtype=torch.int32, device=dev)
    area = torch.empty(16 * len(tf), dtype=torch.float64, device=dev)
    nnz, end, ok = tree.intersect_faces_into(far, dtf, indptr, indices, area, chained=True).result()
    assert ok and nnz == 0 and end == 9 and bool((indptr == 9).all())


def test_regridder_queues_regrid_behind_construction(xb):
    """host (pinned) values in, host tensor out: the public path of bench.py's e2e leg"""
    sxy, sf, txy, tf = _meshes()
    src = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    tgt = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    values = torch.from_numpy(np.random.default_rng(7).random((2, len(sf)))).pin_memory()
    r = xb.OverlapRegridder(src, tgt, method="mean")
    assert r.__dict__.get("_pending") is not None  # nothing has been waited for yet
    out = r.regrid(values)
    assert r.__dict__.get("_pending") is not None or r.__dict__.get("_settled") is not None
    # reference: the waiting path
    tree = xb.CellTree2d(sxy, sf)
    ip, ix, ar = tree.intersect_faces_csr(txy, tf)
    from xugrid_b200 import sparse

    A = sparse.MatrixCSR(ar, ix, ip, len(tf), len(sf), int(ar.numel()))
    expect = sparse.csr_apply(0, 0.0, values.cuda(), A, len(tf)).cpu()
    assert torch.equal(out.view(torch.int64), expect.view(torch.int64))
    # the settled weights are the exact matrix
    W = r._weights
    assert W.nnz == int(ar.numel()) and torch.equal(W.indices, ix) and torch.equal(W.indptr, ip)
    # out=: the result lands in the caller's host tensor
    mine = torch.empty((2, len(tf)), dtype=torch.float64).pin_memory()
    got = r.regrid(values, out=mine)
    assert got is mine and torch.equal(mine.view(torch.int64), expect.view(torch.int64))
    with pytest.raises(ValueError):
        r.regrid(values, out=torch.empty((2, 3), dtype=torch.float64))


def test_regridder_falls_back_when_the_bound_is_too_small(xb, monkeypatch):
    from xugrid_b200 import regridder

    # (at least 65536 entries are always allocated: enough rows that 1 entry per row is too few)
    sxy, sf = synthetic.jittered_lattice(250, 240, 41, "mixed")
    txy, tf = synthetic.jittered_lattice(260, 250, 42, "tri")
    src = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    tgt = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    values = torch.from_numpy(np.random.default_rng(8).random((1, len(sf)))).pin_memory()
    good = xb.OverlapRegridder(src, tgt, method="mean").regrid(values)
    monkeypatch.setattr(regridder.UnstructuredGrid2d, "_ENTRIES_PER_ROW", 1)
    tgt2 = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    r = xb.OverlapRegridder(src, tgt2, method="mean")
    pending = r.__dict__.get("_pending")
    assert pending is not None and pending.data.numel() == len(tf)
    out = r.regrid(values)
    assert not pending.complete()
    assert torch.equal(out.view(torch.int64), good.view(torch.int64))
    assert r._weights.nnz > len(tf)


@pytest.mark.parametrize("bound", [16, 1])
def test_ranges_are_applied_and_downloaded_as_they_finish(xb, monkeypatch, bound):
    """a target mesh uploaded in face ranges: chained jobs, apply + D2H per range on the side
    stream (regridder._drain_ranges); bound=1: the arrays are too small and everything is redone"""
    from xugrid_b200 import regridder

    sxy, sf = synthetic.jittered_lattice(250, 240, 41, "mixed")
    txy, tf = synthetic.jittered_lattice(260, 250, 42, "tri")
    values = torch.from_numpy(np.random.default_rng(9).random((3, len(sf)))).pin_memory()
    values[1, ::7] = float("nan")
    src = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    whole = xb.OverlapRegridder(src, xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf), method="mean")
    assert len(whole.__dict__["_pending"].ranges) == 1
    good = whole.regrid(values)
    monkeypatch.setattr(regridder.BaseRegridder, "_CHUNK_MIN_FACES", 1000)
    monkeypatch.setattr(regridder.BaseRegridder, "_TARGET_CHUNKS", (0.1, 0.45, 0.8, 1.0))
    monkeypatch.setattr(regridder.UnstructuredGrid2d, "_ENTRIES_PER_ROW", bound)
    for cls, method in ((xb.OverlapRegridder, "mean"), (xb.RelativeOverlapRegridder, "first_order_conservative")):
        tgt = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)  # fresh: nothing resident, the upload is chunked
        r = cls(src, tgt, method=method)
        pending = r.__dict__["_pending"]
        assert len(pending.ranges) == 4 and pending.ranges[-1][1] == len(tf)
        out = r.regrid(values)
        assert pending.complete() == (bound == 16)
        if cls is xb.OverlapRegridder:
            assert torch.equal(out.view(torch.int64), good.view(torch.int64))
        else:
            expect = cls(src, xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf), method=method).regrid(values.cuda()).cpu()
            assert torch.equal(out.view(torch.int64), expect.view(torch.int64))
        # and once more on the settled weights (the ordinary path)
        again = r.regrid(values)
        assert torch.equal(again.view(torch.int64), out.view(torch.int64))


def test_resident_range_is_built_as_two_chained_jobs(xb, monkeypatch):
    """a large resident target range (a rank's shard, or a whole resident mesh) is split into two
    chained jobs so that regrid() can bring the first part down under the second job; results
    and out= semantics unchanged"""
    from xugrid_b200 import regridder

    sxy, sf = synthetic.jittered_lattice(150, 140, 51, "tri")
    txy, tf = synthetic.jittered_lattice(160, 130, 52, "mixed")
    src = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    tgt = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    values = torch.from_numpy(np.random.default_rng(10).random((2, len(sf)))).pin_memory()
    good = xb.OverlapRegridder(src, tgt, method="mean").regrid(values)
    monkeypatch.setattr(regridder.UnstructuredGrid2d, "_SPLIT_MIN_ROWS", 1000)
    n_t = len(tf)
    b, e = n_t // 5, n_t - n_t // 7
    for rng in (None, (b, e)):
        r = xb.OverlapRegridder(src, tgt, method="mean", target_range=rng)
        pending = r.__dict__["_pending"]
        assert len(pending.ranges) == 2 and pending.ranges[0][0] == 0 and pending.ranges[1][1] == pending.n
        expect = good if rng is None else good[:, b:e]
        out = r.regrid(values)
        assert torch.equal(out.view(torch.int64), expect.contiguous().view(torch.int64))
        # into a strided view of a larger host array (what a rank does with shared_host_array)
        r2 = xb.OverlapRegridder(src, tgt, method="mean", target_range=rng)
        big = torch.full((2, n_t + 10), -1.0, dtype=torch.float64).pin_memory()
        lo = 3 if rng is None else b
        view = big[:, lo: lo + expect.shape[1]]
        got = r2.regrid(values, out=view)
        assert got is view and torch.equal(view.contiguous().view(torch.int64), expect.contiguous().view(torch.int64))
        assert bool((big[:, :lo] == -1.0).all()) and bool((big[:, lo + expect.shape[1]:] == -1.0).all())
        assert r2._weights.nnz == r._weights.nnz


def test_job_as_one_cuda_graph(xb):
    """XG_RUN_GRAPH: same CSR; the first call may fall back to direct launches (scratch cache
    cold), the repeats go out as graphs"""
    sxy, sf, txy, tf = _meshes()
    dev = torch.device("cuda", 0)
    tree = xb.CellTree2d(sxy, sf)
    dxy, dtf = torch.from_numpy(txy).to(dev), torch.from_numpy(tf).to(dev)
    n_t = len(tf)
    ip0, ix0, ar0 = tree.intersect_faces_csr(dxy, dtf)
    as_graph = []
    for rep in range(4):
        indptr = torch.full((n_t + 1,), -7, dtype=torch.int64, device=dev)
        indices = torch.empty(16 * n_t, dtype=torch.int32, device=dev)
        area = torch.empty(16 * n_t, dtype=torch.float64, device=dev)
        cut = n_t // 3
        jobs = [tree.intersect_faces_into(dxy, dtf, indptr[: cut + 1], indices, area, t_begin=0, t_end=cut, graph=True),
                tree.intersect_faces_into(dxy, dtf, indptr[cut:], indices, area, t_begin=cut, t_end=n_t, chained=True, graph=True)]
        end = 0
        for j in jobs:
            _, end, ok = j.result()
            assert ok
            as_graph.append(j.launched_as_graph)
        assert end == int(ip0[-1]) and torch.equal(indptr, ip0) and torch.equal(indices[:end], ix0)
        assert torch.equal(area[:end].view(torch.int64), ar0.view(torch.int64))
    assert all(as_graph[-2:]), as_graph
