"""Host -> device transport of a mesh: 24-bit connectivity entries widened on the device
(xg_unpack_u24), in face ranges, for a face range only and completed later; the regridders on
top of it give the same weights as with resident int64 arrays."""
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


@pytest.mark.parametrize("kind", ["tri", "mixed"])
def test_packed_upload_reproduces_the_connectivity(xb, monkeypatch, kind):
    from xugrid_b200 import grid as grid_mod

    monkeypatch.setattr(grid_mod.Ugrid2d, "_PACK_MIN_FACES", 100)
    xy, faces = synthetic.jittered_lattice(37, 29, 3, kind)
    want = torch.from_numpy(faces.astype(np.int32))
    # whole mesh, one piece and several pieces (odd cut points)
    for chunks in (1, 3, (0.13, 0.61, 1.0)):
        g = xb.Ugrid2d(xy[:, 0], xy[:, 1], -1, faces)
        assert g._packs_u24()
        g.prefetch(n_chunks=chunks)
        assert g._faces_pin_u24 is not None and g._faces_pin_u24.numel() == faces.size * 3
        dxy, df = g.device_arrays()
        assert df.dtype == torch.int32 and torch.equal(df.cpu(), want)
        assert torch.equal(dxy.cpu(), torch.from_numpy(xy))
        assert g.upload_nbytes() == xy.size * 8 + faces.size * 3
    # a face range first (what a rank does with its own target rows), the rest on demand
    g = xb.Ugrid2d(xy[:, 0], xy[:, 1], -1, faces)
    b, e = 301, 1777
    g.prefetch(face_range=(b, e))
    _, df = g.device_arrays(face_range=(b, e))
    assert torch.equal(df[b:e].cpu(), want[b:e])
    _, df = g.device_arrays()
    assert torch.equal(df.cpu(), want)
    # release + upload again from the staged copy
    g.release_device()
    _, df = g.device_arrays()
    assert torch.equal(df.cpu(), want)


def test_unpack_edge_sizes(xb):
    import ctypes

    from xugrid_b200 import _native

    L = _native.lib()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 5, 1023, 4099):
        v = rng.integers(0, (1 << 24) - 1, n).astype(np.int64)
        v[rng.random(n) < 0.2] = -1
        v32 = v.astype(np.int32)
        packed = np.ascontiguousarray(v32.view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1) if n else np.zeros(0, np.uint8)
        d_packed = torch.from_numpy(np.concatenate([packed, np.zeros(8, np.uint8)])).to(dev)
        out = torch.full((n + 4,), 77, dtype=torch.int32, device=dev)
        _native.check(L.xg_unpack_u24(d_packed.data_ptr(), n, out.data_ptr(), _native.stream_ptr()))
        assert np.array_equal(out[:n].cpu().numpy(), v32)
        assert bool((out[n:] == 77).all())  # nothing written behind n
    # (the value 2^24 - 1 itself is the fill marker: meshes with that many nodes are not packed)


def test_regridder_on_packed_uploads_equals_resident(xb, monkeypatch):
    from xugrid_b200 import grid as grid_mod
    from xugrid_b200 import regridder

    sxy, sf = synthetic.jittered_lattice(120, 110, 21, "mixed")
    txy, tf = synthetic.jittered_lattice(131, 97, 22, "tri")
    values = torch.from_numpy(np.random.default_rng(5).random((2, len(sf)))).pin_memory()
    expect = xb.OverlapRegridder(xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf), xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf),
                                 method="mean").regrid(values)
    monkeypatch.setattr(grid_mod.Ugrid2d, "_PACK_MIN_FACES", 100)
    monkeypatch.setattr(regridder.BaseRegridder, "_CHUNK_MIN_FACES", 1000)
    src = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    tgt = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    for rep in range(2):  # second time: re-upload from the staged copies
        src.release_device()
        tgt.release_device()
        r = xb.OverlapRegridder(src, tgt, method="mean")
        assert src._packed_d is not None and tgt._packed_d is not None
        out = r.regrid(values)
        assert torch.equal(out.view(torch.int64), expect.view(torch.int64))
    b, e = len(tf) // 4, len(tf) - 100
    tgt.release_device()
    shard = xb.OverlapRegridder(src, tgt, method="mean", target_range=(b, e)).regrid(values)
    assert torch.equal(shard.view(torch.int64), expect[:, b:e].contiguous().view(torch.int64))
