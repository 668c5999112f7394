"""GPU tests of the regridder classes, modelled on the reference's
tests/test_regrid/test_regridder.py and test_unstructured.py, plus end-to-end
parity of BASELINE config 1 (5k-face Delaunay -> 50x50 raster) against the
oracle pipeline."""
import numpy as np
import pytest
import torch

import oracle
from tests import cases
from xugrid_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xugrid_b200

    assert torch.cuda.is_available()
    return xugrid_b200


@pytest.fixture(scope="module")
def disk(xb):
    xy, faces = synthetic.disk()
    grid = xb.Ugrid2d(xy[:, 0], xy[:, 1], -1, faces)
    cent = oracle.face_centroids(xy, faces)
    z = synthetic.function_z(cent[:, 0], cent[:, 1])
    return xb.UgridDataArray(z, grid)


def quads(xb, dx):
    # tests/fixtures/fixture_regridder.py:39-90
    xmin, ymin, xmax, ymax = -10.0, -10.0, 10.0, 10.0
    x = np.arange(xmin, xmax, dx) + 0.5 * dx
    y = np.arange(ymin, ymax, dx) + 0.5 * dx
    return xb.StructuredGrid2d(x, y)


def test_check_source_target_types(xb, disk):
    for cls in (xb.CentroidLocatorRegridder, xb.OverlapRegridder, xb.RelativeOverlapRegridder):
        with pytest.raises(TypeError):
            cls(source=disk, target=1.0)
        with pytest.raises(TypeError):
            cls(source=1.0, target=disk)


def test_overlap_regridder(xb, disk):
    # test_regridder.py:152-167
    square = quads(xb, 2.0).to_ugrid2d()
    assert square.n_face == 100
    regridder = xb.OverlapRegridder(disk, square, method="mean")
    result = regridder.regrid(disk)
    assert isinstance(result, xb.UgridDataArray)
    v = result.values
    assert np.isfinite(v).any()
    assert np.nanmin(v) >= disk.values.min() and np.nanmax(v) <= disk.values.max()
    layered = xb.UgridDataArray(np.ones(5)[:, None] * disk.values[None, :], disk.grid, ("layer", disk.grid.face_dimension))
    broadcasted = regridder.regrid(layered)
    assert broadcasted.dims == ("layer", square.face_dimension)
    assert broadcasted.shape == (5, 100)
    assert np.array_equal(broadcasted.values[3], v, equal_nan=True)


def test_centroid_locator_regridder(xb, disk):
    # test_regridder.py:100-127
    square = quads(xb, 2.0).to_ugrid2d()
    regridder = xb.CentroidLocatorRegridder(source=disk, target=square)
    result = regridder.regrid(disk)
    v = result.values
    assert np.isfinite(v).any()
    assert np.nanmin(v) >= disk.values.min() and np.nanmax(v) <= disk.values.max()
    # parity with the oracle's locate
    cent = oracle.face_centroids(square.node_coordinates, square.face_node_connectivity)
    idx = oracle.CellTree2d(disk.grid.node_coordinates, disk.grid.face_node_connectivity).locate_points(cent)
    expect = np.where(idx >= 0, disk.values[idx.clip(0)], np.nan)
    assert np.array_equal(v, expect, equal_nan=True)
    back = xb.CentroidLocatorRegridder(source=result, target=disk).regrid(result)
    assert back.grid.n_face == disk.grid.n_face


@pytest.mark.parametrize("method", ["mean", "sum", "minimum", "maximum", "mode", "median", "max_overlap",
                                    "harmonic_mean", "geometric_mean", "p25"])
def test_config1_delaunay_to_raster_matches_oracle(xb, method):
    """BASELINE.json configs[0]: 5k-face Delaunay -> 50x50 structured target."""
    sxy = np.random.default_rng(1).random((2600, 2))
    from scipy.spatial import Delaunay

    sf = Delaunay(sxy).simplices.astype(np.int64)
    source = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    cx = np.linspace(0.01, 0.99, 50)
    target = xb.StructuredGrid2d(cx, cx)
    values = np.random.default_rng(2).random(len(sf))
    values[::17] = np.nan
    out = xb.OverlapRegridder(source, target, method=method).regrid(values)
    assert out.shape == (50, 50)
    # oracle pipeline: promote raster, intersect, to_csr, reduce
    tgrid = target.to_ugrid2d()
    si, ti, w = oracle.overlap_triplets(sxy, sf, tgrid.node_coordinates, tgrid.face_node_connectivity)
    o = cases.canon(ti, si, w)
    indptr = oracle.coo_to_csr(o[0], 2500)
    ref = oracle.csr_apply(method, values, o[2], o[1], indptr, 2500).reshape(50, 50)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    if method == "geometric_mean":
        assert np.allclose(out, ref, rtol=1e-10, atol=0, equal_nan=True)
    else:
        assert np.array_equal(out, ref, equal_nan=True)


def test_relative_overlap_conservative(xb):
    sxy, sf = synthetic.jittered_lattice(30, 30, 1, "mixed")
    txy, tf = synthetic.jittered_lattice(11, 12, 2, "tri", bounds=(0.1, 0.9, 0.1, 0.9))
    source = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    values = np.random.default_rng(3).random((4, len(sf)))
    r = xb.RelativeOverlapRegridder(source, target)
    out = r.regrid(values)
    si, ti, w = oracle.overlap_triplets(sxy, sf, txy, tf, relative=True)
    o = cases.canon(ti, si, w)
    ref = oracle.csr_apply("first_order_conservative", values, o[2], o[1], oracle.coo_to_csr(o[0], len(tf)), len(tf))
    assert np.array_equal(out, ref, equal_nan=True)
    with pytest.raises(ValueError):
        xb.RelativeOverlapRegridder(source, target, method="mean")


def test_method_errors(xb, disk):
    # regridder.py:124-141
    square = quads(xb, 4.0)
    with pytest.raises(ValueError, match="Invalid regridding method"):
        xb.OverlapRegridder(disk, square, method="nope")
    with pytest.raises(TypeError):
        xb.OverlapRegridder(disk, square, method=3.5)
    with pytest.raises(NotImplementedError):
        xb.OverlapRegridder(disk, square, method=lambda v, w, ws: 0.0)
    r = xb.OverlapRegridder(disk, square, method=xb.OverlapRegridder.create_percentile_method(50))
    m = xb.OverlapRegridder(disk, square, method="median")
    assert np.array_equal(r.regrid(disk), m.regrid(disk), equal_nan=True)
    with pytest.raises(TypeError):
        r.regrid([1, 2, 3])
    with pytest.raises(ValueError):
        r.regrid(np.zeros(7))
    median = xb.OverlapRegridder.create_percentile_method(percentile=50)
    assert median(np.array([0, 1, 2, 3, 4], dtype=float), np.ones(5), np.zeros(5)) == 2


@pytest.mark.parametrize("cls_name", ["CentroidLocatorRegridder", "OverlapRegridder", "RelativeOverlapRegridder"])
def test_regridder_from_weights_and_dataset(xb, disk, cls_name):
    # test_regridder.py:203-257
    cls = getattr(xb, cls_name)
    square = quads(xb, 2.0).to_ugrid2d()
    regridder = cls(source=disk, target=square)
    result = regridder.regrid(disk)
    weights = regridder.weights
    assert weights["__regrid_n"] == 100 and weights["__source_type"] == "UnstructuredGrid2d"
    new = cls.from_weights(weights, target=square)
    assert np.array_equal(new.regrid(disk).values, result.values, equal_nan=True)
    layered = np.stack([disk.values * k for k in (1.0, 2.0, 3.0)])
    assert np.array_equal(new.regrid(layered)[0], result.values, equal_nan=True)
    again = cls.from_dataset(regridder.to_dataset())
    assert np.array_equal(again.regrid(disk).values, result.values, equal_nan=True)
    df = regridder.weights_as_dataframe()
    assert {"source_index", "target_index", "weight"} <= set(df.columns)
    regridder._weights = None
    with pytest.raises(ValueError):
        regridder.weights_as_dataframe()


def test_structured_source_to_unstructured_target(xb, disk):
    # test_regridder.py:41-55
    src = quads(xb, 1.0)
    data = np.ones(src.shape)
    for cls in (xb.CentroidLocatorRegridder, xb.OverlapRegridder, xb.RelativeOverlapRegridder):
        out = cls(src, disk).regrid(data)
        assert out.shape == (disk.grid.n_face,)
        assert np.nanmax(out) <= 1.0 + 1e-12


def test_self_overlap_locate_identity(xb, disk):
    # tests/test_regrid/test_unstructured.py:32-59
    g = disk.grid
    tree = g.celltree
    ti, si, w = tree.intersect_faces(g.node_coordinates, g.face_node_connectivity)
    valid = w > 1e-5
    assert np.array_equal(ti[valid], np.arange(g.n_face)) and np.array_equal(si[valid], np.arange(g.n_face))
    assert np.allclose(w[valid], g.area)
    assert np.array_equal(g.locate_points(g.centroids), np.arange(g.n_face))


def test_torch_tensor_in_tensor_out_and_target_range(xb):
    sxy, sf = synthetic.delaunay(1500, 3)
    txy, tf = synthetic.delaunay(1200, 4)
    source = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    v = torch.rand(3, len(sf), dtype=torch.float64, device="cuda")
    full = xb.OverlapRegridder(source, target).regrid(v)
    assert isinstance(full, torch.Tensor) and full.is_cuda and full.shape == (3, len(tf))
    parts = []
    from xugrid_b200.distributed import shard_range

    for r in range(3):
        b, e = shard_range(len(tf), r, 3)
        parts.append(xb.OverlapRegridder(source, target, target_range=(b, e)).regrid(v))
    assert torch.equal(torch.cat(parts, dim=1).nan_to_num(-1.0), full.nan_to_num(-1.0))


# ---------------------------------------------------------------------------- barycentric
@pytest.mark.parametrize("name", ["disk", "mixed"])
def test_barycentric_weights_match_reference_pipeline(xb, golden_dir, name):
    """BarycentricInterpolator weights against the golden triplets produced with the
    reference's voronoi.py + replace_interpolated_weights (oracle/make_golden.py)."""
    g = np.load(f"{golden_dir}/barycentric_golden.npz")
    xy, faces, points = g[f"{name}_xy"], g[f"{name}_faces"], g[f"{name}_points"]
    source = xb.Ugrid2d(xy[:, 0], xy[:, 1], -1, faces)
    # target mesh whose face centroids are exactly the golden points: tiny triangles
    d = 1e-3
    tri = np.array([[-d, -d], [2 * d, -d], [-d, 2 * d]])  # centroid at the origin
    txy = (points[:, None, :] + tri[None, :, :]).reshape(-1, 2)
    tf = np.arange(3 * len(points)).reshape(-1, 3)
    target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    assert np.allclose(oracle.face_centroids(txy, tf), points, atol=1e-15)
    r = xb.BarycentricInterpolator(source, target)
    data, indices, indptr = r._weights.to_numpy()
    tgt = np.repeat(np.arange(len(points)), np.diff(indptr))
    order = np.lexsort((data, indices, tgt))
    # centroids of the tiny triangles differ from the golden points by rounding (1e-15), and
    # Wachspress weights near a polygon edge amplify that: the pair set is identical, weights
    # agree to 1e-6 here (the exact-point comparison at 1e-10 is tests/test_voronoi_cpu.py and
    # test_barycentric_parity)
    assert np.array_equal(tgt[order], g[f"{name}_target"])
    assert np.array_equal(indices[order], g[f"{name}_source"])
    assert np.allclose(data[order], g[f"{name}_weight"], rtol=1e-6, atol=1e-9)


def test_barycentric_interpolator(xb, disk):
    # test_regridder.py:187-201: 0.25-spaced quads over the disk
    square = quads(xb, 0.25).to_ugrid2d()
    regridder = xb.BarycentricInterpolator(source=disk, target=square)
    result = regridder.regrid(disk).values
    assert np.isfinite(result).any()
    assert np.nanmin(result) >= disk.values.min() and np.nanmax(result) <= disk.values.max()
    layered = xb.UgridDataArray(np.ones(5)[:, None] * disk.values[None, :], disk.grid, ("layer", disk.grid.face_dimension))
    broadcasted = regridder.regrid(layered)
    assert broadcasted.dims == ("layer", square.face_dimension) and broadcasted.shape == (5, 6400)
    # self-interpolation is the identity (test_unstructured.py:53-59)
    me = xb.BarycentricInterpolator(disk, disk)
    data, indices, indptr = me._weights.to_numpy()
    strong = data > 1e-9
    rows = np.repeat(np.arange(disk.grid.n_face), np.diff(indptr))
    assert np.array_equal(rows[strong], np.arange(disk.grid.n_face))
    assert np.array_equal(indices[strong], np.arange(disk.grid.n_face))
    assert np.allclose(data[strong], 1.0)
    # round trips like the other regridders
    again = xb.BarycentricInterpolator.from_weights(regridder.weights, target=square)
    assert np.array_equal(again.regrid(disk).values, result, equal_nan=True)


def test_barycentric_reproduces_linear_fields(xb):
    sxy, sf = synthetic.jittered_lattice(40, 40, 5, "mixed")
    txy, tf = synthetic.jittered_lattice(23, 19, 6, "tri", bounds=(0.1, 0.9, 0.1, 0.9))
    source = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    f = lambda p: 2.0 + 3.0 * p[:, 0] - 1.5 * p[:, 1]
    out = xb.BarycentricInterpolator(source, target).regrid(f(source.centroids))
    assert np.allclose(out, f(target.centroids), rtol=0, atol=1e-12)


def test_barycentric_concave_known_answer(xb):
    # /root/reference/tests/test_regrid/test_regridder.py:334-369
    vertices = np.array([[0.0, 0.0], [3.0, 0.0], [1.0, 1.0], [0.0, 2.0], [3.0, 2.0]])
    faces = np.array([[0, 1, 2], [0, 2, 3], [2, 4, 3]])
    grid = xb.Ugrid2d(vertices[:, 0], vertices[:, 1], -1, faces)
    dx = 0.1
    x = np.arange(0.0, 3.0, dx) + 0.5 * dx
    y = np.arange(0.0, 2.0, dx) + 0.5 * dx
    other = xb.StructuredGrid2d(x, y)
    result = xb.BarycentricInterpolator(source=grid, target=other).regrid(np.array([2.0, 0.5, 2.0]))
    assert result.shape == (20, 30)
    assert np.nanmin(result) >= 0.5 and np.nanmax(result) <= 2.0
    # There should be a large enough "gap" of NaNs on the right side.
    assert np.isnan(result).sum() == 200


def test_barycentric_structured_source_known_answer(xb):
    # /root/reference/tests/test_regrid/test_regridder.py:371-405 (the reference takes its
    # separable linear path for a structured source; here the raster is a quad mesh and the
    # same numbers come out of the Voronoi / Wachspress path)
    source = xb.StructuredGrid2d(x=np.array([0.5, 1.5]), y=np.array([1.5, 0.5]))
    data = np.array([[1.0, 2.0], [3.0, 4.0]])
    c = np.arange(0.0, 2.0, 0.25) + 0.125
    target = xb.StructuredGrid2d(c, c)
    out = xb.BarycentricInterpolator(source, target).regrid(data)
    assert out.shape == (8, 8)
    assert np.allclose(out[0], [3.0, 3.0, 3.125, 3.375, 3.625, 3.875, 4.0, 4.0])
    assert np.allclose(out[-1], [1.0, 1.0, 1.125, 1.375, 1.625, 1.875, 2.0, 2.0])
    assert np.allclose(out[:, 0], [3.0, 3.0, 2.75, 2.25, 1.75, 1.25, 1.0, 1.0])


def test_chunked_target_upload_gives_the_same_weights(xb):
    """Targets of >= 2^20 faces are uploaded in face ranges and regridded range by
    range while the next range is in flight; the CSR must not depend on that."""
    sxy, sf = synthetic.jittered_lattice(300, 300, 41, "tri")
    txy, tf = synthetic.jittered_lattice(740, 745, 42, "tri")
    assert len(tf) >= 1 << 20
    source = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    values = np.random.default_rng(5).random(len(sf))
    r4 = xb.OverlapRegridder(source, target)
    assert target._chunks is None or len(target._chunks) <= 4
    out4 = r4.regrid(values)
    w4 = r4._weights.to_numpy()
    target.release_device()
    source.release_device()
    old = xb.OverlapRegridder._TARGET_CHUNKS
    xb.OverlapRegridder._TARGET_CHUNKS = 1
    try:
        r1 = xb.OverlapRegridder(source, target)
    finally:
        xb.OverlapRegridder._TARGET_CHUNKS = old
    w1 = r1._weights.to_numpy()
    for a, b in zip(w4, w1):
        assert np.array_equal(a, b)
    assert np.array_equal(out4, r1.regrid(values), equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("flip_x,flip_y", [(False, False), (True, False), (False, True)])
def test_raster_target_as_points_equals_promoted_mesh(xb, flip_x, flip_y):
    """CentroidLocatorRegridder / BarycentricInterpolator onto a raster use only its cell
    centres, laid out on the device; same points and weights as the quad mesh the
    reference promotes the raster to (structured.py:489-501)."""
    from xugrid_b200 import synthetic
    from xugrid_b200.regridder import _StructuredAdapter

    sxy, sf = synthetic.delaunay(1500, 3)
    source = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf)
    x = np.linspace(0.05, 0.95, 37)
    y = np.linspace(0.1, 0.9, 23)
    x = x[::-1].copy() if flip_x else x
    y = y[::-1].copy() if flip_y else y
    raster = xb.StructuredGrid2d(x, y)
    adapter = _StructuredAdapter(raster)
    pts = adapter.centroids_device(torch.device("cuda", 0)).cpu().numpy()
    mesh = raster.to_ugrid2d()
    assert np.array_equal(pts, mesh.centroids)
    values = np.random.default_rng(5).random(len(sf))
    for cls in (xb.CentroidLocatorRegridder, xb.BarycentricInterpolator):
        a = cls(source, raster).regrid(values)
        b = cls(source, mesh).regrid(values)
        assert a.shape == (23, 37)
        assert np.array_equal(a.ravel(), np.asarray(b).ravel(), equal_nan=True)


def _random_network(n_edge, seed, lo=-0.1, hi=1.1):
    rng = np.random.default_rng(seed)
    a = rng.uniform(lo, hi, (n_edge, 2))
    b = a + rng.normal(0.0, 0.08, (n_edge, 2))
    return np.stack([a, b], axis=1)


@pytest.mark.parametrize("mesh", ["delaunay", "mixed", "hexagons"])
def test_intersect_edges_parity(xb, mesh):
    """CellTree2d.intersect_edges against the oracle: same (edge, face) pairs, clipped
    segments bit-identical (same Cyrus-Beck arithmetic, -fmad=false)"""
    if mesh == "delaunay":
        xy, f = synthetic.delaunay(3000, 61)
        edges = _random_network(4000, 62)
    elif mesh == "mixed":
        xy, f = synthetic.jittered_lattice(30, 25, 63, "mixed")
        edges = _random_network(3000, 64)
    else:
        xy, f = cases.hexagon_mesh(6, 5, 0)
        edges = _random_network(1500, 65, -1.0, 9.0)
    # also: degenerate (zero-length), NaN and far-away segments
    edges[7] = edges[7][[0, 0]]
    edges[11, 0, 0] = np.nan
    edges[13] += 100.0
    gi, gf, gxy = xb.CellTree2d(xy, f).intersect_edges(edges)
    oi, of, oxy = oracle.CellTree2d(xy, f).intersect_edges(edges)
    assert len(gi) > len(edges) // 2
    assert np.array_equal(gi, oi) and np.array_equal(gf, of)
    assert np.array_equal(gxy, oxy)


def test_intersect_edges_arena_regrows(xb):
    """long segments crossing many faces: more pairs than the first arena holds"""
    xy, f = synthetic.jittered_lattice(60, 60, 66, "tri")
    rng = np.random.default_rng(67)
    edges = np.stack([rng.uniform(0, 0.1, (300, 2)), rng.uniform(0.9, 1.0, (300, 2))], axis=1)
    gi, gf, gxy = xb.CellTree2d(xy, f).intersect_edges(edges)
    oi, of, oxy = oracle.CellTree2d(xy, f).intersect_edges(edges)
    assert len(gi) > 4 * len(edges) + 1024
    assert np.array_equal(gi, oi) and np.array_equal(gf, of) and np.array_equal(gxy, oxy)


@pytest.mark.parametrize("structured", [False, True])
def test_network_gridder_reference_known_answers(xb, structured):
    # /root/reference/tests/test_regrid/test_network_gridder.py:10-140
    raster = xb.StructuredGrid2d(np.arange(0.5, 4.5, 1.0), np.arange(3.5, -0.5, -1.0))
    target = raster if structured else raster.to_ugrid2d()
    node_xy = np.array([[0.0, 0.0], [1.5, 1.5], [2.5, 1.5], [4.0, 0.0], [4.0, 3.0]])
    edge_nodes = np.array([[0, 1], [1, 2], [2, 3], [2, 4]])
    network = xb.Ugrid1d(node_xy[:, 0], node_xy[:, 1], -1, edge_nodes)
    data = np.array([1.0, 2.0, 4.0, -4.0])
    gridder = xb.NetworkGridder(network, target, method="mean")
    assert gridder._weights.n == 16 and gridder._weights.m == 4 and gridder._weights.nnz == 8
    out = np.asarray(gridder.regrid(data))
    assert out.shape == ((4, 4) if structured else (16,))
    assert np.count_nonzero(np.isnan(out)) == 11
    diag = 0.5 * np.sqrt(2)
    expected = np.array([1.0, (diag * 1 + 0.5 * 2) / (diag + 0.5),
                         (0.5 * 2 + diag * -4 + diag * 4) / (2 * diag + 0.5), -4.0, 4.0])
    # rows of the raster run from y = 3.5 down to 0.5
    cells = [(3, 0), (2, 1), (2, 2), (1, 3), (3, 3)]
    got = np.array([out.reshape(4, 4)[r, c] for r, c in cells])
    assert np.allclose(got, expected)
    # transient data (test_network_gridder.py:97-120) and the weights round trip
    both = np.asarray(gridder.regrid(np.stack([data, 2.0 * data])))
    assert np.count_nonzero(np.isnan(both)) == 22
    assert np.allclose(both.reshape(2, 4, 4)[1][tuple(zip(*cells))], 2.0 * expected)
    again = xb.NetworkGridder.from_weights(gridder.weights, target, method="mean")
    assert np.array_equal(np.asarray(again.regrid(data)), out, equal_nan=True)
    with pytest.raises(TypeError):
        xb.NetworkGridder(raster.to_ugrid2d(), target)


def test_target_range_of_the_point_regridders(xb, disk):
    """CentroidLocatorRegridder / BarycentricInterpolator built for a range of target faces
    give exactly the columns of the full regridder (rows are independent)."""
    txy, tf = synthetic.jittered_lattice(23, 19, 61, "mixed", bounds=(-9.0, 9.0, -9.0, 9.0))
    target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    raster = quads(xb, 0.5)
    values = np.stack([disk.values, disk.values[::-1]])
    for cls in (xb.CentroidLocatorRegridder, xb.BarycentricInterpolator):
        for tgt in (target, raster):
            full = cls(disk, tgt).regrid(values).reshape(2, -1)
            n = full.shape[1]
            for b, e in ((0, n // 3), (n // 3, n), (5, 5)):
                part = cls(disk, tgt, target_range=(b, e)).regrid(values)
                assert part.shape == (2, e - b)
                assert np.array_equal(part, full[:, b:e], equal_nan=True), (cls.__name__, b, e)


def test_curve_shards_partition_the_mesh_into_compact_patches(xb):
    """distributed.curve_shard: disjoint, complete, balanced, and compact although the
    faces are numbered at random"""
    from xugrid_b200.distributed import curve_shard

    txy, tf = synthetic.jittered_lattice(90, 80, 71, "tri")
    tf = np.ascontiguousarray(tf[np.random.default_rng(7).permutation(len(tf))])
    dxy, dtf = torch.from_numpy(txy).cuda(), torch.from_numpy(tf).cuda()
    for world in (1, 3, 8):
        shards = [curve_shard(dxy, dtf, r, world).cpu().numpy() for r in range(world)]
        allids = np.concatenate(shards)
        assert np.array_equal(np.sort(allids), np.arange(len(tf)))
        sizes = np.array([len(s) for s in shards])
        assert sizes.max() - sizes.min() <= 0.02 * len(tf) + 64
        if world > 1:
            # bounding boxes of the patches: together not much more than the domain
            area = 0.0
            for s in shards:
                p = txy[tf[s]].reshape(-1, 2)
                area += (p[:, 0].max() - p[:, 0].min()) * (p[:, 1].max() - p[:, 1].min())
            assert area < 2.0  # unit square; random numbering with range shards would give ~ world


def test_target_ids_give_the_columns_of_the_full_regridder(xb, disk):
    from xugrid_b200.distributed import curve_shard

    txy, tf = synthetic.jittered_lattice(31, 27, 73, "mixed", bounds=(-9.0, 9.0, -9.0, 9.0))
    tf = np.ascontiguousarray(tf[np.random.default_rng(9).permutation(len(tf))])
    target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
    values = np.stack([disk.values, disk.values[::-1]])
    dxy, dtf = target.device_arrays()
    for cls in (xb.OverlapRegridder, xb.RelativeOverlapRegridder, xb.CentroidLocatorRegridder,
                xb.BarycentricInterpolator):
        full = cls(disk, target).regrid(values)
        seen = np.zeros(len(tf), dtype=bool)
        for r in range(3):
            ids = curve_shard(dxy, dtf, r, 3)
            part = cls(disk, target, target_ids=ids).regrid(values)
            idn = ids.cpu().numpy()
            assert part.shape == (2, len(idn))
            assert np.array_equal(part, full[:, idn], equal_nan=True), (cls.__name__, r)
            seen[idn] = True
        assert seen.all()
    empty = xb.OverlapRegridder(disk, target, target_ids=torch.empty(0, dtype=torch.int32, device="cuda"))
    assert empty.regrid(values).shape == (2, 0)
