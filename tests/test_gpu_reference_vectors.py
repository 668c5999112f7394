"""GPU: the overlap / centroid / linear vectors the reference's own tests hold
(/root/reference/tests/test_regrid/test_structured.py:51-205,249-262,
tests/fixtures/fixture_regridder.py:240-362, consumed by
tests/test_regrid/test_regridder.py:100-150,371-405), reproduced

* through the CUDA overlap kernels (k_traverse -> k_sat -> k_clip_survivors): the
  structured fixtures a..e promoted to QUAD MESHES as StructuredGrid2d.convert_to does
  (structured.py:489-501) -- the only reference-held vectors that pin intersect_faces;
* through the regridder classes on the rasters themselves (the reference's separable
  raster -> raster path, xugrid_b200/structured.py, applied by the CUDA CSR kernels), with
  full equality instead of the reference's weak ``.any()``.

tests/test_reference_vectors_cpu.py asserts the same vectors on the oracle."""
import numpy as np
import pytest
import torch

import oracle
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xugrid_b200

    assert torch.cuda.is_available()
    return xugrid_b200


def quad_mesh(xb, **kw):
    return xb.StructuredGrid2d(**kw).to_ugrid2d()


def gpu_overlap(xb, src_kw, tgt_kw):
    s, t = quad_mesh(xb, **src_kw), quad_mesh(xb, **tgt_kw)
    ti, si, area = s.celltree.intersect_faces(t.node_coordinates, t.face_node_connectivity)
    oti, osi, oarea = oracle.CellTree2d(s.node_coordinates, s.face_node_connectivity).intersect_faces(
        t.node_coordinates, t.face_node_connectivity)
    # and bit-identical to the oracle, as everywhere else
    assert cases.pair_set(si, ti, area) == cases.pair_set(osi, oti, oarea)
    return si, ti, area, s


def test_overlap_2d_is_the_reference_answer(xb):
    si, ti, area, _ = gpu_overlap(xb, cases.FIX_A, cases.FIX_B)
    assert cases.pair_set(si, ti, area) == cases.pair_set(cases.OVERLAP_2D_SOURCE, cases.OVERLAP_2D_TARGET,
                                                          np.full(36, 625.0))


@pytest.mark.parametrize("name", list(cases.OVERLAP_1D))
def test_overlap_1d_as_quad_strips(xb, name):
    strip = cases.strip_axis_y()
    src = dict(x=cases.FIX_A["x"], **strip)
    if name == "a_b":
        tgt = dict(x=cases.FIX_B["x"], **strip)
    elif name == "a_b_flipped":
        tgt = dict(x=cases.FIX_B["y"], **strip)
    else:
        tgt = dict(x=cases.FIX_E_X, xbounds=cases.FIX_E_XBOUNDS, **strip)
    es, et, ew = cases.OVERLAP_1D[name]
    if name.endswith("relative"):
        # through RelativeOverlapRegridder's weights on the quad strips (unstructured.py:133-134)
        r = xb.RelativeOverlapRegridder(quad_mesh(xb, **src), quad_mesh(xb, **tgt))
        data, col, indptr = r._weights.to_numpy()
        row = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
        assert cases.pair_set(col, row, data) == cases.pair_set(es, et, ew)
    else:
        si, ti, area, _ = gpu_overlap(xb, src, tgt)
        assert cases.pair_set(si, ti, area) == cases.pair_set(es, et, ew)


def test_centroids_2d_is_the_reference_answer(xb):
    s, t = quad_mesh(xb, **cases.FIX_A), quad_mesh(xb, **cases.FIX_B)
    idx = s.celltree.locate_points(t.centroids)
    # corner shared by four source cells -> smallest index, the reference's pairs
    assert np.array_equal(idx[cases.CENTROID_2D_TARGET], cases.CENTROID_2D_SOURCE)
    cent = t.centroids
    outside = (cent[:, 0] < 25) | (cent[:, 0] > 175) | (cent[:, 1] < 25) | (cent[:, 1] > 175)
    assert np.all(idx[outside] == -1)


# ---- regridder classes on the rasters (test_regridder.py:100-150) ---------------------------------
def rasters(xb):
    return xb.StructuredGrid2d(**cases.FIX_A), xb.StructuredGrid2d(**cases.FIX_B)


def test_overlap_regridder_structured(xb):
    a, b = rasters(xb)
    regridder = xb.OverlapRegridder(source=a, target=b)
    result = regridder.regrid(cases.FIX_A_DATA)
    assert result.shape == (4, 4)
    assert np.array_equal(result, cases.EXPECTED_OVERLAP)
    broadcasted = regridder.regrid(cases.FIX_A_LAYERED)
    assert broadcasted.shape == (2, 4, 4)
    assert np.array_equal(broadcasted[0], cases.EXPECTED_OVERLAP)
    assert np.array_equal(broadcasted[1], cases.EXPECTED_OVERLAP + 9.0)
    # weights: the 36 pairs, 625.0 each, as CSR
    data, col, indptr = regridder._weights.to_numpy()
    row = np.repeat(np.arange(16), np.diff(indptr))
    assert cases.pair_set(col, row, data) == cases.pair_set(cases.OVERLAP_2D_SOURCE, cases.OVERLAP_2D_TARGET,
                                                            np.full(36, 625.0))
    # the same rasters as quad meshes through the overlap kernels: same values
    quads = xb.OverlapRegridder(a.to_ugrid2d(), b.to_ugrid2d()).regrid(cases.FIX_A_DATA.ravel())
    assert np.array_equal(np.asarray(quads).reshape(4, 4), cases.EXPECTED_OVERLAP)
    # "mode" must not repeat the first value (xugrid issue 236, test_regridder.py:143-150)
    adapted = cases.FIX_A_DATA.copy()
    adapted[0, 0] = 99
    result = xb.OverlapRegridder(source=a, target=a, method="mode").regrid(adapted)
    assert not np.all(result == 99.0)
    assert np.array_equal(result, adapted)


def test_centroid_locator_regridder_structured(xb):
    a, b = rasters(xb)
    regridder = xb.CentroidLocatorRegridder(source=a, target=b)
    result = regridder.regrid(cases.FIX_A_DATA)
    assert np.array_equal(result, cases.EXPECTED_CENTROID, equal_nan=True)
    broadcasted = regridder.regrid(cases.FIX_A_LAYERED)
    assert broadcasted.shape == (2, 4, 4)
    assert np.array_equal(broadcasted[0], cases.EXPECTED_CENTROID, equal_nan=True)
    data, row, col = regridder._weights.to_numpy()
    assert cases.pair_set(col, row, data) == cases.pair_set(cases.CENTROID_2D_SOURCE, cases.CENTROID_2D_TARGET,
                                                            np.ones(4))


def test_barycentric_interpolator_structured(xb):
    a, b = rasters(xb)
    regridder = xb.BarycentricInterpolator(source=a, target=b)
    result = regridder.regrid(cases.FIX_A_DATA)
    assert np.array_equal(result, cases.EXPECTED_LINEAR, equal_nan=True)
    broadcasted = regridder.regrid(cases.FIX_A_LAYERED)
    assert np.array_equal(broadcasted[0], cases.EXPECTED_LINEAR, equal_nan=True)


def test_structured_weights_round_trip(xb):
    a, b = rasters(xb)
    for cls in (xb.OverlapRegridder, xb.RelativeOverlapRegridder, xb.CentroidLocatorRegridder,
                xb.BarycentricInterpolator):
        r = cls(source=a, target=b)
        ds = r.to_dataset()
        assert ds["__source_type"] == "StructuredGrid2d" and ds["__target_type"] == "StructuredGrid2d"
        again = cls.from_dataset(ds)
        assert np.array_equal(again.regrid(cases.FIX_A_DATA), r.regrid(cases.FIX_A_DATA), equal_nan=True)


def test_structured_golden_through_the_apply_kernels(xb, golden_dir):
    """every golden raster pair (weights from the reference's structured.py): the CSR built on
    the device gives the values the reference's triplets give through the oracle's apply"""
    import os

    g = np.load(os.path.join(golden_dir, "structured_golden.npz"))
    for name in g["names"]:
        def grid(key):
            kw = {ax + "bounds": g[f"{name}_{key}{ax}bounds"] for ax in ("x", "y") if f"{name}_{key}{ax}bounds" in g.files}
            return xb.StructuredGrid2d(g[f"{name}_{key}x"], g[f"{name}_{key}y"], **kw)
        src, tgt = grid("s"), grid("t")
        values = np.random.default_rng(3).random((2,) + src.shape)
        values[1, ::3, ::2] = np.nan
        for kind, cls in (("overlap", xb.OverlapRegridder), ("relative", xb.RelativeOverlapRegridder),
                          ("linear", xb.BarycentricInterpolator)):
            if f"{name}_{kind}_weight" not in g.files:
                continue
            si, ti, w = g[f"{name}_{kind}_source"], g[f"{name}_{kind}_target"], g[f"{name}_{kind}_weight"]
            method = "first_order_conservative" if kind == "relative" else "mean"
            ref = oracle.csr_apply(method, values.reshape(2, -1), w, si, oracle.coo_to_csr(ti, tgt.size), tgt.size)
            out = cls(src, tgt).regrid(values)
            # same weights; entries of a row may be summed in another order (last-bit differences)
            assert np.allclose(out.reshape(2, -1), ref, rtol=1e-12, atol=0, equal_nan=True), (name, kind)


# ---- sel_points (/root/reference/tests/test_ugrid2d.py:826-890) ------------------------------------
def test_sel_points(xb):
    grid = xb.Ugrid2d(cases.VERTICES[:, 0], cases.VERTICES[:, 1], -1, cases.FACES)
    obj = np.array([0, 1, 2, 3])
    x, y = [0.5, 1.5], [0.5, 1.25]
    with pytest.raises(ValueError, match="method must be one of"):
        grid.sel_points(obj=obj, x=x, y=y, method="nothing")
    with pytest.raises(ValueError, match="out_of_bounds must be one of"):
        grid.sel_points(obj=obj, x=x, y=y, out_of_bounds="nothing")
    with pytest.raises(ValueError, match="shape of x does not match shape of y"):
        grid.sel_points(obj=obj, x=[0.5, 1.5], y=[0.5])
    with pytest.raises(ValueError, match="x and y must be 1d"):
        grid.sel_points(obj=obj, x=[x], y=[y])
    actual = grid.sel_points(obj=obj, x=x, y=y)
    assert np.array_equal(np.asarray(actual), [0, 3])
    assert actual.dim == "mesh2d_points"
    assert np.array_equal(actual.x, x) and np.array_equal(actual.y, y)


def test_sel_points_out_of_bounds(xb):
    grid = xb.Ugrid2d(cases.VERTICES[:, 0], cases.VERTICES[:, 1], -1, cases.FACES)
    obj = np.array([0, 1, 2, 3])
    x = [-10.0, 0.5, -20.0, 1.5, -30.0]
    y = [-10.0, 0.5, -20.0, 1.25, -30.0]
    with pytest.raises(ValueError, match="Not all points are located on the topology"):
        grid.sel_points(obj=obj, x=x, y=y, out_of_bounds="raise")
    dropped = grid.sel_points(obj=obj, x=x, y=y, out_of_bounds="drop")
    assert np.array_equal(np.asarray(dropped), [0, 3]) and np.array_equal(dropped.index, [1, 3])
    with pytest.warns(UserWarning, match="Not all points are located on the topology"):
        actual = grid.sel_points(obj=obj, x=x, y=y, out_of_bounds="warn")
    assert np.allclose(np.asarray(actual), [np.nan, 0, np.nan, 3, np.nan], equal_nan=True)
    actual = grid.sel_points(obj=obj, x=x, y=y, out_of_bounds="ignore")
    assert np.allclose(np.asarray(actual), [np.nan, 0, np.nan, 3, np.nan], equal_nan=True)
    actual = grid.sel_points(obj=obj, x=x, y=y, out_of_bounds="ignore", fill_value=-1)
    assert np.allclose(np.asarray(actual), [-1, 0, -1, 3, -1])
    actual = grid.sel_points(obj=obj, x=x, y=y, out_of_bounds="drop", tolerance=11.0)
    assert np.asarray(actual).shape == (2,)
    # layered data and device tensors keep their leading axes
    layered = np.arange(8.0).reshape(2, 4)
    sel = grid.sel_points(obj=layered, x=x, y=y, out_of_bounds="ignore")
    assert np.allclose(np.asarray(sel), [[np.nan, 0, np.nan, 3, np.nan], [np.nan, 4, np.nan, 7, np.nan]], equal_nan=True)
    sel_t = grid.sel_points(obj=torch.from_numpy(layered).cuda(), x=x, y=y, out_of_bounds="ignore")
    assert sel_t.values.is_cuda and np.allclose(np.asarray(sel_t), np.asarray(sel), equal_nan=True)


def test_weights_validation_raises_instead_of_reading_out_of_bounds(xb):
    """a stale / mismatched weights dict must raise like the reference's numpy indexing would"""
    a, b = rasters(xb)
    ds = xb.OverlapRegridder(source=a, target=b).to_dataset()
    bad = dict(ds)
    bad["__regrid_indices"] = ds["__regrid_indices"].copy()
    bad["__regrid_indices"][3] = 9  # m == 9: out of range
    with pytest.raises(IndexError):
        xb.OverlapRegridder.from_dataset(bad)
    bad = dict(ds)
    bad["__regrid_indptr"] = ds["__regrid_indptr"][:-1]
    with pytest.raises(ValueError):
        xb.OverlapRegridder.from_dataset(bad)
    r = xb.OverlapRegridder(source=a, target=b)
    with pytest.raises(Exception, match="float64"):
        r.regrid_device(torch.zeros((1, 9), dtype=torch.float32, device="cuda"))
    with pytest.raises(ValueError, match="columns"):
        r.regrid_device(torch.zeros((1, 10), dtype=torch.float64, device="cuda"))
