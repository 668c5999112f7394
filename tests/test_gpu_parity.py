"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle,
the committed golden vectors and the reference's known answers.

Bars (BASELINE.json north_star): indices bit-exact, float64 areas / values
within 1e-10 relative.  Because libxgregrid is compiled with -fmad=false the
areas are in fact bit-identical to the oracle's, and that is what is asserted."""
import numpy as np
import pytest
import torch

import oracle
from tests import cases
from xugrid_b200 import synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-10


@pytest.fixture(scope="module")
def xb():
    import xugrid_b200

    assert torch.cuda.is_available()
    return xugrid_b200


def gpu_triplets(xb, sxy, sf, txy, tf, **kw):
    tree = xb.CellTree2d(sxy, sf, **kw)
    return cases.canon(*tree.intersect_faces(txy, tf, kw.get("fill_value", -1))), tree


def oracle_triplets(sxy, sf, txy, tf):
    return cases.canon(*oracle.CellTree2d(sxy, sf).intersect_faces(txy, tf))


def assert_same_triplets(a, b):
    assert a[0].shape == b[0].shape
    assert np.array_equal(a[0], b[0])  # target index, bit-exact
    assert np.array_equal(a[1], b[1])  # source index, bit-exact
    assert np.array_equal(a[2], b[2])  # area: bit-identical (stronger than rtol 1e-10)


MESH_PAIRS = {
    "delaunay_tri_tri": lambda: (*synthetic.delaunay(3000, 1), *synthetic.delaunay(2500, 2)),
    "mixed_to_tri": lambda: (*synthetic.jittered_lattice(40, 30, 3, "mixed"), *synthetic.jittered_lattice(37, 33, 4, "tri")),
    "tri_to_mixed": lambda: (*synthetic.jittered_lattice(37, 33, 4, "tri"), *synthetic.jittered_lattice(40, 30, 3, "mixed")),
    "quad_to_quad": lambda: (*synthetic.jittered_lattice(21, 17, 5, "quad"), *synthetic.jittered_lattice(16, 23, 6, "quad")),
    "self_overlap_slivers": lambda: (*synthetic.jittered_lattice(25, 25, 7, "mixed"), *synthetic.jittered_lattice(25, 25, 7, "mixed")),
    "raster_on_raster": lambda: (*synthetic.jittered_lattice(20, 20, 8, "quad", jitter=0.0), *synthetic.jittered_lattice(10, 10, 9, "quad", jitter=0.0)),
    "hexagons_k6": lambda: (*cases.hexagon_mesh(6, 5, 0), *synthetic.jittered_lattice(15, 14, 10, "tri", bounds=(-1.0, 9.0, -1.0, 9.0))),
    "tri_onto_hexagons": lambda: (*synthetic.jittered_lattice(15, 14, 10, "tri", bounds=(-1.0, 9.0, -1.0, 9.0)), *cases.hexagon_mesh(6, 5, 0)),
    # 61k -> 61k faces: enough pairs that every thread of the persistent SAT / clip grids takes
    # several grid-stride trips (the software-pipelined loop bodies; smaller cases leave after one)
    "tri_tri_multi_trip": lambda: (*synthetic.jittered_lattice(175, 175, 12, "tri"), *synthetic.jittered_lattice(173, 178, 13, "tri")),
    "disk_to_coarse": lambda: (*synthetic.disk(), *synthetic.jittered_lattice(10, 10, 11, "quad", jitter=0.0, bounds=(-10.0, 10.0, -10.0, 10.0))),
}


@pytest.mark.parametrize("name", list(MESH_PAIRS))
def test_overlap_bit_exact(xb, name):
    sxy, sf, txy, tf = MESH_PAIRS[name]()
    g, tree = gpu_triplets(xb, sxy, sf, txy, tf)
    o = oracle_triplets(sxy, sf, txy, tf)
    assert_same_triplets(g, o)
    assert tree.last_stats.nnz == len(o[0])


def polygon_mesh_k(k, n=40, seed=0):
    """n disjoint regular k-gons with a little jitter (exercises the generic k path)"""
    rng = np.random.default_rng(seed)
    verts, faces = [], []
    for i in range(n):
        cx, cy = (i % 8) * 2.2, (i // 8) * 2.2
        base = len(verts)
        ang = np.linspace(0, 2 * np.pi, k, endpoint=False) + rng.uniform(0, 1)
        for a in ang:
            verts.append([cx + np.cos(a), cy + np.sin(a)])
        faces.append(list(range(base, base + k)))
    return np.array(verts), np.array(faces, dtype=np.int64)


@pytest.mark.parametrize("k", [5, 8, 12, 20])
def test_overlap_many_vertex_polygons(xb, k):
    sxy, sf = polygon_mesh_k(k, seed=k)
    txy, tf = synthetic.jittered_lattice(30, 20, k, "mixed", bounds=(-1.5, 17.5, -1.5, 10.5))
    assert_same_triplets(gpu_triplets(xb, sxy, sf, txy, tf)[0], oracle_triplets(sxy, sf, txy, tf))
    # and with the polygons as queries
    assert_same_triplets(gpu_triplets(xb, txy, tf, sxy, sf)[0], oracle_triplets(txy, tf, sxy, sf))


def test_overlap_orientation_independent(xb):
    # tests/test_regrid/test_regridder.py:295-332
    sxy, sf = synthetic.jittered_lattice(19, 18, 1, "mixed")
    txy, tf = synthetic.jittered_lattice(17, 21, 2, "tri")
    cw_s, cw_t = sf.copy(), tf[:, ::-1].copy()
    for r in range(len(cw_s)):
        n = 4 if cw_s[r, 3] != -1 else 3
        cw_s[r, :n] = cw_s[r, :n][::-1]
    a = gpu_triplets(xb, sxy, sf, txy, tf)[0]
    b = gpu_triplets(xb, sxy, cw_s, txy, cw_t)[0]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.allclose(a[2], b[2], rtol=1e-12, atol=0)
    assert_same_triplets(b, oracle_triplets(sxy, cw_s, txy, cw_t))


def test_overlap_int32_and_fill_value(xb):
    sxy, sf = synthetic.jittered_lattice(12, 11, 1, "mixed")
    txy, tf = synthetic.jittered_lattice(9, 13, 2, "mixed")
    o = oracle_triplets(sxy, sf, txy, tf)
    g32 = cases.canon(*xb.CellTree2d(sxy, sf.astype(np.int32)).intersect_faces(txy, tf.astype(np.int32)))
    assert_same_triplets(g32, o)
    sf9, tf9 = sf.copy(), tf.copy()
    sf9[sf9 < 0] = -999
    tf9[tf9 < 0] = 999999
    g = cases.canon(*xb.CellTree2d(sxy, sf9, -999).intersect_faces(txy, tf9, 999999))
    assert_same_triplets(g, o)


def test_overlap_empty_and_disjoint(xb):
    sxy, sf = synthetic.delaunay(200, 1)
    tree = xb.CellTree2d(sxy, sf)
    ti, si, w = tree.intersect_faces(sxy + 10.0, sf)
    assert ti.size == 0 and si.size == 0 and w.size == 0
    indptr, indices, area = tree.intersect_faces_csr(sxy, sf, t_begin=5, t_end=5)
    assert indptr.tolist() == [0] and indices.numel() == 0
    # touching only: unit squares sharing an edge have no positive overlap
    a_xy = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    b_xy = a_xy + [1.0, 0.0]
    q = np.array([[0, 1, 2, 3]])
    ti, si, w = xb.CellTree2d(a_xy, q).intersect_faces(b_xy, q)
    assert w.size == 0


def test_overlap_target_range_equals_full(xb):
    sxy, sf = synthetic.delaunay(2000, 5)
    txy, tf = synthetic.delaunay(1800, 6)
    tree = xb.CellTree2d(sxy, sf)
    indptr, indices, area = (t.cpu().numpy() for t in tree.intersect_faces_csr(txy, tf))
    n = len(tf)
    for b, e in ((0, n // 3), (n // 3, n - 7), (n - 7, n)):
        ip, ix, ar = (t.cpu().numpy() for t in tree.intersect_faces_csr(txy, tf, t_begin=b, t_end=e))
        assert np.array_equal(ip, indptr[b : e + 1] - indptr[b])
        assert np.array_equal(ix, indices[indptr[b] : indptr[e]])
        assert np.array_equal(ar, area[indptr[b] : indptr[e]])
    # rows are sorted by source index
    rows = np.repeat(np.arange(n), np.diff(indptr))
    order = np.lexsort((indices, rows))
    assert np.array_equal(order, np.arange(len(order)))


def test_overlap_relative(xb):
    sxy, sf = synthetic.jittered_lattice(14, 12, 1, "mixed")
    txy, tf = synthetic.jittered_lattice(11, 15, 2, "tri")
    tree = xb.CellTree2d(sxy, sf)
    indptr, indices, w = (t.cpu().numpy() for t in tree.intersect_faces_csr(txy, tf, relative=True))
    si, ti, ow = oracle.overlap_triplets(sxy, sf, txy, tf, relative=True)
    o = cases.canon(ti, si, ow)
    rows = np.repeat(np.arange(len(tf)), np.diff(indptr))
    assert np.array_equal(rows, o[0]) and np.array_equal(indices, o[1])
    assert np.array_equal(w, o[2])


# ---------------------------------------------------------------------------- points
def test_locate_known_answers(xb):
    t = xb.CellTree2d(cases.VERTICES, cases.FACES)
    assert np.array_equal(t.locate_points(cases.CENTROIDS), [0, 1, 2, 3])
    assert np.array_equal(t.locate_points(np.array([[-0.01, 1.0], [-0.01, 0.5]]), 0.011), [0, 0])
    g = xb.Ugrid2d(cases.VERTICES[:, 0], cases.VERTICES[:, 1], -1, cases.FACES)
    x, y, index = g.rasterize(resolution=0.5)
    assert np.allclose(x, [0.25, 0.75, 1.25, 1.75]) and np.allclose(y, [1.75, 1.25, 0.75, 0.25])
    assert np.array_equal(index, cases.RASTER_INDEX)
    x, y, index = g.rasterize(resolution=0.5, bounds=(-1.0, -1.0, 2.0, 2.0))
    assert np.array_equal(index, cases.RASTER_INDEX_BOUNDS)


@pytest.mark.parametrize("mesh", ["delaunay", "mixed", "hexagons"])
def test_locate_parity(xb, mesh):
    if mesh == "delaunay":
        sxy, sf = synthetic.delaunay(5000, 5)
        lo, hi = -0.05, 1.05
    elif mesh == "mixed":
        sxy, sf = synthetic.jittered_lattice(40, 35, 3, "mixed")
        lo, hi = -0.05, 1.05
    else:
        sxy, sf = cases.hexagon_mesh(6, 5, 0)
        lo, hi = -1.5, 9.5
    rng = np.random.default_rng(6)
    pts = rng.random((50000, 2)) * (hi - lo) + lo
    # exact vertices and edge midpoints: ambiguous between neighbouring faces
    tri0 = sf[: min(200, len(sf))]
    amb = np.vstack([sxy[tri0[:, 0]], 0.5 * (sxy[tri0[:, 0]] + sxy[tri0[:, 1]])])
    pts = np.vstack([pts, amb, [[np.nan, 0.5]]])
    g = xb.CellTree2d(sxy, sf).locate_points(pts)
    ot = oracle.CellTree2d(sxy, sf)
    o = ot.locate_points(pts)
    assert np.array_equal(g < 0, o < 0)
    diff = np.nonzero(g != o)[0]
    # a point on a shared edge/vertex may resolve to either face (regridder.py:369-370);
    # the GPU picks the smallest qualifying face index
    assert len(diff) <= 2 * len(amb)
    for i in diff:
        allf = ot.locate_point_all(pts[i, 0], pts[i, 1])
        assert o[i] in allf and g[i] == allf.min()
    interior = np.arange(50000)
    assert np.array_equal(g[interior], o[interior])


def test_barycentric_known_answers(xb):
    t = xb.CellTree2d(cases.VERTICES, cases.FACES)
    face, w = t.compute_barycentric_weights(cases.BARY_XY)
    assert np.array_equal(face, cases.BARY_FACE) and np.allclose(w, cases.BARY_W)
    xy = cases.BARY_XY.copy()
    xy[:, 0] -= 0.01
    face, w = t.compute_barycentric_weights(xy, tolerance=0.01)
    assert np.array_equal(face, cases.BARY_FACE_TOL) and np.allclose(w, cases.BARY_W_TOL, atol=0.05)


@pytest.mark.parametrize("mesh", ["delaunay", "mixed", "hexagons"])
def test_barycentric_parity(xb, mesh):
    if mesh == "delaunay":
        sxy, sf = synthetic.delaunay(3000, 7)
        lo, hi = -0.02, 1.02
    elif mesh == "mixed":
        sxy, sf = synthetic.jittered_lattice(30, 25, 8, "mixed")
        lo, hi = -0.02, 1.02
    else:
        sxy, sf = cases.hexagon_mesh(6, 5, 0)
        lo, hi = -1.5, 9.5
    pts = np.random.default_rng(9).random((20000, 2)) * (hi - lo) + lo
    gi, gw = xb.CellTree2d(sxy, sf).compute_barycentric_weights(pts)
    oi, ow = oracle.CellTree2d(sxy, sf).compute_barycentric_weights(pts)
    assert np.array_equal(gi, oi)
    assert np.allclose(gw, ow, rtol=RTOL, atol=1e-15)
    inside = gi >= 0
    assert np.allclose(gw[inside].sum(axis=1), 1.0, rtol=1e-12)
    # weights reproduce the point: sum_j w_j * vertex_j == p
    conn = sf[gi[inside]]
    vx = np.where(conn >= 0, sxy[conn.clip(0), 0], 0.0)
    vy = np.where(conn >= 0, sxy[conn.clip(0), 1], 0.0)
    assert np.allclose((gw[inside] * vx).sum(1), pts[inside, 0], atol=1e-9)
    assert np.allclose((gw[inside] * vy).sum(1), pts[inside, 1], atol=1e-9)


def test_barycentric_clockwise_faces_keep_vertex_order(xb):
    sxy, sf = synthetic.jittered_lattice(10, 9, 3, "quad")
    pts = np.random.default_rng(1).random((2000, 2))
    gi, gw = xb.CellTree2d(sxy, sf).compute_barycentric_weights(pts)
    gi2, gw2 = xb.CellTree2d(sxy, sf[:, ::-1].copy()).compute_barycentric_weights(pts)
    assert np.array_equal(gi, gi2)
    assert np.allclose(gw, gw2[:, ::-1], rtol=1e-9, atol=1e-14)


# ---------------------------------------------------------------------------- geometry
def test_area_centroid_golden(xb, golden_dir):
    k = np.load(f"{golden_dir}/connectivity_golden.npz")
    g = xb.Ugrid2d(k["xy"][:, 0], k["xy"][:, 1], -1, k["faces"])
    assert np.array_equal(g.area, k["area"])
    assert np.array_equal(g.centroids, k["centroids"])
    g3 = xb.Ugrid2d(k["xy"][:, 0], k["xy"][:, 1], -1, k["tri"].astype(np.int32))
    assert np.array_equal(g3.area, k["tri_area"])
    assert np.array_equal(g3.centroids, k["tri_centroids"])


# ---------------------------------------------------------------------------- apply
def test_apply_golden_all_methods(xb, golden_dir):
    from xugrid_b200 import sparse

    g = np.load(f"{golden_dir}/apply_golden.npz")
    A = sparse.MatrixCSR.from_arrays(g["data"], g["indices"], g["indptr"], int(g["n"]), int(g["m"]))
    src = torch.from_numpy(g["source"]).cuda()
    transcendental = {"geometric_mean"}
    n_checked = 0
    for name, (code, p) in sparse.METHOD_CODES.items():
        out = sparse.csr_apply(code, p, src, A, A.n).cpu().numpy()
        ref = g["out_" + name]
        assert np.array_equal(np.isnan(out), np.isnan(ref)), name
        if name in transcendental:
            assert np.allclose(out, ref, rtol=RTOL, atol=0, equal_nan=True), name
        else:
            assert np.array_equal(out, ref, equal_nan=True), name
        n_checked += 1
    for p in (0.0, 33.3, 100.0):
        out = sparse.csr_apply(7, p, src, A, A.n).cpu().numpy()
        assert np.array_equal(out, g[f"out_percentile_{p}"], equal_nan=True)
    assert n_checked == 18
    isrc = torch.from_numpy(g["int_source"]).cuda().to(torch.float64)
    out = sparse.csr_apply(0, 0.0, isrc, A, A.n).cpu().numpy()
    assert np.array_equal(out, g["out_int_mean"], equal_nan=True)
    # row shards write into column blocks of a wider output (out_ld > rows)
    full = torch.full((src.shape[0], A.n), -7.0, dtype=torch.float64, device="cuda")
    for b, e in ((0, 150), (150, A.n)):
        sub = sparse.MatrixCSR(A.data, A.indices, A.indptr[b : e + 1].contiguous(), e - b, A.m, 0)
        sparse.csr_apply(0, 0.0, src, sub, e - b, out=full[:, b:e])
    assert np.array_equal(full.cpu().numpy(), g["out_mean"], equal_nan=True)


def test_apply_many_steps_equals_step_by_step(xb, golden_dir):
    """k_csr_apply_steps (8 time steps per thread) against the per-step kernel, bitwise:
    19 steps = two full groups + a remainder, NaNs, zero weights, empty rows"""
    from xugrid_b200 import sparse

    g = np.load(f"{golden_dir}/apply_golden.npz")
    A = sparse.MatrixCSR.from_arrays(g["data"], g["indices"], g["indptr"], int(g["n"]), int(g["m"]))
    rng = np.random.default_rng(7)
    src = rng.random((19, int(g["m"]))) - 0.3
    src[rng.random(src.shape) < 0.1] = np.nan
    src[rng.random(src.shape) < 0.05] = 0.0
    src = torch.from_numpy(src).cuda()
    for name in ("mean", "harmonic_mean", "sum", "minimum", "maximum", "max_overlap", "first_order_conservative",
                 "mode", "median", "geometric_mean"):
        code, p = sparse.METHOD_CODES[name]
        many = sparse.csr_apply(code, p, src, A, A.n).cpu().numpy()
        one = np.concatenate([sparse.csr_apply(code, p, src[e:e + 1], A, A.n).cpu().numpy() for e in range(19)])
        assert np.array_equal(many, one, equal_nan=True), name


@pytest.mark.parametrize("case", cases.reduce_cases_as_csr(), ids=lambda c: f"{c[0]}-{c[1]}")
def test_reduce_known_answers(xb, case):
    from xugrid_b200 import sparse

    method, p, src, data, indices, indptr, expected = case
    code, pp = (7, p) if method == "percentile" else sparse.METHOD_CODES[method]
    A = sparse.MatrixCSR.from_arrays(data, indices, indptr, len(expected), src.shape[1])
    out = sparse.csr_apply(code, pp, torch.from_numpy(src).cuda(), A, A.n).cpu().numpy()[0]
    assert np.allclose(out, expected, equal_nan=True)


def test_coo_gather_golden(xb, golden_dir):
    from xugrid_b200 import sparse

    c = np.load(f"{golden_dir}/coo_golden.npz")
    A = sparse.MatrixCOO.from_triplet(c["row"], c["col"], np.ones(len(c["row"])), n=int(c["size"]), m=c["source"].shape[1])
    out = sparse.coo_gather(torch.from_numpy(c["source"]).cuda(), A, int(c["size"])).cpu().numpy()
    assert np.array_equal(out, c["out"], equal_nan=True)
    assert np.array_equal(A.to_csr().indptr.cpu().numpy(), c["indptr"])


def test_csr_apply_rejects_bad_percentile(xb):
    from xugrid_b200 import _native, sparse

    A = sparse.MatrixCSR.from_arrays(np.ones(1), np.zeros(1, np.int64), np.array([0, 1]), 1, 1)
    with pytest.raises(_native.XgError):
        sparse.csr_apply(7, 101.0, torch.ones(1, 1, dtype=torch.float64, device="cuda"), A, 1)
    with pytest.raises(_native.XgError):
        sparse.csr_apply(99, 0.0, torch.ones(1, 1, dtype=torch.float64, device="cuda"), A, 1)


# ---------------------------------------------------------------------------- full size
def test_full_size_properties(xb):
    """10M -> 10M triangles (the size the BASELINE metric is quoted on):
    size-independent properties + oracle parity on a sampled target range against
    the full-size source."""
    from xugrid_b200 import sparse

    n = 2236
    sxy, sf = synthetic.jittered_lattice(n, n, 10, "tri")
    txy, tf = synthetic.jittered_lattice(n + 3, n - 2, 11, "tri", bounds=(0.05, 0.95, 0.05, 0.95))
    tree = xb.CellTree2d(sxy, sf)
    indptr, indices, area = tree.intersect_faces_csr(txy, tf)
    n_t, n_s = len(tf), len(sf)
    # (1) targets are interior: every row sums to the target's own area
    t_area = torch.from_numpy(oracle.face_area(txy, tf)).cuda()
    rows = torch.repeat_interleave(torch.arange(n_t, device="cuda"), indptr[1:] - indptr[:-1])
    row_sum = torch.zeros(n_t, dtype=torch.float64, device="cuda").index_add_(0, rows, area)
    assert torch.allclose(row_sum, t_area, rtol=1e-9, atol=0)
    # (2) sources fully covered by the target domain: column sums equal source areas
    s_area = torch.from_numpy(oracle.face_area(sxy, sf)).cuda()
    col_sum = torch.zeros(n_s, dtype=torch.float64, device="cuda").index_add_(0, indices.long(), area)
    cent = torch.from_numpy(oracle.face_centroids(sxy, sf)).cuda()
    inner = ((cent > 0.06) & (cent < 0.94)).all(dim=1)
    assert torch.allclose(col_sum[inner], s_area[inner], rtol=1e-9, atol=0)
    # (3) rows sorted by source index, no duplicates
    same_row = rows[1:] == rows[:-1]
    assert bool((indices[1:][same_row] > indices[:-1][same_row]).all())
    # (4) mean of a constant field is the constant; linearity of the conservative sum
    A = sparse.MatrixCSR(area, indices, indptr, n_t, n_s, int(area.numel()))
    ones = torch.full((1, n_s), 3.25, dtype=torch.float64, device="cuda")
    assert torch.allclose(sparse.csr_apply(0, 0.0, ones, A, n_t), ones[:, :1].expand(1, n_t), rtol=1e-14, atol=0)
    v1 = torch.rand(1, n_s, dtype=torch.float64, device="cuda")
    v2 = torch.rand(1, n_s, dtype=torch.float64, device="cuda")
    c = lambda v: sparse.csr_apply(9, 0.0, v, A, n_t)
    assert torch.allclose(c(v1 + v2), c(v1) + c(v2), rtol=1e-12, atol=0)
    # (5) oracle parity on a contiguous target sample against the full source tree
    b, e = n_t // 2, n_t // 2 + 20000
    oti, osi, oa = cases.canon(*oracle.CellTree2d(sxy, sf).intersect_faces(txy, tf[b:e]))
    ip = indptr[b : e + 1].cpu().numpy()
    assert np.array_equal(np.repeat(np.arange(e - b), np.diff(ip)), oti)
    assert np.array_equal(indices[ip[0] : ip[-1]].cpu().numpy(), osi)
    assert np.array_equal(area[ip[0] : ip[-1]].cpu().numpy(), oa)


def test_overlap_dense_sources_rerun_path(xb):
    """~130 candidates per target: more than the arena guess (16 per target), so
    the walk is re-run with the exact size it counted; also rows longer than the
    8-entry register sort."""
    sxy, sf = synthetic.jittered_lattice(160, 160, 21, "tri")
    txy, tf = synthetic.jittered_lattice(20, 20, 22, "mixed", bounds=(0.03, 0.97, 0.03, 0.97))
    g, tree = gpu_triplets(xb, sxy, sf, txy, tf)
    assert tree.last_stats.candidates > 16 * len(tf)
    assert tree.last_stats.walks == 2
    assert_same_triplets(g, oracle_triplets(sxy, sf, txy, tf))
    indptr, indices, _ = (t.cpu().numpy() for t in tree.intersect_faces_csr(txy, tf))
    assert np.diff(indptr).max() > 8
    rows = np.repeat(np.arange(len(tf)), np.diff(indptr))
    assert np.array_equal(np.lexsort((indices, rows)), np.arange(len(indices)))


@pytest.mark.parametrize("n_coarse,min_row", [(6, 17), (1, 1025)])
def test_overlap_long_rows_are_sorted(xb, n_coarse, min_row):
    """rows longer than the register networks (warp rank sort in shared memory) and
    longer than its capacity (one target covering thousands of source faces)"""
    sxy, sf = synthetic.jittered_lattice(48, 48, 41, "tri")
    txy, tf = synthetic.jittered_lattice(n_coarse, n_coarse, 42, "quad", jitter=0.0, bounds=(0.01, 0.99, 0.01, 0.99))
    g, tree = gpu_triplets(xb, sxy, sf, txy, tf)
    assert_same_triplets(g, oracle_triplets(sxy, sf, txy, tf))
    indptr, indices, _ = (t.cpu().numpy() for t in tree.intersect_faces_csr(txy, tf))
    assert np.diff(indptr).max() >= min_row
    rows = np.repeat(np.arange(len(tf)), np.diff(indptr))
    assert np.array_equal(np.lexsort((indices, rows)), np.arange(len(indices)))


def test_overlap_fine_source_arena_sized_by_sampling(xb):
    """a source much finer than the targets: the candidate density is sampled first, so
    one walk is enough (without it the default arena overflows and the walk runs twice)"""
    sxy, sf = synthetic.jittered_lattice(420, 420, 51, "tri")
    txy, tf = synthetic.jittered_lattice(95, 95, 52, "tri", bounds=(0.02, 0.98, 0.02, 0.98))
    g, tree = gpu_triplets(xb, sxy, sf, txy, tf)
    assert tree.last_stats.candidates > 16 * len(tf)
    assert tree.last_stats.walks == 1
    assert_same_triplets(g, oracle_triplets(sxy, sf, txy, tf))


def test_clipped_index_equals_full_index_on_a_target_range(xb):
    from xugrid_b200.distributed import shard_range, target_range_box

    sxy, sf = synthetic.delaunay(4000, 31)
    txy, tf = synthetic.jittered_lattice(40, 40, 32, "tri", bounds=(0.02, 0.98, 0.02, 0.98))
    full = xb.CellTree2d(sxy, sf)
    dxy, dtf = torch.from_numpy(txy).cuda(), torch.from_numpy(tf).cuda()
    for rank in range(4):
        b, e = shard_range(len(tf), rank, 4)
        box = target_range_box(dxy, dtf, b, e)
        part = xb.CellTree2d(sxy, sf, clip_box=box)
        assert 0 < part.n_indexed < full.n_indexed
        ref = [t.cpu().numpy() for t in full.intersect_faces_csr(dxy, dtf, t_begin=b, t_end=e)]
        got = [t.cpu().numpy() for t in part.intersect_faces_csr(dxy, dtf, t_begin=b, t_end=e)]
        for x, y in zip(ref, got):
            assert np.array_equal(x, y)
    # a box that misses the mesh indexes nothing and finds nothing
    empty = xb.CellTree2d(sxy, sf, clip_box=(5.0, 6.0, 5.0, 6.0))
    assert empty.n_indexed == 0
    ip, ix, ar = empty.intersect_faces_csr(dxy, dtf)
    assert ix.numel() == 0 and int(ip[-1]) == 0
