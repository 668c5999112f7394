"""CPU tests: the oracle against the reference's golden vectors and known
answers.  No GPU needed."""
import numpy as np
import pytest

import oracle
from tests import cases
from xugrid_b200 import synthetic


def test_apply_golden(golden_dir):
    g = np.load(f"{golden_dir}/apply_golden.npz")
    n = int(g["n"])
    checked = 0
    for key in g.files:
        if not key.startswith("out_"):
            continue
        name = key[4:]
        if name.startswith("percentile_"):
            out = oracle.csr_apply("percentile", g["source"], g["data"], g["indices"], g["indptr"], n,
                                   percentile=float(name.split("_")[1]))
        elif name == "int_mean":
            out = oracle.csr_apply("mean", g["int_source"], g["data"], g["indices"], g["indptr"], n)
        else:
            out = oracle.csr_apply(name, g["source"], g["data"], g["indices"], g["indptr"], n)
        # the C restatement is bit-identical to the reference's numba code
        assert np.array_equal(out, g[key], equal_nan=True), name
        checked += 1
    assert checked == 22
    # the parallel-rows schedule gives the same numbers
    out = oracle.csr_apply("mean", g["source"], g["data"], g["indices"], g["indptr"], n, parallel_rows=True)
    assert np.array_equal(out, g["out_mean"], equal_nan=True)


def test_coo_golden(golden_dir):
    c = np.load(f"{golden_dir}/coo_golden.npz")
    out = oracle.coo_gather(c["source"], c["row"], c["col"], int(c["size"]))
    assert np.array_equal(out, c["out"], equal_nan=True)
    assert np.array_equal(oracle.coo_to_csr(c["row"], int(c["size"])), c["indptr"])


def test_connectivity_golden(golden_dir):
    k = np.load(f"{golden_dir}/connectivity_golden.npz")
    assert np.array_equal(oracle.face_area(k["xy"], k["faces"]), k["area"])
    assert np.array_equal(oracle.face_centroids(k["xy"], k["faces"]), k["centroids"])
    assert np.array_equal(oracle.face_area(k["xy"], k["tri"]), k["tri_area"])
    assert np.array_equal(oracle.face_centroids(k["xy"], k["tri"]), k["tri_centroids"])
    assert np.array_equal(oracle.counter_clockwise(k["xy"], k["cw"]), k["ccw"])


@pytest.mark.parametrize("case", cases.reduce_cases_as_csr(), ids=lambda c: f"{c[0]}-{c[1]}")
def test_reduce_known_answers(case):
    method, p, src, data, indices, indptr, expected = case
    out = oracle.csr_apply(method, src, data, indices, indptr, len(expected), percentile=p)[0]
    assert np.allclose(out, expected, equal_nan=True)


def test_locate_points_known_answers():
    # tests/test_ugrid2d.py:724-730
    t = oracle.CellTree2d(cases.VERTICES, cases.FACES, -1)
    assert np.array_equal(t.locate_points(cases.CENTROIDS), [0, 1, 2, 3])
    assert np.array_equal(t.locate_points([[-0.01, 1.0], [-0.01, 0.5]], 0.011), [0, 0])


def test_barycentric_known_answers():
    # tests/test_ugrid2d.py:751-791
    t = oracle.CellTree2d(cases.VERTICES, cases.FACES, -1)
    face, w = t.compute_barycentric_weights(cases.BARY_XY)
    assert np.array_equal(face, cases.BARY_FACE)
    assert np.allclose(w, cases.BARY_W)
    xy = cases.BARY_XY.copy()
    xy[:, 0] -= 0.01
    face, w = t.compute_barycentric_weights(xy, tolerance=0.01)
    assert np.array_equal(face, cases.BARY_FACE_TOL)
    assert np.allclose(w, cases.BARY_W_TOL, atol=0.05)


def test_rasterize_known_answers():
    # tests/test_ugrid2d.py:794-825
    t = oracle.CellTree2d(cases.VERTICES, cases.FACES, -1)
    x = np.array([0.25, 0.75, 1.25, 1.75]); y = np.array([1.75, 1.25, 0.75, 0.25])
    yy, xx = np.meshgrid(y, x, indexing="ij")
    idx = t.locate_points(np.column_stack([xx.ravel(), yy.ravel()])).reshape(4, 4)
    assert np.array_equal(idx, cases.RASTER_INDEX)
    x = np.array([-0.75, -0.25, 0.25, 0.75, 1.25, 1.75]); y = np.array([1.75, 1.25, 0.75, 0.25, -0.25, -0.75])
    yy, xx = np.meshgrid(y, x, indexing="ij")
    idx = t.locate_points(np.column_stack([xx.ravel(), yy.ravel()])).reshape(6, 6)
    assert np.array_equal(idx, cases.RASTER_INDEX_BOUNDS)


def test_self_overlap_is_identity():
    # tests/test_regrid/test_unstructured.py:32-45
    xy, faces = synthetic.disk()
    tree = oracle.CellTree2d(xy, faces)
    ti, si, w = tree.intersect_faces(xy, faces)
    valid = w > 1e-5
    assert np.array_equal(ti[valid], np.arange(len(faces)))
    assert np.array_equal(si[valid], np.arange(len(faces)))
    assert np.allclose(w[valid], oracle.face_area(xy, faces))
    # self-locate / self-barycentric identity (test_unstructured.py:47-59)
    cent = oracle.face_centroids(xy, faces)
    assert np.array_equal(tree.locate_points(cent), np.arange(len(faces)))


def test_tree_equals_bruteforce_and_invariants():
    sxy, sf = synthetic.delaunay(1500, 3)
    txy, tf = synthetic.jittered_lattice(23, 19, 4, "mixed", bounds=(0.1, 0.9, 0.1, 0.9))
    tree = oracle.CellTree2d(sxy, sf)
    a = cases.canon(*tree.intersect_faces(txy, tf))
    b = cases.canon(*tree.intersect_faces_bruteforce(txy, tf))
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # targets lie strictly inside the source hull: row sums equal target areas
    ta = oracle.face_area(txy, tf)
    rs = np.bincount(a[0], weights=a[2], minlength=len(ta))
    assert np.allclose(rs, ta, rtol=1e-9)
    # symmetry: overlap(A, B)^T == overlap(B, A)
    tree2 = oracle.CellTree2d(txy, tf)
    ti2, si2, w2 = tree2.intersect_faces(sxy, sf)
    c = cases.canon(si2, ti2, w2)
    assert np.array_equal(c[0], a[0]) and np.array_equal(c[1], a[1])
    assert np.allclose(c[2], a[2], rtol=1e-9, atol=1e-18)


def test_orientation_independence():
    # tests/test_regrid/test_regridder.py:295-332: flipping faces changes nothing
    sxy, sf = synthetic.jittered_lattice(9, 8, 1, "mixed")
    txy, tf = synthetic.jittered_lattice(7, 11, 2, "tri")
    cw = sf.copy()
    for r in range(len(cw)):
        n = 4 if cw[r, 3] != -1 else 3
        cw[r, :n] = cw[r, :n][::-1]
    a = cases.canon(*oracle.CellTree2d(sxy, sf).intersect_faces(txy, tf))
    b = cases.canon(*oracle.CellTree2d(sxy, cw).intersect_faces(txy, tf))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.allclose(a[2], b[2], rtol=1e-12)


def test_empty_inputs():
    xy, faces = synthetic.disk()
    tree = oracle.CellTree2d(xy, faces)
    ti, si, w = tree.intersect_faces(xy + 100.0, faces)
    assert ti.size == 0 and si.size == 0 and w.size == 0
    assert np.array_equal(tree.locate_points(np.array([[50.0, 50.0]])), [-1])


def test_intersect_edges_reference_known_answer():
    """/root/reference/tests/test_regrid/test_network_gridder.py:10-95: a 4x4 raster of unit
    cells (y descending) and a 4-edge network; nnz == 8 (point contacts do not count) and the
    length-weighted means the reference expects in five cells."""
    from xugrid_b200.grid import StructuredGrid2d

    g = StructuredGrid2d(np.arange(0.5, 4.5, 1.0), np.arange(3.5, -0.5, -1.0)).to_ugrid2d()
    node_xy = np.array([[0.0, 0.0], [1.5, 1.5], [2.5, 1.5], [4.0, 0.0], [4.0, 3.0]])
    edge_nodes = np.array([[0, 1], [1, 2], [2, 3], [2, 4]])
    values = np.array([1.0, 2.0, 4.0, -4.0])
    tree = oracle.CellTree2d(g.node_coordinates, g.face_node_connectivity)
    ei, fi, xy = tree.intersect_edges(node_xy[edge_nodes])
    assert len(ei) == 8
    assert np.array_equal(np.lexsort((fi, ei)), np.arange(8))
    length = np.linalg.norm(xy[:, 1] - xy[:, 0], axis=1)
    order = np.lexsort((ei, fi))
    indptr = oracle.coo_to_csr(fi[order], g.n_face)
    out = oracle.csr_apply("mean", values[None, :], length[order], ei[order], indptr, g.n_face)[0]
    assert np.count_nonzero(np.isnan(out)) == 11
    diag = 0.5 * np.sqrt(2)
    expected = np.array([1.0, (diag * 1 + 0.5 * 2) / (diag + 0.5),
                         (0.5 * 2 + diag * -4 + diag * 4) / (2 * diag + 0.5), -4.0, 4.0])
    cen = oracle.face_centroids(g.node_coordinates, g.face_node_connectivity)
    x_loc = np.array([0.5, 1.5, 2.5, 3.5, 3.5])
    y_loc = np.array([0.5, 1.5, 1.5, 2.5, 0.5])
    cells = [int(np.argmin((cen[:, 0] - x) ** 2 + (cen[:, 1] - y) ** 2)) for x, y in zip(x_loc, y_loc)]
    assert np.allclose(out[cells], expected)


def test_intersect_edges_segment_on_shared_edge_counts_once():
    """a segment lying on the edge shared by two faces belongs to the face on its left"""
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [0.0, 1.0], [1.0, 1.0], [2.0, 1.0]])
    faces = np.array([[0, 1, 4, 3], [1, 2, 5, 4]])
    tree = oracle.CellTree2d(xy, faces)
    up = np.array([[[1.0, 0.2], [1.0, 0.8]]])      # direction +y: the left face is face 0
    ei, fi, seg = tree.intersect_edges(up)
    assert fi.tolist() == [0] and np.allclose(seg[0], up[0])
    ei, fi, seg = tree.intersect_edges(up[:, ::-1])  # direction -y: the left face is face 1
    assert fi.tolist() == [1]


def test_intersect_edges_conserves_length():
    """segments inside the mesh hull: the clipped pieces of a segment add up to its length"""
    from xugrid_b200 import synthetic

    xy, f = synthetic.jittered_lattice(20, 17, 71, "mixed")
    rng = np.random.default_rng(72)
    a = rng.uniform(0.05, 0.95, (400, 2))
    b = np.clip(a + rng.normal(0.0, 0.15, (400, 2)), 0.05, 0.95)
    edges = np.stack([a, b], axis=1)
    ei, fi, seg = oracle.CellTree2d(xy, f).intersect_edges(edges)
    piece = np.linalg.norm(seg[:, 1] - seg[:, 0], axis=1)
    total = np.bincount(ei, weights=piece, minlength=len(edges))
    assert np.allclose(total, np.linalg.norm(b - a, axis=1), rtol=1e-12, atol=1e-14)
    # every piece lies on its segment and inside the face's bounding box
    bb = oracle.build_bboxes(xy, f)[fi]
    lo = np.minimum(seg[:, 0], seg[:, 1]); hi = np.maximum(seg[:, 0], seg[:, 1])
    assert (lo[:, 0] >= bb[:, 0] - 1e-12).all() and (hi[:, 0] <= bb[:, 1] + 1e-12).all()
    assert (lo[:, 1] >= bb[:, 2] - 1e-12).all() and (hi[:, 1] <= bb[:, 3] + 1e-12).all()


def test_clip_area_properties():
    """Sutherland-Hodgman area of random convex polygon pairs: symmetric, bounded by both
    areas, equal to the polygon's own area when clipped by itself or by a polygon around it"""
    rng = np.random.default_rng(9)

    def convex(n, cx, cy, r):
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        return np.column_stack([cx + r * np.cos(ang), cy + r * np.sin(ang)])

    def area(p):
        x, y = p[:, 0], p[:, 1]
        return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))

    for _ in range(300):
        a = convex(rng.integers(3, 8), rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.3, 1.5))
        b = convex(rng.integers(3, 8), rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.3, 1.5))
        ab, ba = oracle.clip_area(a, b), oracle.clip_area(b, a)
        assert abs(ab - ba) <= 1e-12 * max(area(a), area(b))
        assert ab <= min(area(a), area(b)) * (1 + 1e-12)
        assert abs(oracle.clip_area(a, a) - area(a)) <= 1e-12 * area(a)
        t = np.linspace(0.0, 2.0 * np.pi, 12, endpoint=False)  # a regular 12-gon around a
        big = np.column_stack([a[:, 0].mean() + 10.0 * np.cos(t), a[:, 1].mean() + 10.0 * np.sin(t)])
        assert abs(oracle.clip_area(a, big) - area(a)) <= 1e-12 * area(a)
        if ab > 0:
            assert oracle.polygons_intersect(a, b)
