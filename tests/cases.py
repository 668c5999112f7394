"""Known answers taken from the reference's own tests, shared by the oracle
(CPU) tests and the CUDA (GPU) tests."""
import numpy as np

nan = np.nan

# /root/reference/tests/test_regrid/test_reduce.py:8-110,173-199
# (method, values, weights, expected, percentile)
_FWD_V = [0.0, 1.0, 2.0, nan]
_FWD_W = [0.5, 0.5, 0.5, 0.5]
REDUCE_CASES = []
for _v in (_FWD_V, _FWD_V[::-1]):
    REDUCE_CASES += [
        ("mean", _v, _FWD_W, 1.0, None),
        ("harmonic_mean", _v, _FWD_W, 1.0 / (0.5 / 1.0 + 0.5 / 2.0), None),
        ("geometric_mean", _v, _FWD_W, np.sqrt(1.0 * 2.0), None),
        ("sum", _v, _FWD_W, 3.0, None),
        ("minimum", _v, _FWD_W, 0.0, None),
        ("maximum", _v, _FWD_W, 2.0, None),
        ("mode", _v, _FWD_W, 2.0, None),
        ("median", _v, _FWD_W, 1.0, None),
        ("conductance", _v, _FWD_W, 1.5, None),
        ("first_order_conservative", _v, _FWD_W, 1.5, None),
        ("max_overlap", _v, _FWD_W, 2.0, None),
    ]
REDUCE_CASES += [
    ("max_overlap", [0.0, 1.0, 2.0, nan], [0.5, 1.5, 0.5, 2.5], 1.0, None),
    ("mode", [0.0, 1.0, 1.0, 2.0, nan], [0.5] * 5, 1.0, None),
    ("mode", [1.0, 1.0, 3.0], [1.0, 1.0, 1.0], 1.0, None),
    ("mode", [4.0, 5.0, 6.0], [0.5, 0.5, 0.5], 6.0, None),
    ("percentile", list(np.arange(8) * 0.5), [1.0] * 8, 0.0, 0.0),
    ("percentile", list(np.arange(8) * 0.5), [1.0] * 8, 3.5, 100.0),
    ("percentile", list(np.arange(8) * 0.5), [1.0] * 8, 1.75, 50.0),
    ("percentile", list(np.arange(12.0)), [1.0] * 12, 2.75, 25.0),
    ("percentile", list(np.arange(12.0)), [1.0] * 12, 5.5, 50.0),
    ("percentile", list(np.arange(12.0)), [1.0] * 12, 11.0, 100.0),
    ("percentile", [1.0, 2.0, 3.0], [1.0] * 3, 1.0, 0.0),
]
for _m in ("mean", "harmonic_mean", "geometric_mean", "sum", "minimum", "maximum", "mode",
           "first_order_conservative", "conductance", "max_overlap", "median"):
    REDUCE_CASES.append((_m, [1.0] * 5, [0.0] * 5, nan, None))   # test_weights_all_zeros
    REDUCE_CASES.append((_m, [nan] * 5, [1.0] * 5, nan, None))   # test_values_all_nan


def reduce_cases_as_csr():
    """One CSR row per case; returns per-method dict name -> (rows, data, indices, indptr, source, expected, p)."""
    groups = {}
    for method, v, w, exp, p in REDUCE_CASES:
        key = (method, p)
        groups.setdefault(key, []).append((v, w, exp))
    out = []
    for (method, p), rows in groups.items():
        src, data, indices, indptr, expected = [], [], [], [0], []
        for v, w, exp in rows:
            for vi, wi in zip(v, w):
                indices.append(len(src))
                src.append(vi)
                data.append(wi)
            indptr.append(len(indices))
            expected.append(exp)
        out.append((method, p, np.array(src)[None, :], np.array(data), np.array(indices, dtype=np.int64),
                    np.array(indptr, dtype=np.int64), np.array(expected)))
    return out


# /root/reference/tests/test_ugrid2d.py:31-48,77-91
VERTICES = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [0.0, 1.0], [1.0, 1.0], [2.0, 1.0], [1.0, 2.0]])
FACES = np.array([[0, 1, 4, 3], [1, 2, 5, 4], [3, 4, 6, -1], [4, 5, 6, -1]])
CENTROIDS = np.array([[0.5, 0.5], [1.5, 0.5], [2.0 / 3.0, 4.0 / 3.0], [4.0 / 3.0, 4.0 / 3.0]])

# tests/test_ugrid2d.py:751-791
BARY_XY = np.array([[0.0, 0.0], [0.5, 0.5], [1.5, 0.5], [0.5, 1.5], [2.0, 2.0]])
BARY_FACE = np.array([0, 0, 1, 2, -1])
BARY_W = np.array([
    [1.0, 0.0, 0.0, 0.0],
    [0.25, 0.25, 0.25, 0.25],
    [0.25, 0.25, 0.25, 0.25],
    [0.5, 0.0, 0.5, 0.0],
    [0.0, 0.0, 0.0, 0.0],
])
BARY_FACE_TOL = np.array([-1, 0, 1, 2, -1])
BARY_W_TOL = np.array([
    [0.0, 0.0, 0.0, 0.0],
    [0.25, 0.25, 0.25, 0.25],
    [0.25, 0.25, 0.25, 0.25],
    [0.5, 0.0, 0.5, 0.0],
    [0.0, 0.0, 0.0, 0.0],
])

# tests/test_ugrid2d.py:794-825
RASTER_INDEX = np.array([[-1, 2, 3, -1], [2, 2, 3, 3], [0, 0, 1, 1], [0, 0, 1, 1]])
RASTER_INDEX_BOUNDS = np.array([
    [-1, -1, -1, 2, 3, -1],
    [-1, -1, 2, 2, 3, 3],
    [-1, -1, 0, 0, 1, 1],
    [-1, -1, 0, 0, 1, 1],
    [-1, -1, -1, -1, -1, -1],
    [-1, -1, -1, -1, -1, -1],
])


def canon(ti, si, a):
    o = np.lexsort((si, ti))
    return ti[o], si[o], a[o]


def hexagon_mesh(nx=6, ny=5, seed=0):
    """regular hexagons (k = 6) plus a few triangles padded with -1: exercises k > 4"""
    rng = np.random.default_rng(seed)
    verts, faces = [], []
    for j in range(ny):
        for i in range(nx):
            cx = i * 1.5
            cy = j * np.sqrt(3.0) + (i % 2) * np.sqrt(3.0) / 2
            base = len(verts)
            for q in range(6):
                ang = np.pi / 3 * q
                verts.append([cx + np.cos(ang) * 0.98, cy + np.sin(ang) * 0.98])
            faces.append([base + q for q in range(6)])
    verts = np.array(verts) + rng.uniform(-0.005, 0.005, (len(verts), 2))
    faces = np.array(faces, dtype=np.int64)
    # append triangles (fill -1)
    tri = np.array([[0, 1, 2, -1, -1, -1], [6, 7, 9, -1, -1, -1]], dtype=np.int64)
    return verts, np.vstack([faces, tri])


# ---- structured fixtures a..e of /root/reference/tests/fixtures/fixture_regridder.py:94-182 -----
# (midpoints; dx = 50 / dy = -50 unless stated) and the expectations the reference holds for them
FIX_A = dict(x=[50.0, 100.0, 150.0], y=[150.0, 100.0, 50.0])
FIX_B = dict(x=[25.0, 75.0, 125.0, 175.0], y=[175.0, 125.0, 75.0, 25.0])
FIX_C = dict(x=[40.0, 90.0, 140.0, 190.0], y=[175.0, 125.0, 75.0, 25.0])
FIX_E_X = [30.0, 67.5, 105.0]
FIX_E_XBOUNDS = np.column_stack(([17.5, 42.5, 92.5], [42.5, 92.5, 117.5]))
FIX_A_DATA = np.arange(9.0).reshape(3, 3)                       # fixture_regridder.py:94-105
FIX_A_LAYERED = np.arange(18.0).reshape(2, 3, 3)                # :108-121

# tests/test_regrid/test_structured.py:108-205: 36 (source, target) pairs, every weight 625.0
OVERLAP_2D_SOURCE = np.repeat(np.arange(9), 4)
OVERLAP_2D_TARGET = np.array([0, 4, 5, 1, 2, 6, 5, 1, 2, 3, 7, 6, 8, 9, 5, 4, 9, 5, 10, 6, 10, 11, 7, 6,
                              9, 8, 12, 13, 10, 14, 13, 9, 10, 11, 14, 15])
# test_structured.py:249-262
CENTROID_2D_SOURCE = np.array([0, 1, 3, 4])
CENTROID_2D_TARGET = np.array([5, 6, 9, 10])
# test_structured.py:51-105: the 1-D cases (source, target, length)
OVERLAP_1D = {
    "a_b": ([0, 0, 1, 1, 2, 2], [0, 1, 1, 2, 2, 3], [25.0] * 6),
    "a_b_flipped": ([0, 0, 1, 1, 2, 2], [2, 3, 1, 2, 0, 1], [25.0] * 6),
    "a_e": ([0, 0, 1, 1], [0, 1, 1, 2], [17.5, 32.5, 17.5, 25.0]),
    "a_e_relative": ([0, 0, 1, 1], [0, 1, 1, 2], [17.5 / 50.0, 32.5 / 50.0, 17.5 / 50.0, 25.0 / 50.0]),
}
# fixture_regridder.py:240-362 (consumed by test_regridder.py:100-150 with a weak `.any()`;
# asserted as full equality here)
EXPECTED_CENTROID = np.array([nan, nan, nan, nan, nan, 0, 1, nan, nan, 3, 4, nan, nan, nan, nan, nan]).reshape(4, 4)
EXPECTED_OVERLAP = np.array([0.0, 0.5, 1.5, 2.0, 1.5, 2.0, 3.0, 3.5, 4.5, 5.0, 6.0, 6.5, 6.0, 6.5, 7.5, 8.0]).reshape(4, 4)
EXPECTED_LINEAR = np.array([nan, nan, nan, nan, nan, 2.0, 3.0, nan, nan, 5.0, 6.0, nan, nan, nan, nan, nan]).reshape(4, 4)


def strip_axis_y():
    """one row of unit height: a 1-D grid as a strip of quads (areas == lengths)"""
    return dict(y=[0.5], ybounds=np.array([[0.0, 1.0]]))


def pair_set(source, target, weight):
    """the comparison of test_structured.py:29-48: the set of (target, source) pairs with weights"""
    return sorted(zip(np.asarray(target).tolist(), np.asarray(source).tolist(), np.asarray(weight, float).tolist()))
