"""
CPU ORACLE -- test infrastructure, not product code.

Python face of ``oracle/xg_oracle.c`` (see that file's header for what it
restates and which parts are pinned against the reference).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package; ``xugrid_b200`` never does.

The classes mirror the seam the reference talks to
(/root/reference/xugrid/ugrid/ugrid2d.py:908-921 and
xugrid/regrid/unstructured.py:124-139): ``CellTree2d(vertices, faces,
fill_value)`` with ``intersect_faces``, ``locate_points`` and
``compute_barycentric_weights``.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libxgoracle.so")
_lib = None

METHODS = {
    "mean": (0, 0.0),
    "harmonic_mean": (1, 0.0),
    "geometric_mean": (2, 0.0),
    "sum": (3, 0.0),
    "minimum": (4, 0.0),
    "maximum": (5, 0.0),
    "mode": (6, 0.0),
    "median": (7, 50.0),
    "max_overlap": (8, 0.0),
    "first_order_conservative": (9, 0.0),
    "conductance": (9, 0.0),
}
for _p in (5, 10, 25, 50, 75, 90, 95):
    METHODS[f"p{_p}"] = (7, float(_p))


def build(force: bool = False) -> str:
    """Compile the C restatement with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "xg_oracle.c")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    ):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def _p_i64(a: np.ndarray):
    return a.ctypes.data_as(_i64p)


def _p_f64(a: np.ndarray):
    return a.ctypes.data_as(_f64p)


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.xo_celltree_create.restype = ctypes.c_void_p
        L.xo_celltree_create.argtypes = [
            _f64p, ctypes.c_int64, _i64p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ]
        L.xo_celltree_destroy.argtypes = [ctypes.c_void_p]
        L.xo_celltree_n_nodes.restype = ctypes.c_int64
        L.xo_celltree_n_nodes.argtypes = [ctypes.c_void_p]
        L.xo_celltree_n_candidates.restype = ctypes.c_int64
        L.xo_celltree_n_candidates.argtypes = [ctypes.c_void_p]
        L.xo_default_tolerance.restype = ctypes.c_double
        L.xo_default_tolerance.argtypes = [ctypes.c_void_p]
        L.xo_intersect_faces.restype = ctypes.c_int64
        L.xo_intersect_faces.argtypes = [
            ctypes.c_void_p, _f64p, ctypes.c_int64, _i64p, ctypes.c_int64, ctypes.c_int,
        ]
        L.xo_intersect_faces_bruteforce.restype = ctypes.c_int64
        L.xo_intersect_faces_bruteforce.argtypes = [
            ctypes.c_void_p, _f64p, _i64p, ctypes.c_int64, ctypes.c_int,
        ]
        L.xo_fetch_result.argtypes = [ctypes.c_void_p, _i64p, _i64p, _f64p]
        L.xo_intersect_edges.restype = ctypes.c_int64
        L.xo_intersect_edges.argtypes = [ctypes.c_void_p, _f64p, ctypes.c_int64, ctypes.c_int]
        L.xo_fetch_edges.argtypes = [ctypes.c_void_p, _i64p, _i64p, _f64p]
        L.xo_locate_points.argtypes = [ctypes.c_void_p, _f64p, ctypes.c_int64, ctypes.c_double, _i64p]
        L.xo_locate_point_all.restype = ctypes.c_int
        L.xo_locate_point_all.argtypes = [
            ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, _i64p, ctypes.c_int,
        ]
        L.xo_barycentric.argtypes = [
            ctypes.c_void_p, _f64p, ctypes.c_int64, ctypes.c_double, _i64p, _f64p,
        ]
        L.xo_face_area.argtypes = [_f64p, _i64p, ctypes.c_int64, ctypes.c_int, _f64p]
        L.xo_face_centroid.argtypes = [_f64p, _i64p, ctypes.c_int64, ctypes.c_int, _f64p]
        L.xo_build_bboxes.argtypes = [_f64p, _i64p, ctypes.c_int64, ctypes.c_int, _f64p]
        L.xo_counter_clockwise.argtypes = [_f64p, _i64p, ctypes.c_int64, ctypes.c_int]
        L.xo_csr_apply.restype = ctypes.c_int
        L.xo_csr_apply.argtypes = [
            ctypes.c_int, ctypes.c_double, _i64p, _i64p, _f64p, _f64p,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _f64p, ctypes.c_int,
        ]
        L.xo_coo_gather.argtypes = [
            _i64p, _i64p, ctypes.c_int64, _f64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _f64p,
        ]
        L.xo_coo_to_csr.argtypes = [_i64p, ctypes.c_int64, ctypes.c_int64, _i64p]
        L.xo_clip_area.restype = ctypes.c_double
        L.xo_clip_area.argtypes = [_f64p, ctypes.c_int, _f64p, ctypes.c_int]
        L.xo_polygons_intersect.restype = ctypes.c_int
        L.xo_polygons_intersect.argtypes = [_f64p, ctypes.c_int, _f64p, ctypes.c_int]
        L.xo_point_in_polygon_or_on_edge.restype = ctypes.c_int
        L.xo_point_in_polygon_or_on_edge.argtypes = [
            ctypes.c_double, ctypes.c_double, _f64p, ctypes.c_int, ctypes.c_double,
        ]
        L.xo_replace_interpolated_weights.argtypes = [
            _f64p, _i64p, ctypes.c_int, _i64p, _f64p, ctypes.c_int64, _i64p, ctypes.c_int64,
        ]
        L.xo_num_threads.restype = ctypes.c_int
        L.xo_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().xo_num_threads())


def set_num_threads(n: int) -> None:
    lib().xo_set_num_threads(int(n))


def _faces64(faces, fill_value=-1) -> np.ndarray:
    f = np.array(faces, dtype=np.int64, order="C", copy=True)
    if f.ndim != 2:
        raise ValueError("faces must be (n_face, k)")
    if fill_value != -1:
        f[f == fill_value] = -1
    return f


def _xy(vertices) -> np.ndarray:
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    if v.ndim != 2 or v.shape[1] != 2:
        raise ValueError("vertices must be (n, 2)")
    return v


class CellTree2d:
    """numba_celltree.CellTree2d stand-in (the B1 seam of SURVEY.md §8b)."""

    def __init__(self, vertices, faces, fill_value=-1, n_buckets: int = 4, cells_per_leaf: int = 2):
        self.vertices = _xy(vertices)
        self.faces = _faces64(faces, fill_value)
        self.n_face, self.k = self.faces.shape
        self._h = lib().xo_celltree_create(
            _p_f64(self.vertices), self.vertices.shape[0], _p_i64(self.faces), self.n_face,
            self.k, n_buckets, cells_per_leaf,
        )

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.xo_celltree_destroy(h)
            self._h = None

    @property
    def n_nodes(self) -> int:
        return int(lib().xo_celltree_n_nodes(self._h))

    @property
    def n_candidates(self) -> int:
        """bbox candidates of the last intersect_faces call"""
        return int(lib().xo_celltree_n_candidates(self._h))

    def default_tolerance(self) -> float:
        return float(lib().xo_default_tolerance(self._h))

    def _fetch(self, nnz: int):
        ii = np.empty(nnz, dtype=np.int64)
        jj = np.empty(nnz, dtype=np.int64)
        area = np.empty(nnz, dtype=np.float64)
        if nnz:
            lib().xo_fetch_result(self._h, _p_i64(ii), _p_i64(jj), _p_f64(area))
        return ii, jj, area

    def intersect_faces(self, vertices, faces, fill_value=-1):
        v = _xy(vertices)
        f = _faces64(faces, fill_value)
        nnz = lib().xo_intersect_faces(self._h, _p_f64(v), v.shape[0], _p_i64(f), f.shape[0], f.shape[1])
        return self._fetch(int(nnz))

    def intersect_edges(self, edge_coords):
        """numba_celltree.CellTree2d.intersect_edges: (edge index ascending, face index,
        intersections (n, 2, 2)); call site /root/reference/xugrid/regrid/unstructured.py:203-215"""
        e = np.ascontiguousarray(np.asarray(edge_coords, dtype=np.float64).reshape(-1, 2, 2))
        n = int(lib().xo_intersect_edges(self._h, _p_f64(e), e.shape[0], 0))
        ii = np.empty(n, dtype=np.int64)
        jj = np.empty(n, dtype=np.int64)
        xy = np.empty((n, 2, 2), dtype=np.float64)
        if n:
            lib().xo_fetch_edges(self._h, _p_i64(ii), _p_i64(jj), _p_f64(xy))
        return ii, jj, xy

    def intersect_faces_bruteforce(self, vertices, faces, fill_value=-1):
        v = _xy(vertices)
        f = _faces64(faces, fill_value)
        nnz = lib().xo_intersect_faces_bruteforce(self._h, _p_f64(v), _p_i64(f), f.shape[0], f.shape[1])
        return self._fetch(int(nnz))

    def locate_points(self, points, tolerance: Optional[float] = None) -> np.ndarray:
        p = _xy(points)
        out = np.empty(p.shape[0], dtype=np.int64)
        tol = -1.0 if tolerance is None else float(tolerance)
        lib().xo_locate_points(self._h, _p_f64(p), p.shape[0], tol, _p_i64(out))
        return out

    def locate_point_all(self, x: float, y: float, tolerance: Optional[float] = None) -> np.ndarray:
        out = np.empty(64, dtype=np.int64)
        tol = -1.0 if tolerance is None else float(tolerance)
        n = lib().xo_locate_point_all(self._h, float(x), float(y), tol, _p_i64(out), 64)
        return out[: min(n, 64)].copy()

    def compute_barycentric_weights(self, points, tolerance: Optional[float] = None):
        p = _xy(points)
        idx = np.empty(p.shape[0], dtype=np.int64)
        w = np.zeros((p.shape[0], self.k), dtype=np.float64)
        tol = -1.0 if tolerance is None else float(tolerance)
        lib().xo_barycentric(self._h, _p_f64(p), p.shape[0], tol, _p_i64(idx), _p_f64(w))
        return idx, w


def face_area(vertices, faces, fill_value=-1) -> np.ndarray:
    v, f = _xy(vertices), _faces64(faces, fill_value)
    out = np.empty(f.shape[0], dtype=np.float64)
    lib().xo_face_area(_p_f64(v), _p_i64(f), f.shape[0], f.shape[1], _p_f64(out))
    return out


def face_centroids(vertices, faces, fill_value=-1) -> np.ndarray:
    v, f = _xy(vertices), _faces64(faces, fill_value)
    out = np.empty((f.shape[0], 2), dtype=np.float64)
    lib().xo_face_centroid(_p_f64(v), _p_i64(f), f.shape[0], f.shape[1], _p_f64(out))
    return out


def build_bboxes(vertices, faces, fill_value=-1) -> np.ndarray:
    v, f = _xy(vertices), _faces64(faces, fill_value)
    out = np.empty((f.shape[0], 4), dtype=np.float64)
    lib().xo_build_bboxes(_p_f64(v), _p_i64(f), f.shape[0], f.shape[1], _p_f64(out))
    return out


def counter_clockwise(vertices, faces, fill_value=-1) -> np.ndarray:
    v, f = _xy(vertices), _faces64(faces, fill_value)
    lib().xo_counter_clockwise(_p_f64(v), _p_i64(f), f.shape[0], f.shape[1])
    return f


def clip_area(subject, clipper) -> float:
    a = np.ascontiguousarray(subject, dtype=np.float64)
    b = np.ascontiguousarray(clipper, dtype=np.float64)
    return float(lib().xo_clip_area(_p_f64(a), a.shape[0], _p_f64(b), b.shape[0]))


def polygons_intersect(a, b) -> bool:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return bool(lib().xo_polygons_intersect(_p_f64(a), a.shape[0], _p_f64(b), b.shape[0]))


def point_in_polygon_or_on_edge(x, y, poly, tolerance) -> bool:
    a = np.ascontiguousarray(poly, dtype=np.float64)
    return bool(lib().xo_point_in_polygon_or_on_edge(float(x), float(y), _p_f64(a), a.shape[0], float(tolerance)))


def coo_to_csr(row, n: int) -> np.ndarray:
    r = np.ascontiguousarray(row, dtype=np.int64)
    indptr = np.empty(n + 1, dtype=np.int64)
    lib().xo_coo_to_csr(_p_i64(r), r.size, n, _p_i64(indptr))
    return indptr


def csr_apply(method, source, data, indices, indptr, n: int, percentile: Optional[float] = None,
              parallel_rows: bool = False) -> np.ndarray:
    """
    regridder.py:41-67 on CPU.  ``method`` is a reduce.py name or "percentile"
    with ``percentile`` given.  ``source`` is (n_extra, m); returns (n_extra, n)
    float64 with NaN for empty rows.
    """
    if method == "percentile":
        code, p = 7, float(percentile)
    else:
        code, p = METHODS[method]
    src = np.ascontiguousarray(source, dtype=np.float64)
    if src.ndim == 1:
        src = src[None, :]
    n_extra, m = src.shape
    d = np.ascontiguousarray(data, dtype=np.float64)
    ind = np.ascontiguousarray(indices, dtype=np.int64)
    ptr = np.ascontiguousarray(indptr, dtype=np.int64)
    out = np.empty((n_extra, n), dtype=np.float64)
    lib().xo_csr_apply(code, p, _p_i64(ptr), _p_i64(ind), _p_f64(d), _p_f64(src), n_extra, m, n,
                       _p_f64(out), 1 if parallel_rows else 0)
    return out


def coo_gather(source, row, col, size: int) -> np.ndarray:
    src = np.ascontiguousarray(source, dtype=np.float64)
    if src.ndim == 1:
        src = src[None, :]
    n_extra, m = src.shape
    r = np.ascontiguousarray(row, dtype=np.int64)
    c = np.ascontiguousarray(col, dtype=np.int64)
    out = np.empty((n_extra, size), dtype=np.float64)
    lib().xo_coo_gather(_p_i64(r), _p_i64(c), r.size, _p_f64(src), n_extra, m, size, _p_f64(out))
    return out


def replace_interpolated_weights(vertices, faces, face_index, weights, node_to_node_map,
                                 node_index_threshold: int) -> None:
    v, f = _xy(vertices), _faces64(faces)
    fi = np.ascontiguousarray(face_index, dtype=np.int64)
    nn = np.ascontiguousarray(node_to_node_map, dtype=np.int64)
    assert weights.dtype == np.float64 and weights.flags.c_contiguous
    lib().xo_replace_interpolated_weights(
        _p_f64(v), _p_i64(f), f.shape[1], _p_i64(fi), _p_f64(weights), fi.size, _p_i64(nn),
        int(node_index_threshold),
    )


def overlap_triplets(source_vertices, source_faces, target_vertices, target_faces,
                     relative: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """UnstructuredGrid2d.overlap (unstructured.py:109-135): returns
    (source_index, target_index, weights) with target_index ascending."""
    tree = CellTree2d(source_vertices, source_faces, -1)
    ti, si, w = tree.intersect_faces(target_vertices, target_faces, -1)
    if relative:
        w = w / face_area(source_vertices, source_faces)[si]
    return si, ti, w
