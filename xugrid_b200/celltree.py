"""
GPU spatial index with the interface xugrid expects from
``numba_celltree.CellTree2d`` (the object returned by ``Ugrid2d.celltree``,
/root/reference/xugrid/ugrid/ugrid2d.py:908-921):

* ``intersect_faces(vertices, faces, fill_value)`` -- used by
  ``UnstructuredGrid2d.overlap`` (xugrid/regrid/unstructured.py:124-132)
* ``locate_points(points, tolerance)`` -- unstructured.py:138-139,
  xugrid/ugrid/ugridbase.py:1305-1323, xugrid/ugrid/ugrid2d.py:1099
* ``compute_barycentric_weights(points, tolerance)`` -- ugrid2d.py:1054-1078

All work happens in libxgregrid.so (hand-written sm_100a CUDA); torch holds
the device memory.  numpy in -> numpy out (int64 / float64, like the
reference); the ``*_device`` variants keep everything in HBM.
"""

from __future__ import annotations

import ctypes
from typing import NamedTuple, Optional, Tuple, Union

import numpy as np
import torch

from . import _native

ArrayLike = Union[np.ndarray, torch.Tensor]


def _device(device=None) -> torch.device:
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def to_device(a: ArrayLike, dtype: Optional[torch.dtype], device: torch.device) -> torch.Tensor:
    """numpy / torch (host or device) -> contiguous device tensor; pinned host
    tensors are copied asynchronously."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        if t.is_cuda:
            t = t.to(dtype)
        else:
            t = t.to(device, non_blocking=True).to(dtype)
    if not t.is_cuda:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def faces_to_device(faces: ArrayLike, fill_value: int, device: torch.device) -> torch.Tensor:
    """(n_face, k) connectivity as int32 or int64 on the device, fill -> -1."""
    if isinstance(faces, torch.Tensor):
        t = faces
    else:
        a = np.asarray(faces)
        if a.dtype not in (np.int32, np.int64):
            a = a.astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype not in (torch.int32, torch.int64):
        t = t.to(torch.int64)
    if t.dim() != 2:
        raise ValueError("faces must have shape (n_face, n_max_vertex)")
    t = to_device(t, None, device)
    if fill_value != -1:
        t = torch.where(t == fill_value, torch.full_like(t, -1), t)
    return t


PHASES = ("pack", "traverse", "sat", "clip", "rowptr")


class OverlapStats(NamedTuple):
    targets: int
    node_visits: int
    leaf_tests: int
    clip_pairs: int
    nnz: int
    candidates: int
    walks: int = 1  # full walks of the index (2 = the pair arena was too small the first time)


class PendingOverlap:
    """An overlap job that was enqueued without waiting (CellTree2d.intersect_faces_into)."""

    def __init__(self, tree: "CellTree2d", job):
        self._tree, self._job = tree, job
        self._res = None

    def result(self) -> Tuple[int, int, bool]:
        """(nnz of the job, row pointer behind its last row, complete?) -- waits for the job."""
        if self._res is None:
            L = _native.load_library()
            nnz, end, ok = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
            try:
                _native.check(L.xg_overlap_result(self._job, ctypes.byref(nnz), ctypes.byref(end), ctypes.byref(ok)))
                st = (ctypes.c_int64 * 8)()
                _native.check(L.xg_overlap_stats(self._job, st))
                self.stats = OverlapStats(*[int(st[i]) for i in range(7)])
                self.launched_as_graph = bool(st[7])
                ms = (ctypes.c_double * 8)()
                _native.check(L.xg_overlap_timings(self._job, ms))
                self.phase_ms = dict(zip(PHASES, [float(ms[i]) for i in range(len(PHASES))]))
                self._tree.last_stats, self._tree.last_phase_ms = self.stats, self.phase_ms
            finally:
                L.xg_overlap_abort(self._job)
                self._job = None
            self._res = (int(nnz.value), int(end.value), bool(ok.value))
        return self._res

    def __del__(self):
        job = getattr(self, "_job", None)
        if job is not None and job.value:
            try:
                _native.load_library().xg_overlap_abort(job)
            except Exception:
                pass
            self._job = None


class CellTree2d:
    def __init__(self, vertices: ArrayLike, faces: ArrayLike, fill_value: int = -1, device=None,
                 clip_box=None, clip_targets=None, clip_target_ids=None):
        """``clip_box=(xmin, xmax, ymin, ymax)`` indexes only the faces whose bounding
        box meets the box (used by ranks that own a range of target faces);
        ``clip_targets=(xy, faces, begin, end)`` (device tensors of another mesh) takes
        the box from the bounds of those target faces, reduced on the device -- no
        host round trip; ``clip_target_ids=(xy, faces, ids)`` the same for a list of target
        face numbers (int32 device tensor, e.g. from ``distributed.curve_shard``)."""
        L = _native.lib()
        self.device = _device(device)
        with torch.cuda.device(self.device):
            self.vertices = to_device(vertices, torch.float64, self.device)
            if self.vertices.dim() != 2 or self.vertices.shape[1] != 2:
                raise ValueError("vertices must have shape (n_vertex, 2)")
            self.faces = faces_to_device(faces, fill_value, self.device)
            self.n_face, self.k = int(self.faces.shape[0]), int(self.faces.shape[1])
            self.n_node = int(self.vertices.shape[0])
            h = ctypes.c_void_p()
            if clip_target_ids is not None:
                t_xy, t_faces, t_ids = clip_target_ids
                _native.check(
                    L.xg_index_create_for_target_ids(
                        _native.ptr(self.vertices), self.n_node, _native.ptr(self.faces),
                        self.faces.element_size(), self.n_face, self.k, _native.ptr(t_xy, torch.float64),
                        _native.ptr(t_faces), t_faces.element_size(), _native.ptr(t_ids, torch.int32),
                        int(t_ids.numel()), int(t_faces.shape[1]), _native.stream_ptr(), ctypes.byref(h),
                    )
                )
            elif clip_targets is not None:
                t_xy, t_faces, t_begin, t_end = clip_targets
                _native.check(
                    L.xg_index_create_for_targets(
                        _native.ptr(self.vertices), self.n_node, _native.ptr(self.faces),
                        self.faces.element_size(), self.n_face, self.k, _native.ptr(t_xy, torch.float64),
                        _native.ptr(t_faces), t_faces.element_size(), int(t_begin), int(t_end),
                        int(t_faces.shape[1]), _native.stream_ptr(), ctypes.byref(h),
                    )
                )
            else:
                box = None if clip_box is None else (ctypes.c_double * 4)(*[float(v) for v in clip_box])
                _native.check(
                    L.xg_index_create_clipped(
                        _native.ptr(self.vertices), self.n_node, _native.ptr(self.faces),
                        self.faces.element_size(), self.n_face, self.k, box, _native.stream_ptr(),
                        ctypes.byref(h),
                    )
                )
        self._h = h
        self._info = None
        self._tolerance = None
        self.last_stats: Optional[OverlapStats] = None
        self.last_phase_ms: dict = {}

    # sizes are known when the build has been enqueued; the default tolerance is reduced on
    # the device and only waited for when somebody asks (an overlap job never does)
    def _index_info(self):
        if self._info is None:
            info = (ctypes.c_int64 * 8)()
            _native.check(_native.load_library().xg_index_info(self._h, info, None))
            self._info = [int(v) for v in info]
        return self._info

    @property
    def n_indexed(self) -> int:
        return self._index_info()[0]

    @property
    def n_levels(self) -> int:
        return self._index_info()[2]

    @property
    def device_bytes(self) -> int:
        return self._index_info()[3]

    @property
    def default_tolerance(self) -> float:
        if self._tolerance is None:
            tol = ctypes.c_double()
            _native.check(_native.load_library().xg_index_info(self._h, None, ctypes.byref(tol)))
            self._tolerance = float(tol.value)
        return self._tolerance

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _native.load_library().xg_index_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ overlap
    def intersect_faces_csr(
        self,
        vertices: ArrayLike,
        faces: ArrayLike,
        fill_value: int = -1,
        t_begin: int = 0,
        t_end: Optional[int] = None,
        relative: bool = False,
        ids: Optional[torch.Tensor] = None,
    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """
        CSR over the target rows [t_begin, t_end): (indptr int64, indices int32,
        area float64) device tensors, rows in ascending source index.  With ``ids``
        (int32 device tensor of target face numbers) row i belongs to face ``ids[i]``.
        """
        L = _native.lib()
        with torch.cuda.device(self.device):
            xy = to_device(vertices, torch.float64, self.device)
            fc = faces_to_device(faces, fill_value, self.device)
            n_t, k_t = int(fc.shape[0]), int(fc.shape[1])
            if t_end is None:
                t_end = n_t
            rows = t_end - t_begin if ids is None else int(ids.numel())
            indptr = torch.empty(rows + 1, dtype=torch.int64, device=self.device)
            nnz = ctypes.c_int64()
            job = ctypes.c_void_p()
            s = _native.stream_ptr()
            if ids is None:
                _native.check(
                    L.xg_overlap_begin(
                        self._h, _native.ptr(xy), int(xy.shape[0]), _native.ptr(fc), fc.element_size(),
                        n_t, k_t, t_begin, t_end, _native.ptr(indptr), ctypes.byref(nnz), s,
                        ctypes.byref(job),
                    )
                )
            else:
                _native.check(
                    L.xg_overlap_begin_ids(
                        self._h, _native.ptr(xy), int(xy.shape[0]), _native.ptr(fc), fc.element_size(),
                        n_t, k_t, _native.ptr(ids, torch.int32), rows, _native.ptr(indptr),
                        ctypes.byref(nnz), s, ctypes.byref(job),
                    )
                )
            try:
                st = (ctypes.c_int64 * 8)()
                _native.check(L.xg_overlap_stats(job, st))
                self.last_stats = OverlapStats(*[int(st[i]) for i in range(7)])
                ms = (ctypes.c_double * 8)()
                _native.check(L.xg_overlap_timings(job, ms))
                self.last_phase_ms = dict(zip(PHASES, [float(ms[i]) for i in range(len(PHASES))]))
                indices = torch.empty(nnz.value, dtype=torch.int32, device=self.device)
                area = torch.empty(nnz.value, dtype=torch.float64, device=self.device)
            except BaseException:
                # the job holds the pair / survivor / provisional arenas (several GB at 10 M
                # faces) and is only released by xg_overlap_finish: do not leak it
                L.xg_overlap_abort(job)
                raise
            _native.check(
                L.xg_overlap_finish(
                    job, _native.ptr(indptr), _native.ptr(indices), _native.ptr(area),
                    1 if relative else 0, s,
                )
            )
        return indptr, indices, area

    def intersect_faces_into(
        self,
        vertices: torch.Tensor,
        faces: torch.Tensor,
        indptr: torch.Tensor,
        indices: torch.Tensor,
        area: torch.Tensor,
        t_begin: int = 0,
        t_end: Optional[int] = None,
        relative: bool = False,
        ids: Optional[torch.Tensor] = None,
        chained: bool = False,
        graph: bool = False,
    ) -> "PendingOverlap":
        """
        The same CSR rows written into arrays the caller sized by an upper bound, with
        NOTHING waited for (``xg_overlap_run``): ``indptr`` is the (rows + 1) slice of the
        row pointer these target rows own, ``indices`` / ``area`` the whole entry arrays.
        ``chained=True``: ``indptr[0]`` already holds (on the device) the offset of the first
        entry, i.e. the last row pointer of the rows before -- consecutive target ranges
        fill one matrix.  ``graph=True``: the job goes out as one CUDA graph launch (what a
        regridder asks for: under a mesh upload every small launch of a job is slowed down by
        the traffic on the PCIe link); no per-phase timings then.  Returns a handle;
        ``.result()`` waits and tells whether the job is complete (arena and arrays were large
        enough).
        """
        L = _native.lib()
        with torch.cuda.device(self.device):
            n_t, k_t = int(faces.shape[0]), int(faces.shape[1])
            if t_end is None:
                t_end = n_t
            if ids is not None:
                t_begin, t_end = 0, int(ids.numel())
            rows = t_end - t_begin
            if indptr.numel() != rows + 1 or indices.numel() != area.numel():
                raise ValueError("indptr must have rows + 1 entries; indices and area the same capacity")
            job = ctypes.c_void_p()
            _native.check(
                L.xg_overlap_run(
                    self._h, _native.ptr(vertices, torch.float64), int(vertices.shape[0]), _native.ptr(faces),
                    faces.element_size(), n_t, k_t, _native.ptr(ids, torch.int32) if ids is not None else None,
                    t_begin, t_end, (1 if chained else 0) | (2 if graph else 0), _native.ptr(indptr, torch.int64),
                    _native.ptr(indices, torch.int32), _native.ptr(area, torch.float64), int(area.numel()),
                    1 if relative else 0, _native.stream_ptr(), ctypes.byref(job),
                )
            )
        return PendingOverlap(self, job)

    def intersect_faces(self, vertices: ArrayLike, faces: ArrayLike, fill_value: int = -1):
        """numba_celltree-compatible triplets: (face index of the query mesh
        ascending, face index of this tree, area) as numpy int64/int64/float64."""
        indptr, indices, area = self.intersect_faces_csr(vertices, faces, fill_value)
        counts = indptr[1:] - indptr[:-1]
        rows = torch.repeat_interleave(
            torch.arange(counts.numel(), device=self.device, dtype=torch.int64), counts
        )
        return rows.cpu().numpy(), indices.to(torch.int64).cpu().numpy(), area.cpu().numpy()

    # ------------------------------------------------------------------ segments
    def intersect_edges_device(self, edge_coords: ArrayLike):
        """(edge index int32, face index int32, intersections (n, 2, 2) float64) device
        tensors in ascending (edge, face) order."""
        L = _native.lib()
        with torch.cuda.device(self.device):
            e = to_device(edge_coords, torch.float64, self.device).reshape(-1, 2, 2).contiguous()
            n_edge = int(e.shape[0])
            cap = 4 * n_edge + 1024
            while True:
                ei = torch.empty(cap, dtype=torch.int32, device=self.device)
                fi = torch.empty(cap, dtype=torch.int32, device=self.device)
                xy = torch.empty((cap, 2, 2), dtype=torch.float64, device=self.device)
                count = ctypes.c_int64()
                _native.check(
                    L.xg_intersect_edges(self._h, _native.ptr(e), n_edge, cap, _native.ptr(ei), _native.ptr(fi),
                                         _native.ptr(xy), ctypes.byref(count), _native.stream_ptr())
                )
                if count.value <= cap:
                    break
                cap = int(count.value)  # denser than guessed: the first pass counted exactly
            n = int(count.value)
            ei, fi, xy = ei[:n], fi[:n], xy[:n]
            order = torch.argsort(ei.to(torch.int64) * max(self.n_face, 1) + fi.to(torch.int64))
            return ei[order], fi[order], xy[order]

    def intersect_edges(self, edge_coords: ArrayLike):
        """numba_celltree-compatible: (edge index ascending, face index, intersections
        (n, 2, 2)) as numpy int64/int64/float64 (call site: unstructured.py:203-215)."""
        ei, fi, xy = self.intersect_edges_device(edge_coords)
        return ei.to(torch.int64).cpu().numpy(), fi.to(torch.int64).cpu().numpy(), xy.cpu().numpy()

    # ------------------------------------------------------------------ points
    def locate_points_device(self, points: ArrayLike, tolerance: Optional[float] = None,
                             any_face: bool = False) -> torch.Tensor:
        """``any_face=True``: only "inside or not" matters to the caller -- the answer is SOME
        face that holds the point (or -1), found by a walk that stops at the first hit."""
        L = _native.lib()
        with torch.cuda.device(self.device):
            p = to_device(points, torch.float64, self.device)
            if p.dim() != 2 or p.shape[1] != 2:
                raise ValueError("points must have shape (n_point, 2)")
            out = torch.empty(p.shape[0], dtype=torch.int32, device=self.device)
            tol = -1.0 if tolerance is None else float(tolerance)
            fn = L.xg_locate_any if any_face else L.xg_locate_points
            _native.check(fn(self._h, _native.ptr(p), int(p.shape[0]), tol, _native.ptr(out), _native.stream_ptr()))
        return out

    def locate_points(self, points: ArrayLike, tolerance: Optional[float] = None) -> np.ndarray:
        return self.locate_points_device(points, tolerance).to(torch.int64).cpu().numpy()

    def compute_barycentric_weights_device(self, points: ArrayLike, tolerance: Optional[float] = None):
        L = _native.lib()
        with torch.cuda.device(self.device):
            p = to_device(points, torch.float64, self.device)
            if p.dim() != 2 or p.shape[1] != 2:
                raise ValueError("points must have shape (n_point, 2)")
            n = int(p.shape[0])
            idx = torch.empty(n, dtype=torch.int32, device=self.device)
            w = torch.empty((n, self.k), dtype=torch.float64, device=self.device)
            tol = -1.0 if tolerance is None else float(tolerance)
            _native.check(
                L.xg_barycentric(self._h, _native.ptr(p), n, tol, _native.ptr(idx), _native.ptr(w),
                                 _native.stream_ptr())
            )
        return idx, w

    def compute_barycentric_weights(self, points: ArrayLike, tolerance: Optional[float] = None):
        idx, w = self.compute_barycentric_weights_device(points, tolerance)
        return idx.to(torch.int64).cpu().numpy(), w.cpu().numpy()


def face_area_device(vertices: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """connectivity.area (xugrid/ugrid/connectivity.py:615-633) on the device."""
    L = _native.lib()
    out = torch.empty(faces.shape[0], dtype=torch.float64, device=faces.device)
    with torch.cuda.device(faces.device):
        _native.check(
            L.xg_face_area(_native.ptr(vertices), _native.ptr(faces), faces.element_size(),
                           int(faces.shape[0]), int(faces.shape[1]), _native.ptr(out),
                           _native.stream_ptr())
        )
    return out


def face_centroid_device(vertices: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """connectivity.centroids (xugrid/ugrid/connectivity.py:636-664) on the device."""
    L = _native.lib()
    out = torch.empty((faces.shape[0], 2), dtype=torch.float64, device=faces.device)
    with torch.cuda.device(faces.device):
        _native.check(
            L.xg_face_centroid(_native.ptr(vertices), _native.ptr(faces), faces.element_size(),
                               int(faces.shape[0]), int(faces.shape[1]), _native.ptr(out),
                               _native.stream_ptr())
        )
    return out
