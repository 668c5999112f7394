"""
Minimal mesh containers for the regridding path.

``Ugrid2d`` keeps only what the hot path of xugrid's class of the same name
touches (/root/reference/xugrid/ugrid/ugrid2d.py:87-109 constructor
normalisation, ``node_coordinates`` ugridbase.py:576-579, ``centroids``
ugrid2d.py:544-559, ``area`` :575-584, ``celltree`` :908-921,
``locate_points`` ugridbase.py:1305-1323, ``compute_barycentric_weights``
ugrid2d.py:1054-1078, ``rasterize_like`` :1080-1101,
``from_structured_bounds`` :1973-2034).  Geometry is evaluated on the GPU.

``StructuredGrid2d`` covers only the promotion of a raster to a quad mesh
(/root/reference/xugrid/regrid/structured.py:24-100,489-501); the separable
structured-to-structured path of the reference is out of scope and such pairs
are regridded as quad meshes.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _native
from .celltree import CellTree2d, face_area_device, face_centroid_device, faces_to_device, to_device

FILL_VALUE = -1


class Ugrid2d:
    def __init__(self, node_x, node_y, fill_value: int, face_node_connectivity,
                 name: str = "mesh2d", start_index: int = 0, device=None):
        self.node_x = np.ascontiguousarray(_host(node_x), dtype=np.float64)
        self.node_y = np.ascontiguousarray(_host(node_y), dtype=np.float64)
        if self.node_x.ndim != 1 or self.node_x.shape != self.node_y.shape:
            raise ValueError("node_x and node_y must be 1D arrays of equal length")
        if not isinstance(face_node_connectivity, (np.ndarray, torch.Tensor)):
            raise TypeError(
                "face_node_connectivity should be an array of integers or a sparse matrix"
            )
        conn = np.array(_host(face_node_connectivity), copy=True)
        if conn.dtype not in (np.int32, np.int64):
            conn = conn.astype(np.int64)
        if conn.ndim != 2:
            raise ValueError("face_node_connectivity must be (n_face, n_max_node)")
        # ugrid2d.py:103-109: fill value -> -1, indices 0-based
        if fill_value != FILL_VALUE or start_index != 0:
            is_fill = conn == fill_value
            if start_index != 0:
                conn[~is_fill] -= start_index
            if fill_value != FILL_VALUE:
                conn[is_fill] = FILL_VALUE
        self.face_node_connectivity = conn
        self.fill_value = FILL_VALUE
        self.start_index = 0
        self.name = name
        self.device = device
        self._clear()

    def _clear(self):
        self._celltree = None
        self._xy_d = None
        self._faces_d = None
        self._centroids = None
        self._area = None
        self._xy_pin = None
        self._faces_pin = None
        self._faces_pin_u24 = None
        self._packed_d = None
        self._ready = None
        self._xy_ready = None
        self._chunks = None
        self._face_range = None

    # ---- basic properties -------------------------------------------------------
    @property
    def n_node(self) -> int:
        return self.node_x.size

    @property
    def n_face(self) -> int:
        return self.face_node_connectivity.shape[0]

    @property
    def n_max_node_per_face(self) -> int:
        return self.face_node_connectivity.shape[1]

    @property
    def face_dimension(self) -> str:
        return f"{self.name}_nFaces"

    @property
    def core_dimension(self) -> str:
        return self.face_dimension

    @property
    def node_coordinates(self) -> np.ndarray:
        return np.column_stack([self.node_x, self.node_y])

    @property
    def bounds(self) -> Tuple[float, float, float, float]:
        return (float(self.node_x.min()), float(self.node_y.min()),
                float(self.node_x.max()), float(self.node_y.max()))

    # ---- device mirrors ------------------------------------------------------------
    def _dev(self) -> torch.device:
        if self.device is None:
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device(self.device)

    # meshes of at least this many faces (and fewer than 2^24 - 1 nodes) cross PCIe with 3 bytes
    # per connectivity entry (xg_unpack_u24 widens them on the device)
    _PACK_MIN_FACES = 1 << 20

    def _packs_u24(self) -> bool:
        return self.n_node < (1 << 24) - 1 and self.n_face >= self._PACK_MIN_FACES

    def _stage(self, packed: Optional[bool] = None) -> None:
        """Pinned host staging copies, made once per grid.  Connectivity is staged as
        int32 whenever the node count allows it (the kernels take either width), and for
        large meshes of fewer than 2^24 - 1 nodes also as 24-bit little-endian entries
        (0xFFFFFF = fill): 3 instead of 8 bytes per entry over PCIe on every later upload.
        ``packed``: which copy is needed now (None: what ``prefetch`` will use)."""
        if self._xy_pin is None:
            self._xy_pin = torch.from_numpy(self.node_coordinates).pin_memory()
        if packed is None:
            packed = self._packs_u24()
        if packed and self._faces_pin_u24 is None:
            conn = self.face_node_connectivity
            c32 = conn if conn.dtype == np.int32 else conn.astype(np.int32)
            # (little-endian int32: the low three bytes; -1 becomes 0xFFFFFF)
            b = np.ascontiguousarray(np.ascontiguousarray(c32).view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1)
            self._faces_pin_u24 = torch.from_numpy(b).pin_memory()
        if not packed and self._faces_pin is None:
            conn = self.face_node_connectivity
            if conn.dtype != np.int32 and self.n_node < (1 << 31):
                conn = conn.astype(np.int32)
            self._faces_pin = torch.from_numpy(np.ascontiguousarray(conn)).pin_memory()

    def upload_nbytes(self, face_range: Optional[Tuple[int, int]] = None, sharded_over: int = 1) -> int:
        """Bytes one (re-)upload of the mesh moves over PCIe: coordinates + connectivity as it is
        staged (``sharded_over`` ranks share the replicated arrays; a face range is not shared)."""
        k = self.n_max_node_per_face
        xy = self.n_node * 16
        if face_range is None:
            per = 3 if self._packs_u24() else (4 if self.n_node < (1 << 31) else 8)
            return (xy + self.n_face * k * per) // sharded_over
        packed = self._packs_u24()
        per = 3 if packed else (4 if self.n_node < (1 << 31) else 8)
        return xy // sharded_over + max(0, face_range[1] - face_range[0]) * k * per

    def prefetch(self, n_chunks: int = 1, face_range: Optional[Tuple[int, int]] = None,
                 sharded: bool = False) -> None:
        """Start the host-to-device copy of the mesh on the side (copy) stream.
        The first call pins a host staging copy; later uploads (after
        ``release_device``) are asynchronous DMA transfers that overlap with
        kernels on the compute stream.  With ``n_chunks > 1`` the connectivity
        goes up in contiguous face ranges, each with its own event, so that a
        consumer can start on the first range while the rest is in flight.
        ``face_range=(b, e)``: only that range of the connectivity is uploaded
        (a rank that owns a range of target faces needs nothing else).
        ``sharded=True`` (inside an initialised process group whose ranks all
        hold the same host mesh): every rank uploads 1/world of the arrays over
        its own PCIe link and the full mesh is assembled with an all-gather over
        NVLink."""
        if self._xy_d is not None:
            return
        L = _native.lib()  # raises when there is no GPU: no CPU fallback
        dev = self._dev()
        whole_sharded = sharded and face_range is None
        packed = self._packs_u24()
        self._stage(packed)
        copy_stream = _copy_stream(dev)
        n = self.n_face
        k = self.n_max_node_per_face
        if sharded:
            from .distributed import sharded_upload
        if sharded and face_range is None:
            with torch.cuda.stream(copy_stream):
                self._xy_d = sharded_upload(self._xy_pin, dev)
                if packed:
                    # 1/world of the 24-bit entries per rank, assembled over NVLink, widened here
                    self._packed_d = sharded_upload(self._faces_pin_u24, dev)
                    self._faces_d = torch.empty((n, k), dtype=torch.int32, device=dev)
                    _native.check(L.xg_unpack_u24(self._packed_d.data_ptr(), n * k, self._faces_d.data_ptr(),
                                                  _native.stream_ptr()))
                else:
                    self._faces_d = sharded_upload(self._faces_pin, dev)
                self._xy_ready = torch.cuda.Event(enable_timing=True)
                self._xy_ready.record(copy_stream)
                self._upload_events = [("mesh replicated", self._xy_ready)]
                self._chunks = [(0, n, self._xy_ready)]
                self._ready = self._xy_ready
                self._face_range = None
            return
        fb, fe = (0, n) if face_range is None else (max(0, int(face_range[0])), min(n, int(face_range[1])))
        m = max(0, fe - fb)
        if isinstance(n_chunks, (tuple, list)):
            # cumulative fractions of the range, e.g. (0.125, 0.5, 1.0): a small first chunk
            # lets the consumer start early, large later chunks keep its fixed costs low
            fr = [min(1.0, max(0.0, float(f))) for f in n_chunks]
            bounds = sorted(set([fb] + [fb + int(round(m * f)) for f in fr] + [fe]))
        else:
            n_chunks = max(1, min(int(n_chunks), m)) if m > 0 else 1
            bounds = [fb + (m * c) // n_chunks for c in range(n_chunks + 1)]
        if packed:
            # groups of 4 entries: every piece starts at a multiple of 4 faces (a few rows before
            # the range may come along)
            bounds = sorted(set([bounds[0] - bounds[0] % 4] + [x - x % 4 for x in bounds[1:-1]] + [bounds[-1]]))
        with torch.cuda.stream(copy_stream):
            if sharded:
                self._xy_d = sharded_upload(self._xy_pin, dev)
            else:
                self._xy_d = self._xy_pin.to(dev, non_blocking=True)
            self._xy_ready = torch.cuda.Event(enable_timing=True)
            self._xy_ready.record(copy_stream)
            # (name, event) of every upload milestone, for scripts/e2e_timeline.py
            self._upload_events = [("coordinates uploaded", self._xy_ready)]
            if packed:
                self._faces_d = torch.empty((n, k), dtype=torch.int32, device=dev)
                self._packed_d = torch.empty(n * k * 3, dtype=torch.uint8, device=dev)
            else:
                self._faces_d = torch.empty(self._faces_pin.shape, dtype=self._faces_pin.dtype, device=dev)
            self._chunks = []
            for b, e in zip(bounds[:-1], bounds[1:]):
                self._upload_faces(L, b, e)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(copy_stream)
                self._chunks.append((b, e, ev))
                self._upload_events.append((f"faces [{b},{e}) uploaded", ev))
            self._ready = self._chunks[-1][2] if self._chunks else self._xy_ready
            self._face_range = None if (fb, fe) == (0, n) else (fb, fe)

    def _upload_faces(self, L, b: int, e: int) -> None:
        """rows [b, e) of the connectivity on the current (copy) stream; b a multiple of 4 when
        the packed copy travels"""
        if e <= b:
            return
        if self._packed_d is None:
            self._faces_d[b:e].copy_(self._faces_pin[b:e], non_blocking=True)
            return
        k = self.n_max_node_per_face
        lo, hi = b * k * 3, e * k * 3
        self._packed_d[lo:hi].copy_(self._faces_pin_u24[lo:hi], non_blocking=True)
        _native.check(L.xg_unpack_u24(self._packed_d.data_ptr() + lo, (e - b) * k,
                                      self._faces_d.data_ptr() + b * k * 4, _native.stream_ptr()))

    def _complete_upload(self) -> None:
        """After a range-limited prefetch: bring up the rest of the connectivity."""
        if self._face_range is None:
            return
        fb, fe = self._face_range
        copy_stream = _copy_stream(self._dev())
        L = _native.lib()
        with torch.cuda.stream(copy_stream):
            self._upload_faces(L, 0, fb)
            self._upload_faces(L, fe - fe % 4 if self._packed_d is not None else fe, self.n_face)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        torch.cuda.current_stream(self._xy_d.device).wait_event(ev)
        self._face_range = None

    def _adopt(self, cur) -> None:
        # tensors were allocated on the copy stream: tell the allocator who uses them
        self._xy_d.record_stream(cur)
        self._faces_d.record_stream(cur)

    def device_arrays(self, face_range: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(n_node, 2) float64 and (n_face, k) int tensors in HBM (cached).  With
        ``face_range`` the caller promises to touch only those rows of the
        connectivity (the rest may not have been uploaded)."""
        self.prefetch()
        if self._face_range is not None and (
            face_range is None or face_range[0] < self._face_range[0] or face_range[1] > self._face_range[1]
        ):
            self._complete_upload()
        if self._ready is not None:
            cur = torch.cuda.current_stream(self._xy_d.device)
            cur.wait_event(self._ready)
            self._adopt(cur)
            self._ready = None
            self._chunks = None
        return self._xy_d, self._faces_d

    def device_chunks(self):
        """(xy, faces, [(begin, end)]) with the compute stream made to wait only
        for the coordinates; call ``wait_chunk(i)`` before using face range i.
        None when the mesh is already fully resident."""
        self.prefetch()
        if self._ready is None or not self._chunks or len(self._chunks) < 2 or self._face_range is not None:
            return None
        cur = torch.cuda.current_stream(self._xy_d.device)
        cur.wait_event(self._xy_ready)
        self._adopt(cur)
        return self._xy_d, self._faces_d, [(b, e) for b, e, _ in self._chunks]

    def wait_chunk(self, i: int) -> None:
        cur = torch.cuda.current_stream(self._xy_d.device)
        cur.wait_event(self._chunks[i][2])
        if i == len(self._chunks) - 1:
            self._ready = None

    def release_device(self) -> None:
        """Drop the device mirrors and the spatial index (the pinned host
        staging copy is kept); the next use uploads again."""
        self._celltree = None
        self._xy_d = None
        self._faces_d = None
        self._packed_d = None
        self._ready = None
        self._xy_ready = None
        self._chunks = None
        self._face_range = None

    @property
    def celltree(self) -> CellTree2d:
        if self._celltree is None:
            xy, faces = self.device_arrays()
            self._celltree = CellTree2d(xy, faces, FILL_VALUE, device=self._dev())
        return self._celltree

    @property
    def centroids_device(self) -> torch.Tensor:
        xy, faces = self.device_arrays()
        return face_centroid_device(xy, faces)

    @property
    def centroids(self) -> np.ndarray:
        if self._centroids is None:
            self._centroids = self.centroids_device.cpu().numpy()
        return self._centroids

    @property
    def area_device(self) -> torch.Tensor:
        xy, faces = self.device_arrays()
        return face_area_device(xy, faces)

    @property
    def area(self) -> np.ndarray:
        if self._area is None:
            self._area = self.area_device.cpu().numpy()
        return self._area

    # ---- queries ----------------------------------------------------------------------
    def locate_points(self, points, tolerance: Optional[float] = None) -> np.ndarray:
        return self.celltree.locate_points(points, tolerance)

    def compute_barycentric_weights(self, points, tolerance: Optional[float] = None):
        return self.celltree.compute_barycentric_weights(points, tolerance)

    def sel_points(self, obj, x, y, method: Optional[str] = None, out_of_bounds: str = "warn",
                   fill_value=np.nan, tolerance: Optional[float] = None) -> "PointSelection":
        """
        Values of face data at points (/root/reference/xugrid/ugrid/ugridbase.py:1125-1259):
        every point is located in the mesh on the GPU (``xg_locate_points``) and the
        data of the face that holds it is selected.  ``obj`` is a ``UgridDataArray`` or an
        array whose LAST axis is the face dimension.  ``out_of_bounds``: "raise", "warn"
        (warning + ``fill_value``), "ignore" (``fill_value``) or "drop" (such points are
        left out; ``PointSelection.index`` tells which were kept).  Only containment is
        offered: ``method="nearest"`` and data on nodes or edges need the reference's
        KD-trees, which are not part of the regridding path.
        """
        import warnings

        method_options = (None, "nearest")
        if method not in method_options:
            raise ValueError(f"method must be one of {method_options}, received: {method}")
        bounds_options = ("warn", "raise", "ignore", "drop")
        if out_of_bounds not in bounds_options:
            raise ValueError(
                f"out_of_bounds must be one of {', '.join(bounds_options)}, received: {out_of_bounds}"
            )
        if method == "nearest":
            raise NotImplementedError("sel_points(method='nearest') is not part of the GPU regridding path")
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        y = np.atleast_1d(np.asarray(y, dtype=np.float64))
        if x.shape != y.shape:
            raise ValueError("shape of x does not match shape of y")
        if x.ndim != 1:
            raise ValueError("x and y must be 1d")
        data = obj.data if isinstance(obj, UgridDataArray) else obj
        if data.shape[-1] != self.n_face:
            raise ValueError(
                f"last axis of obj must be the face dimension ({self.n_face}), received shape {tuple(data.shape)}"
            )
        xy = np.column_stack([x, y])
        index = self.locate_points(xy, tolerance)
        valid = index != -1
        keep = np.arange(x.size)
        if not valid.all():
            msg = "Not all points are located on the topology."
            if out_of_bounds == "raise":
                raise ValueError(msg)
            if out_of_bounds == "warn":
                warnings.warn(msg, UserWarning, stacklevel=2)
            if out_of_bounds == "drop":
                keep = keep[valid]
                index, valid = index[valid], valid[valid]
        is_tensor = isinstance(data, torch.Tensor)
        if is_tensor:
            take = torch.from_numpy(np.where(valid, index, 0)).to(data.device)
            values = data.index_select(data.dim() - 1, take)
            if not valid.all():
                values = values.to(torch.result_type(values, torch.tensor(fill_value)))
                values[..., torch.from_numpy(~valid).to(data.device)] = fill_value
        else:
            values = np.asarray(data)[..., np.where(valid, index, 0)]
            if not valid.all():
                values = values.astype(np.result_type(values.dtype, np.asarray(fill_value).dtype))
                values[..., ~valid] = fill_value
        return PointSelection(values, xy[keep, 0], xy[keep, 1], keep, f"{self.name}_points")

    def rasterize_like(self, x, y):
        # ugrid2d.py:1080-1101
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        yy, xx = np.meshgrid(y, x, indexing="ij")
        nodes = np.column_stack([xx.ravel(), yy.ravel()])
        index = self.celltree.locate_points(nodes).reshape((y.size, x.size))
        return x, y, index

    def rasterize(self, resolution: float, bounds=None):
        # ugrid2d.py:1103-1136
        if bounds is None:
            bounds = self.bounds
        xmin, ymin, xmax, ymax = bounds
        d = abs(resolution)
        xmin = np.floor(xmin / d) * d
        xmax = np.ceil(xmax / d) * d
        ymin = np.floor(ymin / d) * d
        ymax = np.ceil(ymax / d) * d
        x = np.arange(xmin + 0.5 * d, xmax, d)
        y = np.arange(ymax - 0.5 * d, ymin, -d)
        return self.rasterize_like(x, y)

    # ---- constructors -------------------------------------------------------------------
    @staticmethod
    def _from_intervals_helper(node_x, node_y, nx: int, ny: int, name: str) -> "Ugrid2d":
        # ugrid2d.py:1894-1912: counter-clockwise quads over a node lattice
        linear_index = np.arange(node_x.size, dtype=np.int64).reshape((ny + 1, nx + 1))
        face_nodes = np.empty((ny * nx, 4), dtype=np.int64)
        left, right = slice(None, -1), slice(1, None)
        lower, upper = slice(None, -1), slice(1, None)
        if node_x[1] < node_x[0]:
            left, right = right, left
        if node_y[nx + 1] < node_y[0]:
            lower, upper = upper, lower
        face_nodes[:, 0] = linear_index[lower, left].ravel()
        face_nodes[:, 1] = linear_index[lower, right].ravel()
        face_nodes[:, 2] = linear_index[upper, right].ravel()
        face_nodes[:, 3] = linear_index[upper, left].ravel()
        return Ugrid2d(node_x, node_y, -1, face_nodes, name=name)

    @staticmethod
    def from_structured_bounds(x_bounds, y_bounds, name: str = "mesh2d") -> "Ugrid2d":
        # ugrid2d.py:1973-2034 (2D bounds only)
        x_bounds = np.asarray(x_bounds, dtype=np.float64)
        y_bounds = np.asarray(y_bounds, dtype=np.float64)
        if x_bounds.ndim != 2 or y_bounds.ndim != 2:
            raise ValueError(f"Expected 2 dimensions on bounds, received: {x_bounds.ndim}")
        nx, ny = x_bounds.shape[0], y_bounds.shape[0]
        x = bounds1d_to_vertices(x_bounds)
        y = bounds1d_to_vertices(y_bounds)
        node_y, node_x = (a.ravel() for a in np.meshgrid(y, x, indexing="ij"))
        return Ugrid2d._from_intervals_helper(node_x, node_y, nx, ny, name)


class PointSelection:
    """Result of ``Ugrid2d.sel_points``: ``values`` (last axis = the selected points,
    dimension ``<mesh>_points``), their coordinates (the reference attaches them as
    ``<mesh>_x`` / ``<mesh>_y``) and ``index`` = positions of the kept points in the
    caller's x / y."""

    def __init__(self, values, x, y, index, dim: str):
        self.values, self.x, self.y, self.index, self.dim = values, x, y, index, dim

    def __array__(self, dtype=None, copy=None):
        v = self.values.cpu().numpy() if isinstance(self.values, torch.Tensor) else np.asarray(self.values)
        return v.astype(dtype) if dtype is not None else v


class Ugrid1d:
    """The part of xugrid.Ugrid1d the NetworkGridder touches
    (/root/reference/xugrid/ugrid/ugrid1d.py: node coordinates + edge_node_connectivity,
    ``edge_node_coordinates``, ``edge_length``, ``edge_dimension``)."""

    def __init__(self, node_x, node_y, fill_value: int, edge_node_connectivity, name: str = "network1d"):
        self.node_x = np.ascontiguousarray(_host(node_x), dtype=np.float64)
        self.node_y = np.ascontiguousarray(_host(node_y), dtype=np.float64)
        if self.node_x.ndim != 1 or self.node_x.shape != self.node_y.shape:
            raise ValueError("node_x and node_y must be 1D arrays of equal length")
        conn = np.array(_host(edge_node_connectivity), copy=True)
        if conn.ndim != 2 or conn.shape[1] != 2:
            raise ValueError("edge_node_connectivity must be (n_edge, 2)")
        self.edge_node_connectivity = conn.astype(np.int64)
        self.fill_value = fill_value
        self.name = name

    @property
    def n_node(self) -> int:
        return self.node_x.size

    @property
    def n_edge(self) -> int:
        return self.edge_node_connectivity.shape[0]

    @property
    def edge_dimension(self) -> str:
        return f"{self.name}_nEdges"

    @property
    def core_dimension(self) -> str:
        return self.edge_dimension

    @property
    def node_coordinates(self) -> np.ndarray:
        return np.column_stack([self.node_x, self.node_y])

    @property
    def edge_node_coordinates(self) -> np.ndarray:
        """(n_edge, 2, 2): start and end point of every edge"""
        return self.node_coordinates[self.edge_node_connectivity]

    @property
    def edge_length(self) -> np.ndarray:
        xy = self.edge_node_coordinates
        return np.linalg.norm(xy[:, 1] - xy[:, 0], axis=1)


_COPY_STREAMS = {}


def _copy_stream(dev: torch.device) -> "torch.cuda.Stream":
    key = (dev.type, dev.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


def _host(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def bounds1d_to_vertices(bounds: np.ndarray) -> np.ndarray:
    # /root/reference/xugrid/conversion.py:272-282
    diff = np.diff(bounds, axis=0)
    ascending = (diff >= 0.0).all()
    descending = (diff <= 0.0).all()
    if ascending:
        return np.concatenate((bounds[:, 0], bounds[-1:, 1]))
    elif descending:
        return np.concatenate((bounds[:, 1], bounds[-1:, 0]))
    raise ValueError("Bounds are not monotonic ascending or monotonic descending")


class StructuredAxis:
    """One raster axis: midpoints -> cell bounds (structured.py:24-100)."""

    def __init__(self, midpoints, name: str, bounds=None, size=None):
        index = np.asarray(midpoints, dtype=np.float64)
        if index.ndim != 1 or index.size < 1:
            raise ValueError(f"{name} must be a 1D coordinate")
        d = np.diff(index)
        if index.size > 1 and (d < 0).all():
            mid = index[::-1]
            self.flipped = True
        elif index.size == 1 or (d > 0).all():
            mid = index
            self.flipped = False
        else:
            raise ValueError(f"{name} is not monotonic")
        if bounds is not None:
            b = np.asarray(bounds, dtype=np.float64)
        else:
            if size is None:
                if mid.size < 2:
                    raise ValueError(f"cannot infer the cell size of {name} from one midpoint")
                size = np.diff(mid)
                atolx = 1.0e-4 * size[0]
                if not np.allclose(size, size[0], atolx):
                    raise ValueError(
                        f"DataArray has to be equidistant along {name}, or "
                        f'explicit bounds must be given as "{name}bounds", or '
                        f'cellsizes must be as "d{name}"'
                    )
                size = np.full_like(mid, size[0])
            abs_size = np.abs(np.broadcast_to(np.asarray(size, dtype=np.float64), mid.shape))
            b = np.column_stack((mid - 0.5 * abs_size, mid + 0.5 * abs_size))
        self.name = name
        self.index = index
        self.midpoints = mid
        self.bounds = b

    @property
    def size(self) -> int:
        return len(self.bounds)

    @property
    def length(self) -> np.ndarray:
        # structured.py:106-108
        return np.abs(self.bounds[:, 1] - self.bounds[:, 0])

    @property
    def directional_bounds(self) -> np.ndarray:
        return self.bounds[::-1, :].copy() if self.flipped else self.bounds


class StructuredGrid2d:
    """A raster given by its x and y cell midpoints (or explicit bounds)."""

    def __init__(self, x, y, xbounds=None, ybounds=None, name_x: str = "x", name_y: str = "y"):
        self.xbounds = StructuredAxis(x, name_x, xbounds)
        self.ybounds = StructuredAxis(y, name_y, ybounds)

    @property
    def ndim(self) -> int:
        return 2

    @property
    def dims(self) -> Tuple[str, str]:
        return (self.ybounds.name, self.xbounds.name)

    @property
    def shape(self) -> Tuple[int, int]:
        return (self.ybounds.size, self.xbounds.size)

    @property
    def size(self) -> int:
        return self.ybounds.size * self.xbounds.size

    @property
    def coords(self) -> dict:
        return {self.ybounds.name: self.ybounds.index, self.xbounds.name: self.xbounds.index}

    @property
    def area(self) -> np.ndarray:
        # structured.py:485-487
        return np.multiply.outer(self.ybounds.length, self.xbounds.length)

    def to_ugrid2d(self) -> Ugrid2d:
        # structured.py:489-501
        return Ugrid2d.from_structured_bounds(
            self.xbounds.directional_bounds, self.ybounds.directional_bounds
        )


class UgridDataArray:
    """Face data on a mesh (or edge data on a Ugrid1d network): the small part of
    xugrid.UgridDataArray the regridders need (values whose LAST axis is the core
    dimension of the grid)."""

    def __init__(self, data, grid, dims: Optional[Tuple[str, ...]] = None):
        self.data = data
        self.grid = grid
        nd = data.ndim if hasattr(data, "ndim") else np.ndim(data)
        n_core = grid.n_edge if isinstance(grid, Ugrid1d) else grid.n_face
        if data.shape[-1] != n_core:
            raise ValueError("last axis of data must be the face dimension")
        self.dims = tuple(dims) if dims is not None else tuple(
            [f"dim_{i}" for i in range(nd - 1)] + [grid.core_dimension]
        )

    @property
    def values(self):
        d = self.data
        return d.cpu().numpy() if isinstance(d, torch.Tensor) else np.asarray(d)

    @property
    def shape(self):
        return tuple(self.data.shape)
