"""
The regridder classes of xugrid.regrid with the same constructor / ``regrid`` /
``weights`` / ``from_weights`` / ``from_dataset`` surface
(/root/reference/xugrid/regrid/regridder.py:99-659), computing on a B200
through libxgregrid.so.

Differences that follow from running on a GPU without xarray:

* data goes in as ``UgridDataArray`` (xugrid_b200.grid), numpy array or torch
  tensor whose trailing axes are the source grid's; the same kind comes back.
* ``method`` is a name from reduce.py or the object returned by
  ``OverlapRegridder.create_percentile_method(p)``; arbitrary Python callables
  (numba-compiled by the reference, regridder.py:135-136) raise
  ``NotImplementedError``.
* ``weights`` / ``to_dataset()`` return a plain ``dict`` of numpy arrays with
  the reference's variable names (``__regrid_data`` ...,
  regridder.py:264-271); ``from_weights`` / ``from_dataset`` read it back.
* a ``target_range=(begin, end)`` keyword builds and applies only a contiguous
  range of target faces (one shard per GPU; see xugrid_b200.distributed).
"""

from __future__ import annotations

import abc
from typing import Callable, Optional, Tuple, Union

import numpy as np
import torch

from . import _native, sparse
from .grid import StructuredGrid2d, Ugrid1d, Ugrid2d, UgridDataArray
from .sparse import MatrixCOO, MatrixCSR


class PercentileMethod:
    """Returned by create_percentile_method (reduce.py:241-248)."""

    def __init__(self, p: float):
        if not (0.0 <= p <= 100.0):
            raise ValueError(f"percentile must be in the range [0, 100], received: {p}")
        self.percentile = float(p)

    def __call__(self, values, weights, workspace=None):
        """Host evaluation of reduce.percentile (reduce.py:161-203) for one row;
        the regridders evaluate the same rule on the GPU."""
        values = np.asarray(values, dtype=np.float64)
        weights = np.asarray(weights, dtype=np.float64)
        if weights.size == 0 or weights.max() == 0.0:
            return np.nan
        v = np.sort(values[~np.isnan(values)])
        if v.size == 0:
            return np.nan
        if self.percentile == 0 or v.size == 1:
            return v[0]
        if self.percentile == 100:
            return v[-1]
        rank = 1 + (v.size - 1) * self.percentile / 100.0
        f = int(np.floor(rank))
        m = rank - f
        return v[f - 1] * (1 - m) + v[f] * m


def create_percentile_method(p: float) -> PercentileMethod:
    return PercentileMethod(p)


ABSOLUTE_OVERLAP_METHODS = (
    "mean", "harmonic_mean", "geometric_mean", "sum", "minimum", "maximum", "mode", "median",
    "max_overlap", "p5", "p10", "p25", "p50", "p75", "p90", "p95",
)
RELATIVE_OVERLAP_METHODS = ("conductance", "first_order_conservative")


class UnstructuredGrid2d:
    """Topology adapter (/root/reference/xugrid/regrid/unstructured.py:60-220)."""

    def __init__(self, obj):
        if isinstance(obj, UgridDataArray):
            self.ugrid_topology = obj.grid
        elif isinstance(obj, Ugrid2d):
            self.ugrid_topology = obj
        else:
            options = {"Ugrid2d", "UgridDataArray", "UgridDataset"}
            raise TypeError(f"Expected one of {options}, received: {type(obj).__name__}")

    ndim = 1

    @property
    def dims(self):
        return (self.ugrid_topology.face_dimension,)

    @property
    def shape(self):
        return (self.ugrid_topology.n_face,)

    @property
    def size(self):
        return self.ugrid_topology.n_face

    @property
    def area(self):
        return self.ugrid_topology.area

    def to_unstructured(self) -> "UnstructuredGrid2d":
        return self

    # candidate pairs per target row the deferred path sizes its arrays for (what the
    # library's own pair arena holds by default: a job that needs more is redone)
    _ENTRIES_PER_ROW = 16
    # jobs of a regridder go out as one CUDA graph each (celltree.intersect_faces_into)
    _JOB_AS_GRAPH = True
    # a resident range of at least this many target rows is built as two chained jobs, so that
    # regrid() can bring the first part down under the second job.  OFF by default: measured at
    # N = 2 (5 M rows per rank) it gains nothing -- 9.59 vs 9.56 ms end to end: the second job
    # runs 0.35 ms slower while the copy engine moves the first part (launch latency under
    # PCIe traffic) and a job has 0.2 ms of fixed cost
    _SPLIT_MIN_ROWS = 1 << 62
    _SPLIT_AT = 0.8

    def overlap(self, other: "UnstructuredGrid2d", relative: bool,
                target_range: Optional[Tuple[int, int]] = None,
                target_ids: Optional[torch.Tensor] = None, deferred: bool = False):
        """unstructured.py:109-135 + sparse.py:119-127, as one device CSR.
        ``deferred=True``: nothing is waited for (``sparse.PendingCSR``): the arrays are
        sized by an upper bound and the jobs of a target mesh that is still arriving in
        face ranges are chained on the device."""
        if deferred:
            rows = (other.size if target_range is None else target_range[1] - target_range[0]) \
                if target_ids is None else int(target_ids.numel())
            if rows > 0:
                pending = self._overlap_deferred(other, relative, target_range, target_ids, rows)
                if pending is not None:
                    return pending
        if target_ids is not None:
            # a share of the targets given by face numbers: index only the source faces
            # they can touch, rows in the order of the list
            from .celltree import CellTree2d

            xy, faces = other.ugrid_topology.device_arrays()
            sxy, sfaces = self.ugrid_topology.device_arrays()
            ids = target_ids.to(xy.device)
            tree = self.__dict__.pop("_tree_for_waiting_path", None)  # (built by the deferred attempt)
            if tree is None:
                tree = CellTree2d(sxy, sfaces, -1, device=sxy.device, clip_target_ids=(xy, faces, ids))
            self.last_tree = tree
            indptr, indices, data = tree.intersect_faces_csr(xy, faces, -1, relative=relative, ids=ids)
            return MatrixCSR(data, indices, indptr, int(ids.numel()), self.size, int(data.numel()))
        # the source index first: its build must not wait for target arrays that are
        # queued behind the source on the copy stream
        tree = self.ugrid_topology.celltree if target_range is None else None
        chunks = other.ugrid_topology.device_chunks() if target_range is None else None
        if chunks is not None:
            # the target connectivity is still arriving in face ranges: overlap the
            # weights of range c with the transfer of range c + 1
            xy, faces, ranges = chunks
            self.last_tree = tree
            parts, offset = [], 0
            for c, (b, e) in enumerate(ranges):
                other.ugrid_topology.wait_chunk(c)
                ip, ix, dt = tree.intersect_faces_csr(xy, faces, -1, t_begin=b, t_end=e, relative=relative)
                parts.append((ip[1:] + offset if c else ip + offset, ix, dt))
                offset += int(dt.numel())
            indptr = torch.cat([p[0] for p in parts])
            indices = torch.cat([p[1] for p in parts])
            data = torch.cat([p[2] for p in parts])
            return MatrixCSR(data, indices, indptr, other.size, self.size, int(data.numel()))
        t_begin, t_end = (0, other.size) if target_range is None else target_range
        xy, faces = other.ugrid_topology.device_arrays(face_range=(t_begin, t_end))
        if target_range is not None and (t_end - t_begin) < other.size and t_end > t_begin:
            # a shard: index only the source faces its targets can touch
            from .celltree import CellTree2d

            sxy, sfaces = self.ugrid_topology.device_arrays()
            # (the clip box = bounds of the target range is reduced on the device)
            tree = self.__dict__.pop("_tree_for_waiting_path", None)  # (built by the deferred attempt)
            if tree is None:
                tree = CellTree2d(sxy, sfaces, -1, device=sxy.device, clip_targets=(xy, faces, t_begin, t_end))
        elif tree is None:
            tree = self.ugrid_topology.celltree
        self.last_tree = tree
        indptr, indices, data = tree.intersect_faces_csr(
            xy, faces, -1, t_begin=t_begin, t_end=t_end, relative=relative
        )
        return MatrixCSR(data, indices, indptr, t_end - t_begin, self.size, int(data.numel()))

    def _overlap_deferred(self, other, relative, target_range, target_ids, rows: int):
        from .celltree import CellTree2d

        tgt = other.ugrid_topology
        tree = self.ugrid_topology.celltree if (target_range is None and target_ids is None) else None
        chunks = tgt.device_chunks() if (target_range is None and target_ids is None) else None
        if chunks is not None:
            xy, faces, ranges = chunks
        elif target_ids is not None:
            xy, faces = tgt.device_arrays()
            ranges = None
        else:
            t_begin, t_end = (0, other.size) if target_range is None else target_range
            xy, faces = tgt.device_arrays(face_range=(t_begin, t_end))
            ranges = [(t_begin, t_end)]
            if rows >= self._SPLIT_MIN_ROWS:
                # two chained jobs: regrid() applies and brings down the first part under the
                # second job (BaseRegridder._drain_ranges), so only a small tail follows the build
                cut = t_begin + int(rows * self._SPLIT_AT)
                ranges = [(t_begin, cut), (cut, t_end)]
        dev = xy.device
        if tree is None:
            sxy, sfaces = self.ugrid_topology.device_arrays()
            if target_ids is not None:
                ids = target_ids.to(dev)
                tree = CellTree2d(sxy, sfaces, -1, device=dev, clip_target_ids=(xy, faces, ids))
            elif rows < other.size:
                tree = CellTree2d(sxy, sfaces, -1, device=dev, clip_targets=(xy, faces, ranges[0][0], ranges[-1][1]))
            else:
                tree = self.ugrid_topology.celltree
        self.last_tree = tree
        # (source faces indexed for these targets; known on the host: a culled build waits for it)
        n_src = tree.n_indexed if tree.n_face > 2 * rows else tree.n_face
        n_ref = rows if (n_src < tree.n_face or target_ids is not None) else max(rows, other.size)
        if n_src > 2 * n_ref:
            # a source much finer than the targets: its rows outgrow any fixed bound, and the
            # library samples the candidate density first (waiting path)
            if tree is not self.ugrid_topology._celltree:
                self._tree_for_waiting_path = tree  # (a culled index: the waiting path takes it over)
            return None
        cap = max(self._ENTRIES_PER_ROW * rows, 65536)
        indptr = torch.empty(rows + 1, dtype=torch.int64, device=dev)
        indices = torch.empty(cap, dtype=torch.int32, device=dev)
        data = torch.empty(cap, dtype=torch.float64, device=dev)
        jobs, range_events = [], []
        if target_ids is not None:
            jobs.append(tree.intersect_faces_into(xy, faces, indptr, indices, data, relative=relative, ids=ids,
                                                  graph=self._JOB_AS_GRAPH))
        else:
            first = ranges[0][0]
            for c, (b, e) in enumerate(ranges):
                if chunks is not None:
                    tgt.wait_chunk(c)  # (the compute stream waits, not the host)
                jobs.append(tree.intersect_faces_into(xy, faces, indptr[b - first: e - first + 1], indices, data,
                                                      t_begin=b, t_end=e, relative=relative, chained=c > 0,
                                                      graph=self._JOB_AS_GRAPH))
                done = torch.cuda.Event()
                done.record()
                range_events.append((b - first, e - first, done))
        rebuild = lambda: self.overlap(other, relative, target_range, target_ids, deferred=False)
        pending = sparse.PendingCSR(data, indices, indptr, rows, self.size, jobs, rebuild)
        pending.ranges = range_events
        return pending

    def locate_centroids(self, other: "UnstructuredGrid2d", tolerance: Optional[float] = None,
                         target_range: Optional[Tuple[int, int]] = None):
        """unstructured.py:137-144: (source_index, target_index, ones) device tensors;
        with ``target_range`` only for those target faces (target_index counts from its start)."""
        tree = self.ugrid_topology.celltree
        points = _target_points(other, tree.device)
        if isinstance(target_range, torch.Tensor):
            points = points[target_range.long()].contiguous()
        elif target_range is not None:
            points = points[target_range[0]:target_range[1]].contiguous()
        idx = tree.locate_points_device(points, tolerance)
        inside = idx != -1
        source_index = idx[inside]
        target_index = torch.arange(points.shape[0], dtype=torch.int32, device=idx.device)[inside]
        weights = torch.ones(source_index.numel(), dtype=torch.float64, device=idx.device)
        return source_index, target_index, weights

    def barycentric(self, other: "UnstructuredGrid2d", tolerance: Optional[float] = None,
                    target_range: Optional[Tuple[int, int]] = None) -> MatrixCSR:
        """
        unstructured.py:146-201 on the device: Voronoi polygons around the
        source nodes, barycentric weights of the target centroids inside them,
        corner weights redistributed, points outside the source mesh dropped.
        Returns the CSR (rows = target faces, columns = source faces).
        """
        from . import voronoi
        from .celltree import CellTree2d

        grid = self.ugrid_topology
        xy, faces = grid.device_arrays()
        points = _target_points(other, xy.device)
        if isinstance(target_range, torch.Tensor):
            points = points[target_range.long()].contiguous()
        elif target_range is not None:
            # a shard: the rows of its own target faces (the Voronoi mesh of the source is
            # replicated, like the source mesh of the overlap regridders)
            points = points[target_range[0]:target_range[1]].contiguous()
        # Voronoi polygons by the library's kernels (csrc/xg_voronoi.cu), barycentric weights
        # inside them, then corner weights / outside mask / CSR rows in two more kernels
        topo = voronoi.voronoi_topology_device(faces, xy, grid.centroids_device)
        vor_tree = CellTree2d(topo.vertices, topo.faces, -1, device=xy.device)
        polygon_index, weights = vor_tree.compute_barycentric_weights_device(points, tolerance)
        # (only "inside the source mesh or not" is asked: unstructured.py:188-190)
        source_face = grid.celltree.locate_points_device(points, any_face=True)
        indptr, indices, data = voronoi.weights_to_csr_device(topo, polygon_index, weights, source_face)
        return MatrixCSR(data, indices, indptr, int(points.shape[0]), self.size, int(data.numel()))

    def intersection_length(self, other: "Network1d", relative: bool):
        """unstructured.py:203-215: (target face index, source edge index, length of the
        edge inside the face) device tensors, ordered by source edge."""
        edges = torch.from_numpy(other.ugrid_topology.edge_node_coordinates)
        ei, fi, xy = self.ugrid_topology.celltree.intersect_edges_device(edges)
        d = xy[:, 1, :] - xy[:, 0, :]
        length = torch.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        if relative:
            length = length / torch.from_numpy(other.length).to(length.device)[ei.long()]
        return fi, ei, length

    def to_dataset(self, name: str) -> dict:
        g = self.ugrid_topology
        return {
            f"{name}_node_x": g.node_x,
            f"{name}_node_y": g.node_y,
            f"{name}_face_node_connectivity": g.face_node_connectivity.astype(np.int64),
            f"{name}_type": "UnstructuredGrid2d",
        }


class _StructuredAdapter:
    """StructuredGrid2d promoted to quads (structured.py:489-501)."""

    def __init__(self, grid: StructuredGrid2d):
        self.grid = grid
        self._promoted: Optional[UnstructuredGrid2d] = None

    ndim = 2

    @property
    def dims(self):
        return self.grid.dims

    @property
    def shape(self):
        return self.grid.shape

    @property
    def size(self):
        return self.grid.size

    def to_unstructured(self) -> UnstructuredGrid2d:
        if self._promoted is None:
            self._promoted = UnstructuredGrid2d(self.grid.to_ugrid2d())
        return self._promoted

    def centroids_device(self, device) -> torch.Tensor:
        """Centroids of the promoted quad mesh without building it on the host: the
        node lattice and the connectivity of Ugrid2d.from_structured_bounds
        (ugrid2d.py:1894-1912,1973-2034) are laid out on the device with a few
        tensor ops and go through the same centroid kernel as any other mesh, so the
        points are bit-identical to those of the host-built mesh (a 4096 x 4096
        raster costs 0.7 s of numpy on the host, a few ms here)."""
        if self._promoted is not None:
            return self._promoted.ugrid_topology.centroids_device
        from .celltree import face_centroid_device
        from .grid import bounds1d_to_vertices

        g = self.grid
        x = torch.from_numpy(bounds1d_to_vertices(g.xbounds.directional_bounds)).to(device)
        y = torch.from_numpy(bounds1d_to_vertices(g.ybounds.directional_bounds)).to(device)
        nx, ny = x.numel() - 1, y.numel() - 1
        xy = torch.stack([x.repeat(ny + 1), y.repeat_interleave(nx + 1)], dim=1).contiguous()
        linear = torch.arange((ny + 1) * (nx + 1), dtype=torch.int32, device=device).reshape(ny + 1, nx + 1)
        left, right = slice(None, -1), slice(1, None)
        lower, upper = slice(None, -1), slice(1, None)
        if float(x[1]) < float(x[0]):
            left, right = right, left
        if float(y[1]) < float(y[0]):
            lower, upper = upper, lower
        faces = torch.stack([linear[lower, left].reshape(-1), linear[lower, right].reshape(-1),
                             linear[upper, right].reshape(-1), linear[upper, left].reshape(-1)], dim=1).contiguous()
        return face_centroid_device(xy, faces)

    def to_dataset(self, name: str) -> dict:
        g = self.grid
        return {
            f"{name}_x": g.xbounds.index,
            f"{name}_y": g.ybounds.index,
            f"{name}_xbounds": g.xbounds.bounds,
            f"{name}_ybounds": g.ybounds.bounds,
            f"{name}_type": "StructuredGrid2d",
        }


_drain_streams = {}


def _drain_stream(dev: torch.device) -> "torch.cuda.Stream":
    """High-priority side stream for apply + download of finished target ranges."""
    key = (dev.type, dev.index)
    if key not in _drain_streams:
        _drain_streams[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _drain_streams[key]


def _target_points(target, device) -> torch.Tensor:
    """(n, 2) device tensor of the target's face centroids (the query points of
    CentroidLocatorRegridder and BarycentricInterpolator)."""
    if isinstance(target, _StructuredAdapter):
        return target.centroids_device(device)
    return target.ugrid_topology.centroids_device


def _both_structured(source, target) -> bool:
    return isinstance(source, _StructuredAdapter) and isinstance(target, _StructuredAdapter)


def _csr_rows(A: MatrixCSR, begin: int, end: int) -> MatrixCSR:
    """rows [begin, end) of a CSR matrix (the shard of a rank that owns a target range)"""
    lo, hi = int(A.indptr[begin].item()), int(A.indptr[end].item())
    return MatrixCSR(A.data[lo:hi].contiguous(), A.indices[lo:hi].contiguous(),
                     (A.indptr[begin:end + 1] - lo).contiguous(), end - begin, A.m, hi - lo)


def setup_grid(obj):
    # regridder.py:72-80
    if isinstance(obj, (Ugrid2d, UgridDataArray)):
        return UnstructuredGrid2d(obj)
    elif isinstance(obj, StructuredGrid2d):
        return _StructuredAdapter(obj)
    elif isinstance(obj, (UnstructuredGrid2d, _StructuredAdapter, Network1d)):
        return obj
    else:
        raise TypeError(
            f"Expected Ugrid2d, UgridDataArray or StructuredGrid2d, received: {type(obj).__name__}"
        )


def _grid_from_dataset(ds: dict, prefix: str):
    kind = ds[f"{prefix}_type"]
    if kind == "Network1d":
        return Network1d(Ugrid1d(ds[f"{prefix}_node_x"], ds[f"{prefix}_node_y"], -1,
                                 ds[f"{prefix}_edge_node_connectivity"]))
    if kind == "UnstructuredGrid2d":
        return Ugrid2d(ds[f"{prefix}_node_x"], ds[f"{prefix}_node_y"], -1,
                       ds[f"{prefix}_face_node_connectivity"])
    return StructuredGrid2d(ds[f"{prefix}_x"], ds[f"{prefix}_y"],
                            xbounds=ds[f"{prefix}_xbounds"], ybounds=ds[f"{prefix}_ybounds"])


class BaseRegridder(abc.ABC):
    _METHODS: Tuple[str, ...] = ()
    # face ranges a large target mesh is uploaded in (cumulative fractions): a small first
    # range lets the first overlap job start early, a small last one leaves little to apply and
    # bring down after the last job (regrid works range by range, _drain_ranges), ranges in
    # between are large to keep the fixed cost of a job (0.3 ms under the upload) low;
    # measured with scripts/e2e_timeline.py (profiles/r2_e2e_timeline_n1.txt)
    _TARGET_CHUNKS = (0.2, 0.85, 1.0)
    _CHUNK_MIN_FACES = 1 << 20  # smaller target meshes go up in one piece
    _TARGET_AS_POINTS = False  # True: the weights need only the target's face centroids

    def __init__(self, source, target, tolerance: Optional[float] = None,
                 target_range: Optional[Tuple[int, int]] = None, sharded_upload: bool = False,
                 target_ids: Optional[torch.Tensor] = None):
        """``target_range=(b, e)``: this process owns only those target faces (one
        process per GPU, see distributed.py).  ``target_ids`` (int32 device tensor of
        target face numbers, e.g. ``distributed.curve_shard``): this process owns those
        faces; row / output column i belongs to face ``target_ids[i]``.
        ``sharded_upload=True`` (all ranks of the default process group hold the same
        host meshes): the source mesh and the target coordinates go up 1/world per rank
        and are replicated over NVLink."""
        self._source = setup_grid(source)
        self._target = setup_grid(target)
        self._weights = None
        self._sharded_upload = bool(sharded_upload)
        if target_ids is not None:
            if target_range is not None:
                raise ValueError("give target_range or target_ids, not both")
            if _both_structured(self._source, self._target):
                raise NotImplementedError("target_ids is not supported for raster -> raster regridding")
            target_ids = target_ids.to(torch.int32).contiguous()
            # (a list of n rows: the range bookkeeping of a shard applies unchanged)
            target_range = (0, int(target_ids.numel()))
        self._target_ids = target_ids
        self._target_range = target_range
        if _both_structured(self._source, self._target):
            # raster -> raster: the separable path of the reference (convert_to_match,
            # regridder.py:83-96); no mesh, no spatial index, nothing to upload
            _native.lib()  # the weights are applied on the GPU: no CPU fallback
            self._compute_weights(self._source, self._target, tolerance)
            return
        # start both uploads on the copy stream: the index build over the source
        # overlaps with the transfer of the target
        src_grid = self._source.to_unstructured().ugrid_topology
        if self._TARGET_AS_POINTS and isinstance(self._target, _StructuredAdapter):
            # only the cell centres of a raster target are needed: no quad mesh, no upload
            src_grid.prefetch(sharded=sharded_upload)
            self._compute_weights(self._source, self._target, tolerance)
            return
        tgt_grid = self._target.to_unstructured().ugrid_topology
        if target_range is None or target_ids is not None:
            src_grid.prefetch(sharded=sharded_upload)
            tgt_grid.prefetch(n_chunks=self._TARGET_CHUNKS if (tgt_grid.n_face >= self._CHUNK_MIN_FACES and target_ids is None) else 1,
                              sharded=sharded_upload and target_ids is not None)
        else:
            # a shard: its (small) target range first, so that the clip box of the
            # source index is known while the source mesh is still arriving
            tgt_grid.prefetch(face_range=target_range, sharded=sharded_upload)
            src_grid.prefetch(sharded=sharded_upload)
        self._compute_weights(self._source, self._target, tolerance)

    # ---- method selection (regridder.py:124-141) --------------------------------------
    def _setup_regrid(self, func) -> None:
        if isinstance(func, str):
            if func not in self._METHODS:
                raise ValueError(
                    "Invalid regridding method. Available methods are: {}".format(self._METHODS)
                )
            self._method = sparse.METHOD_CODES[func]
        elif isinstance(func, PercentileMethod):
            self._method = (7, func.percentile)
        elif callable(func):
            raise NotImplementedError(
                "custom Python reductions cannot run on the GPU; use a named method or "
                "create_percentile_method(p)"
            )
        else:
            raise TypeError(f"method must be string or callable, received: {type(func).__name__}")

    @abc.abstractmethod
    def _compute_weights(self, source, target, tolerance=None):
        pass

    # The weights of the overlap regridders may still be in the making on the GPU when the
    # constructor returns (sparse.PendingCSR): regrid() queues behind them; anything that
    # needs the exact matrix on the host side (export, gather, nnz) settles them first.
    @property
    def _weights(self):
        pending = self.__dict__.get("_pending")
        if pending is not None:
            self.__dict__["_settled"] = pending.finalize()
            self.__dict__["_pending"] = None
        return self.__dict__.get("_settled")

    @_weights.setter
    def _weights(self, value) -> None:
        if isinstance(value, sparse.PendingCSR):
            self.__dict__["_pending"], self.__dict__["_settled"] = value, None
        else:
            self.__dict__["_pending"], self.__dict__["_settled"] = None, value

    def _regrid(self, source: torch.Tensor, A, size: int) -> torch.Tensor:
        code, p = self._method
        return sparse.csr_apply(code, p, source, A, size)

    # ---- apply (regridder.py:143-262) ---------------------------------------------------
    def _rows(self) -> int:
        if self._target_range is None:
            return self._target.size
        return self._target_range[1] - self._target_range[0]

    def regrid_device(self, source: torch.Tensor) -> torch.Tensor:
        """(n_extra, source.size) float64 device tensor -> (n_extra, rows)."""
        pending = self.__dict__.get("_pending")
        A = pending.for_apply() if pending is not None else self._weights
        return self._regrid(source, A, self._rows())

    def _upload_values(self, flat: torch.Tensor, dev, wait: bool = True):
        """Pinned host values -> device on the COPY stream (behind whatever mesh upload is
        still in flight there, next to the kernels that build the weights); 1/world per
        rank + an all-gather over NVLink for a regridder built with sharded_upload."""
        from .grid import _copy_stream

        cs = _copy_stream(dev)
        with torch.cuda.stream(cs):
            if self._sharded_upload and flat.shape[0] == 1:
                from .distributed import sharded_upload

                src = sharded_upload(flat.reshape(-1), dev).reshape(1, -1)
            else:
                src = flat.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        if wait:
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            src.record_stream(cur)
        return src, ev

    def _drain_ranges(self, pending, src: torch.Tensor, values_ready, res2d: torch.Tensor) -> None:
        """The weights are being built range by range (a target mesh that is still arriving over
        PCIe): apply and bring down the columns of every range as soon as its job is done, on a
        high-priority side stream -- the device-to-host copy of the early ranges runs under the
        jobs of the later ones, and only the last range's apply + copy is left at the end."""
        dev = src.device
        code, p = self._method
        drain = _drain_stream(dev)
        out_d = torch.empty((int(src.shape[0]), pending.n), dtype=torch.float64, device=dev)
        drain.wait_event(values_ready)
        with torch.cuda.stream(drain):
            for b, e, done in pending.ranges:
                drain.wait_event(done)
                A = MatrixCSR(pending.data, pending.indices, pending.indptr[b: e + 1], e - b, pending.m, -1)
                sparse.csr_apply(code, p, src, A, e - b, out=out_d[:, b:e], row_begin=b, row_end=e)
                res2d[:, b:e].copy_(out_d[:, b:e], non_blocking=True)
        drain.synchronize()

    def _regrid_array(self, source, out=None):
        source_grid = self._source
        nd = source_grid.ndim
        if source.ndim < nd or tuple(source.shape[source.ndim - nd:]) != tuple(source_grid.shape):
            raise ValueError(
                f"data does not contain regridder source dimensions: expected trailing shape "
                f"{tuple(source_grid.shape)}, received {tuple(source.shape)}"
            )
        first_dims_shape = tuple(source.shape[: source.ndim - nd])
        is_tensor = isinstance(source, torch.Tensor)
        pending = self.__dict__.get("_pending")
        host_tensor = is_tensor and not source.is_cuda
        if pending is not None and not host_tensor:
            # device data in: nothing below waits for the GPU, so settle the weights first
            pending = None
        dev = pending.data.device if pending is not None else self._weights.data.device
        if is_tensor:
            flat = source.reshape((-1, source_grid.size))
            if host_tensor and flat.is_pinned() and flat.dtype == torch.float64 and flat.is_contiguous():
                ranged = (pending is not None and len(pending.ranges) > 1 and isinstance(self, BaseOverlapRegridder))
                res = res2d = None
                if ranged:
                    out_shape = first_dims_shape + (tuple(self._target.shape) if self._target_range is None
                                                    else (self._rows(),))
                    if out is not None:
                        if not isinstance(out, torch.Tensor) or out.is_cuda or tuple(out.shape) != tuple(out_shape) \
                                or out.dtype != torch.float64:
                            raise ValueError(f"out must be a float64 host tensor of shape {tuple(out_shape)}")
                        res = out
                    else:
                        res = torch.empty(out_shape, dtype=torch.float64, pin_memory=True)
                    if res.dim() == 2 and res.shape[0] == flat.shape[0]:
                        res2d = res  # (e.g. the columns of a shared host array: strided, fine)
                    elif res.is_contiguous():
                        res2d = res.view(flat.shape[0], -1)
                    else:
                        ranged = False
                src, values_ready = self._upload_values(flat, dev, wait=not ranged)
                if ranged:
                    self._drain_ranges(pending, src, values_ready, res2d)
                    if not pending.complete():
                        _ = self._weights  # (rebuilt through the waiting path; apply again)
                        return self._regrid_array(source, out)
                    return res
            else:
                src = flat.to(dev, non_blocking=True).to(torch.float64).contiguous()
        else:
            a = np.ascontiguousarray(np.asarray(source).reshape((-1, source_grid.size)))
            src = torch.from_numpy(a).to(dev, non_blocking=True).to(torch.float64)
        res_d = self.regrid_device(src)
        if self._target_range is None:
            out_shape = first_dims_shape + tuple(self._target.shape)
        else:
            out_shape = first_dims_shape + (self._rows(),)
        res_d = res_d.reshape(out_shape)
        if host_tensor:
            # host tensor in -> host tensor out (pinned if the input was): async D2H + one sync
            if out is not None:
                if not isinstance(out, torch.Tensor) or out.is_cuda or tuple(out.shape) != tuple(out_shape) \
                        or out.dtype != torch.float64:
                    raise ValueError(f"out must be a float64 host tensor of shape {tuple(out_shape)}")
                res = out
            else:
                res = torch.empty(out_shape, dtype=torch.float64, pin_memory=source.is_pinned())
            res.copy_(res_d, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            if pending is not None and not pending.complete():
                # the upper-bound arrays were too small: the weights are rebuilt (waiting
                # path) and applied again
                _ = self._weights
                return self._regrid_array(source, out)
            return res
        if out is not None:
            raise ValueError("out= is for host tensors (the result of device data stays on the device)")
        return res_d if is_tensor else res_d.cpu().numpy()

    def regrid(self, data, out=None):
        """regridder.py:212-262.  ``out`` (host tensors only): a float64 host tensor of the
        result's shape -- pinned, or a view of ``distributed.shared_host_array`` -- that
        receives the result instead of a fresh allocation."""
        if isinstance(data, UgridDataArray):
            res = self._regrid_array(data.data, out)
            if isinstance(self._target, UnstructuredGrid2d) and self._target_range is None:
                return UgridDataArray(res, self._target.ugrid_topology,
                                      data.dims[:-1] + (self._target.ugrid_topology.face_dimension,))
            return res
        elif isinstance(data, (np.ndarray, torch.Tensor)):
            return self._regrid_array(data, out)
        raise TypeError(
            f"Expected UgridDataArray, numpy.ndarray or torch.Tensor, received: {type(data).__name__}"
        )

    # ---- persistence (regridder.py:264-361) -----------------------------------------------
    def to_dataset(self) -> dict:
        w = self._weights
        if w is None:
            raise ValueError("Weights have not been computed yet.")
        if self._target_range is not None:
            # the rows of one shard next to the topology of the whole target would read back
            # as a mis-shaped regridder: collect the rows of all ranks first
            raise ValueError(
                "this regridder holds only the target rows [{}, {}); call gather_weights() "
                "before exporting the weights".format(*self._target_range)
            )
        ds = {}
        if isinstance(w, MatrixCSR):
            data, indices, indptr = w.to_numpy()
            ds.update({"__regrid_data": data, "__regrid_indices": indices, "__regrid_indptr": indptr})
        else:
            data, row, col = w.to_numpy()
            ds.update({"__regrid_data": data, "__regrid_row": row, "__regrid_col": col})
        ds.update({"__regrid_n": w.n, "__regrid_m": w.m, "__regrid_nnz": w.nnz})
        ds.update(self._source.to_dataset("__source"))
        ds.update(self._target.to_dataset("__target"))
        return ds

    @property
    def weights(self) -> dict:
        return self.to_dataset()

    def weights_as_dataframe(self):
        import pandas as pd

        matrix = self._weights
        if matrix is None:
            raise ValueError("Weights have not been computed yet.")
        if isinstance(matrix, MatrixCSR):
            matrix = matrix.to_coo()
        data, row, col = matrix.to_numpy()
        return pd.DataFrame({"target_index": row, "source_index": col, "weight": data})

    @staticmethod
    def _csr_from_dataset(ds: dict) -> MatrixCSR:
        return MatrixCSR.from_arrays(ds["__regrid_data"], ds["__regrid_indices"], ds["__regrid_indptr"],
                                     int(ds["__regrid_n"]), int(ds["__regrid_m"]))

    @staticmethod
    def _coo_from_dataset(ds: dict) -> MatrixCOO:
        return MatrixCOO.from_triplet(ds["__regrid_row"], ds["__regrid_col"], ds["__regrid_data"],
                                      n=int(ds["__regrid_n"]), m=int(ds["__regrid_m"]))

    @classmethod
    def _weights_from_dataset(cls, ds: dict):
        return cls._csr_from_dataset(ds)

    def to_xarray(self):
        """``to_dataset()`` as an ``xarray.Dataset`` in the reference's layout
        (needs xarray; see xugrid_b200/xarray_io.py)."""
        from . import xarray_io

        return xarray_io.to_xarray(self.to_dataset())

    @classmethod
    def from_weights(cls, weights, target):
        if not isinstance(weights, dict):
            # an xarray.Dataset as written by the reference's to_dataset()
            from . import xarray_io

            weights = xarray_io.from_xarray(weights)
        instance = cls.__new__(cls)
        instance._weights = cls._weights_from_dataset(weights)
        instance._target = setup_grid(target)
        instance._source = setup_grid(_grid_from_dataset(weights, "__source"))
        instance._target_range = None
        instance._target_ids = None
        instance._sharded_upload = False
        return instance

    @classmethod
    def from_dataset(cls, dataset):
        if not isinstance(dataset, dict):
            from . import xarray_io

            dataset = xarray_io.from_xarray(dataset)
        target = _grid_from_dataset(dataset, "__target")
        return cls.from_weights(dataset, target)


class CentroidLocatorRegridder(BaseRegridder):
    """regridder.py:364-426"""

    _TARGET_AS_POINTS = True

    def _compute_weights(self, source, target, tolerance=None):
        if _both_structured(source, target):
            # structured.py:549-574
            from . import structured

            if self._target_range is not None:
                raise NotImplementedError("target_range is not supported for raster -> raster regridding")
            self._weights = structured.locate_centroids(source.grid, target.grid)
            return
        # (a raster target is only needed as points: it is not promoted to a quad mesh)
        source_index, target_index, weight_values = source.to_unstructured().locate_centroids(
            target, tolerance, self._target_ids if self._target_ids is not None else self._target_range)
        self._weights = MatrixCOO(weight_values, target_index.contiguous(), source_index.contiguous(),
                                  self._rows(), source.size, int(weight_values.numel()))

    def _regrid(self, source, A, size):
        return sparse.coo_gather(source, A, size)

    @classmethod
    def _weights_from_dataset(cls, ds: dict):
        return cls._coo_from_dataset(ds)


class BaseOverlapRegridder(BaseRegridder, abc.ABC):
    def gather_weights(self, group=None) -> None:
        """A regridder built with ``target_range`` holds the rows of its own targets only.
        This collects the rows of all ranks (distributed.all_gather_csr): afterwards
        ``.weights`` / ``.regrid`` cover the whole target mesh on every rank."""
        if self._target_range is None:
            return
        from .distributed import all_gather_csr

        if self._target_ids is not None:
            raise NotImplementedError("gather_weights of a regridder built from a face list (target_ids)")
        self._weights = all_gather_csr(self._weights, self._target.size, group)
        self._target_range = None

    def _overlap_weights(self, source, target, relative: bool) -> None:
        # regridder.py:428-436
        if _both_structured(source, target):
            # structured.py:523-547
            from . import structured

            A = structured.overlap(source.grid, target.grid, relative)
            self._weights = A if self._target_range is None else _csr_rows(A, *self._target_range)
            return
        src, tgt = source.to_unstructured(), target.to_unstructured()
        self._weights = src.overlap(tgt, relative=relative, target_range=self._target_range,
                                    target_ids=self._target_ids, deferred=True)


class OverlapRegridder(BaseOverlapRegridder):
    """regridder.py:454-532"""

    _METHODS = ABSOLUTE_OVERLAP_METHODS

    def __init__(self, source, target, method: Union[str, Callable] = "mean",
                 target_range: Optional[Tuple[int, int]] = None, sharded_upload: bool = False,
                 target_ids: Optional[torch.Tensor] = None):
        super().__init__(source=source, target=target, target_range=target_range,
                         sharded_upload=sharded_upload, target_ids=target_ids)
        self._setup_regrid(method)

    def _compute_weights(self, source, target, tolerance=None) -> None:
        self._overlap_weights(source, target, relative=False)

    @staticmethod
    def create_percentile_method(percentile: float) -> PercentileMethod:
        return create_percentile_method(percentile)

    @classmethod
    def from_weights(cls, weights: dict, target, method: Union[str, Callable] = "mean"):
        instance = super().from_weights(weights, target)
        instance._setup_regrid(method)
        return instance


class RelativeOverlapRegridder(BaseOverlapRegridder):
    """regridder.py:535-586"""

    _METHODS = RELATIVE_OVERLAP_METHODS

    def __init__(self, source, target, method: Union[str, Callable] = "first_order_conservative",
                 target_range: Optional[Tuple[int, int]] = None, sharded_upload: bool = False,
                 target_ids: Optional[torch.Tensor] = None):
        super().__init__(source=source, target=target, tolerance=None, target_range=target_range,
                         sharded_upload=sharded_upload, target_ids=target_ids)
        self._setup_regrid(method)

    def _compute_weights(self, source, target, tolerance=None) -> None:
        self._overlap_weights(source, target, relative=True)

    @classmethod
    def from_weights(cls, weights: dict, target,
                     method: Union[str, Callable] = "first_order_conservative"):
        instance = super().from_weights(weights, target)
        instance._setup_regrid(method)
        return instance


class BarycentricInterpolator(BaseRegridder):
    """
    regridder.py:589-659.  Interpolates between the values of the source faces
    (located at their centroids) with barycentric weights inside the Voronoi
    polygons around the source nodes.  Raster -> raster takes the reference's
    separable ``linear_weights`` path (xugrid_b200/structured.py); a raster source
    with a mesh target is promoted to a quad mesh, as in the reference.
    """

    _METHODS = ("mean",)
    _TARGET_AS_POINTS = True

    def __init__(self, source, target, tolerance: Optional[float] = None,
                 target_range: Optional[Tuple[int, int]] = None, target_ids: Optional[torch.Tensor] = None):
        super().__init__(source, target, tolerance, target_range=target_range, target_ids=target_ids)
        # weights of a target face sum to 1: a weighted mean also handles NaNs
        self._setup_regrid("mean")

    def _compute_weights(self, source, target, tolerance=None):
        if _both_structured(source, target):
            # regridder.py:629-631 -> structured.py:576-601 (bilinear between cell centres)
            from . import structured

            A = structured.linear_weights(source.grid, target.grid)
            self._weights = A if self._target_range is None else _csr_rows(A, *self._target_range)
            return
        # (a raster target is only needed as points: it is not promoted to a quad mesh)
        self._weights = source.to_unstructured().barycentric(
            target, tolerance, self._target_ids if self._target_ids is not None else self._target_range)

    @classmethod
    def from_weights(cls, weights: dict, target):
        instance = super().from_weights(weights, target)
        instance._setup_regrid("mean")
        return instance


class Network1d:
    """regrid/network.py:4-38"""

    ndim = 1

    def __init__(self, obj):
        if isinstance(obj, UgridDataArray) and isinstance(obj.grid, Ugrid1d):
            self.ugrid_topology = obj.grid
        elif isinstance(obj, Ugrid1d):
            self.ugrid_topology = obj
        else:
            options = {"Ugrid1d", "UgridDataArray"}
            raise TypeError(f"Expected one of {options}, received: {type(obj).__name__}")

    @property
    def dims(self):
        return (self.ugrid_topology.edge_dimension,)

    @property
    def shape(self):
        return (self.ugrid_topology.n_edge,)

    @property
    def size(self):
        return self.ugrid_topology.n_edge

    @property
    def length(self):
        return self.ugrid_topology.edge_length

    def to_unstructured(self):
        return self

    def to_dataset(self, name: str) -> dict:
        g = self.ugrid_topology
        return {
            f"{name}_node_x": g.node_x,
            f"{name}_node_y": g.node_y,
            f"{name}_edge_node_connectivity": g.edge_node_connectivity,
            f"{name}_type": "Network1d",
        }


class NetworkGridder(BaseRegridder):
    """
    regrid/gridder.py:24-85.  Grids data on the edges of a 1D network onto the faces
    of a 2D mesh: the weight of (face, edge) is the length of the edge inside the face
    (``CellTree2d.intersect_edges``, Cyrus-Beck clipping on the GPU), reduced with any
    of the absolute overlap methods.
    """

    _METHODS = ABSOLUTE_OVERLAP_METHODS

    def __init__(self, source, target, method: Union[str, Callable] = "mean"):
        self._source = Network1d(source)
        self._target = setup_grid(target)
        self._weights = None
        self._target_range = None
        self._target_ids = None
        self._compute_weights(self._source, self._target, relative=False)
        self._setup_regrid(method)

    def _compute_weights(self, source, target, relative: bool = False) -> None:
        tgt = target.to_unstructured()
        target_index, source_index, length = tgt.intersection_length(source, relative=relative)
        # rows = target faces, entries of a row in ascending edge index (sparse.py:119-127)
        order = torch.argsort(target_index.to(torch.int64) * max(source.size, 1) + source_index.to(torch.int64))
        self._weights = MatrixCSR.from_triplet(target_index[order], source_index[order], length[order],
                                               n=target.size, m=source.size)

    @classmethod
    def _weights_from_dataset(cls, ds: dict):
        return cls._csr_from_dataset(ds)

    @classmethod
    def from_weights(cls, weights: dict, target, method: Union[str, Callable] = "mean"):
        instance = super().from_weights(weights, target)
        instance._setup_regrid(method)
        return instance
