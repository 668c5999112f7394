"""
Weight containers with the fields of xugrid's ``MatrixCOO`` / ``MatrixCSR``
(/root/reference/xugrid/core/sparse.py:22-137) whose arrays live in HBM as
torch tensors: ``data`` float64, ``indices``/``row``/``col`` int32 (the
reference uses intp; exports widen to int64), ``indptr`` int64.
"""

from __future__ import annotations

from typing import NamedTuple, Union

import numpy as np
import torch

from . import _native
from .celltree import to_device


class MatrixCOO(NamedTuple):
    data: torch.Tensor
    row: torch.Tensor
    col: torch.Tensor
    n: int
    m: int
    nnz: int

    @staticmethod
    def from_triplet(row, col, data, n=None, m=None, device=None) -> "MatrixCOO":
        # sparse.py:52-59
        dev = _dev(device, data)
        row = to_device(row, torch.int32, dev)
        col = to_device(col, torch.int32, dev)
        data = to_device(data, torch.float64, dev)
        if row.shape != col.shape or row.shape != data.shape or row.dim() != 1:
            raise ValueError("row, col and data must be 1D arrays of equal length")
        if n is None:
            n = int(row.max().item()) + 1 if row.numel() else 0
        if m is None:
            m = int(col.max().item()) + 1 if col.numel() else 0
        _check_index_range(row, int(n), "row")
        _check_index_range(col, int(m), "col")
        return MatrixCOO(data, row, col, int(n), int(m), int(data.numel()))

    def to_csr(self) -> "MatrixCSR":
        """sparse.py:61-78: bincount + cumsum; rows must already be sorted."""
        counts = torch.bincount(self.row.to(torch.int64), minlength=self.n)
        indptr = torch.zeros(self.n + 1, dtype=torch.int64, device=self.data.device)
        torch.cumsum(counts, 0, out=indptr[1:])
        return MatrixCSR(self.data, self.col, indptr, self.n, self.m, self.nnz)

    def to_numpy(self):
        return (
            self.data.cpu().numpy(),
            self.row.to(torch.int64).cpu().numpy(),
            self.col.to(torch.int64).cpu().numpy(),
        )


class MatrixCSR(NamedTuple):
    data: torch.Tensor
    indices: torch.Tensor
    indptr: torch.Tensor
    n: int
    m: int
    nnz: int

    @staticmethod
    def from_triplet(row, col, data, n=None, m=None, device=None) -> "MatrixCSR":
        # sparse.py:119-127
        return MatrixCOO.from_triplet(row, col, data, n, m, device).to_csr()

    @staticmethod
    def from_arrays(data, indices, indptr, n: int, m: int, device=None) -> "MatrixCSR":
        dev = _dev(device, data)
        data = to_device(data, torch.float64, dev)
        indices = to_device(indices, torch.int32, dev)
        indptr = to_device(indptr, torch.int64, dev)
        # a stale or mismatched weights file must raise here, as the reference's numpy
        # indexing would (IndexError), instead of sending the kernels out of bounds
        if indptr.dim() != 1 or indptr.numel() != int(n) + 1:
            raise ValueError(f"indptr must have n + 1 = {int(n) + 1} entries, received {tuple(indptr.shape)}")
        if indices.shape != data.shape or data.dim() != 1:
            raise ValueError("indices and data must be 1D arrays of equal length")
        first, last = (int(v) for v in indptr[[0, -1]].tolist())
        if first != 0 or last != data.numel():
            raise ValueError(f"indptr must run from 0 to nnz = {data.numel()}, received {first} .. {last}")
        if indptr.numel() > 1 and bool((indptr[1:] < indptr[:-1]).any().item()):
            raise ValueError("indptr must be non-decreasing")
        _check_index_range(indices, int(m), "indices")
        return MatrixCSR(data, indices, indptr, int(n), int(m), int(data.numel()))

    def to_coo(self) -> MatrixCOO:
        # sparse.py:129-137
        counts = self.indptr[1:] - self.indptr[:-1]
        row = torch.repeat_interleave(
            torch.arange(self.n, device=self.data.device, dtype=torch.int32), counts
        )
        return MatrixCOO(self.data, row, self.indices, self.n, self.m, self.nnz)

    def to_numpy(self):
        return (
            self.data.cpu().numpy(),
            self.indices.to(torch.int64).cpu().numpy(),
            self.indptr.cpu().numpy(),
        )


class PendingCSR:
    """
    CSR weights whose overlap jobs were enqueued WITHOUT waiting for them
    (``CellTree2d.intersect_faces_into``): ``data`` / ``indices`` are sized by an upper
    bound, the number of entries is only known on the device (``indptr[-1]``).  The apply
    kernels need nothing else, so a regridder can queue ``regrid`` behind its own
    construction while the host runs ahead.  ``finalize()`` waits, and returns the exact
    ``MatrixCSR`` (views of the same arrays); if a job turned out incomplete -- candidate
    arena or arrays too small -- the weights are rebuilt through the waiting path
    (``rebuild``), which re-sizes.
    """

    def __init__(self, data, indices, indptr, n: int, m: int, jobs, rebuild):
        self.data, self.indices, self.indptr = data, indices, indptr
        self.n, self.m = int(n), int(m)
        self._jobs, self._rebuild = list(jobs), rebuild
        self._all_jobs = list(jobs)  # (kept for profiling: stats, phase times, graph flag)
        self._final = None
        self._complete = None
        # (row begin, row end, event recorded behind the job that writes those rows): what lets a
        # consumer work on a range as soon as it is done, while later ranges are still being built
        self.ranges = []

    def for_apply(self) -> "MatrixCSR":
        """What csr_apply needs (row pointer + arrays); nnz is not known yet (-1)."""
        return MatrixCSR(self.data, self.indices, self.indptr, self.n, self.m, -1)

    def complete(self) -> bool:
        """Waits for the jobs; False when one of them has to be redone."""
        if self._complete is None:
            ok, end = True, 0
            self.job_phase_ms = []  # device time per phase of every job (profiling)
            for job in self._jobs:
                _, end, good = job.result()
                ok = ok and good
                self.job_phase_ms.append(job.phase_ms)
            self._complete, self._nnz = ok, int(end)
            self._jobs = []
        return self._complete

    def finalize(self) -> "MatrixCSR":
        if self._final is None:
            if self.complete():
                nnz = self._nnz
                self._final = MatrixCSR(self.data[:nnz], self.indices[:nnz], self.indptr, self.n, self.m, nnz)
            else:
                self._final = self._rebuild()
            self.data = self.indices = self.indptr = None
        return self._final


def _check_index_range(index: torch.Tensor, size: int, name: str) -> None:
    """0 <= index < size (one min/max pass on the device)."""
    if index.numel() == 0:
        return
    lo, hi = (int(v) for v in torch.aminmax(index))
    if lo < 0 or hi >= size:
        raise IndexError(f"{name} out of range: values {lo} .. {hi} for size {size}")


def _check_source(source: torch.Tensor, A, what: str) -> None:
    """The library reads ``source`` as (n_extra, m) float64 on the device of the
    weights: anything else would be reinterpreted bytes or an out-of-bounds gather."""
    if source.dim() != 2:
        raise ValueError("source must be (n_extra, m)")
    if source.dtype != torch.float64:
        raise _native.XgError(f"{what}: source must be float64, received {source.dtype}")
    if source.device != A.data.device:
        raise _native.XgError(f"{what}: source is on {source.device}, the weights on {A.data.device}")
    if int(source.shape[1]) != A.m:
        raise ValueError(f"source has {int(source.shape[1])} columns, weights expect {A.m}")


def _dev(device, like) -> torch.device:
    if device is not None:
        return torch.device(device)
    if isinstance(like, torch.Tensor) and like.is_cuda:
        return like.device
    return torch.device("cuda", torch.cuda.current_device())


METHOD_CODES = {
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
    METHOD_CODES[f"p{_p}"] = (7, float(_p))


def csr_apply(
    code: int,
    percentile: float,
    source: torch.Tensor,
    A: MatrixCSR,
    size: int,
    out: Union[torch.Tensor, None] = None,
    row_begin: int = 0,
    row_end: Union[int, None] = None,
) -> torch.Tensor:
    """
    ``make_regrid(func)._regrid(source, A, size)`` (regridder.py:41-67) on the
    device.  ``source`` is (n_extra, m) float64; returns (n_extra, rows) float64
    with NaN for empty rows.  ``A.indptr`` covers rows [row_begin, row_end) of
    the target (a shard built with t_begin/t_end) when they are given.
    """
    L = _native.lib()
    _check_source(source, A, "csr_apply")
    n_extra, m = int(source.shape[0]), int(source.shape[1])
    rows = A.n if row_end is None else row_end - row_begin
    if rows != A.indptr.numel() - 1:
        raise ValueError(f"indptr describes {A.indptr.numel() - 1} rows, {rows} requested")
    if out is None:
        out = torch.empty((n_extra, rows), dtype=torch.float64, device=source.device)
    if (not out.is_cuda or out.dtype != torch.float64 or out.device != source.device
            or (rows > 1 and out.stride(1) != 1) or out.shape[0] < n_extra or out.shape[1] < rows):
        raise _native.XgError("out must be a float64 CUDA tensor of at least (n_extra, rows) on the "
                              "device of the source, with unit stride along the rows")
    with torch.cuda.device(source.device):
        _native.check(
            L.xg_csr_apply(
                code, float(percentile), _native.ptr(A.indptr, torch.int64), _native.ptr(A.indices, torch.int32),
                _native.ptr(A.data, torch.float64), _native.ptr(source.contiguous(), torch.float64), n_extra, m,
                rows, out.data_ptr(),
                int(out.stride(0)) if n_extra > 1 else max(rows, int(out.stride(0))),
                _native.stream_ptr(),
            )
        )
    return out


def csr_apply_peers(code: int, percentile: float, source: torch.Tensor, A: MatrixCSR, out_ptrs, out_ld: int,
                    column_begin: int) -> None:
    """``csr_apply`` whose results go to SEVERAL (n_extra, out_ld) arrays at once -- the copies
    of the output on all ranks of a sharded job, given as raw device pointers of peer-mapped
    memory (``distributed.GatheredOutput``): row r of ``A`` lands in column
    ``column_begin + r`` of every array (``xg_csr_apply_peers``)."""
    import ctypes

    L = _native.lib()
    _check_source(source, A, "csr_apply_peers")
    n_extra, m = int(source.shape[0]), int(source.shape[1])
    rows = int(A.indptr.numel()) - 1
    if column_begin < 0 or column_begin + rows > out_ld:
        raise ValueError("the rows do not fit the output arrays")
    ptrs = (ctypes.c_void_p * len(out_ptrs))(*[int(p) + 8 * int(column_begin) for p in out_ptrs])
    with torch.cuda.device(source.device):
        _native.check(
            L.xg_csr_apply_peers(
                code, float(percentile), _native.ptr(A.indptr, torch.int64), _native.ptr(A.indices, torch.int32),
                _native.ptr(A.data, torch.float64), _native.ptr(source.contiguous(), torch.float64), n_extra, m,
                rows, ptrs, len(out_ptrs), int(out_ld), _native.stream_ptr(),
            )
        )


def coo_gather(source: torch.Tensor, A: MatrixCOO, size: int) -> torch.Tensor:
    """``CentroidLocatorRegridder._regrid`` (regridder.py:400-409) on the device."""
    L = _native.lib()
    _check_source(source, A, "coo_gather")
    if size < A.n:
        raise ValueError(f"size {size} is smaller than the {A.n} rows of the weights")
    n_extra, m = int(source.shape[0]), int(source.shape[1])
    out = torch.empty((n_extra, size), dtype=torch.float64, device=source.device)
    with torch.cuda.device(source.device):
        _native.check(
            L.xg_coo_gather(
                _native.ptr(A.row, torch.int32), _native.ptr(A.col, torch.int32), A.nnz,
                _native.ptr(source.contiguous(), torch.float64), n_extra, m,
                _native.ptr(out), size, _native.stream_ptr(),
            )
        )
    return out
