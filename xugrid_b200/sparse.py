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
        if n is None:
            n = int(row.max().item()) + 1
        if m is None:
            m = int(col.max().item()) + 1
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
    if source.dim() != 2:
        raise ValueError("source must be (n_extra, m)")
    n_extra, m = int(source.shape[0]), int(source.shape[1])
    if m != A.m:
        raise ValueError(f"source has {m} columns, weights expect {A.m}")
    rows = A.n if row_end is None else row_end - row_begin
    if out is None:
        out = torch.empty((n_extra, rows), dtype=torch.float64, device=source.device)
    if not out.is_cuda or out.dtype != torch.float64 or (rows > 1 and out.stride(1) != 1):
        raise _native.XgError("out must be a float64 CUDA tensor with unit stride along the rows")
    with torch.cuda.device(source.device):
        _native.check(
            L.xg_csr_apply(
                code, float(percentile), _native.ptr(A.indptr), _native.ptr(A.indices),
                _native.ptr(A.data), _native.ptr(source), n_extra, m, rows, out.data_ptr(),
                int(out.stride(0)) if n_extra > 1 else max(rows, int(out.stride(0))),
                _native.stream_ptr(),
            )
        )
    return out


def coo_gather(source: torch.Tensor, A: MatrixCOO, size: int) -> torch.Tensor:
    """``CentroidLocatorRegridder._regrid`` (regridder.py:400-409) on the device."""
    L = _native.lib()
    n_extra, m = int(source.shape[0]), int(source.shape[1])
    out = torch.empty((n_extra, size), dtype=torch.float64, device=source.device)
    with torch.cuda.device(source.device):
        _native.check(
            L.xg_coo_gather(
                _native.ptr(A.row), _native.ptr(A.col), A.nnz, _native.ptr(source), n_extra, m,
                _native.ptr(out), size, _native.stream_ptr(),
            )
        )
    return out
