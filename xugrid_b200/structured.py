"""
Raster -> raster weights: the separable path the reference takes when BOTH grids
are structured (``convert_to_match``, /root/reference/xugrid/regrid/regridder.py:83-96:
``{StructuredGrid2d} -> StructuredGrid2d``, so the spatial index is never touched;
/root/reference/xugrid/regrid/structured.py:523-601).

Every 2-D weight is the product of one weight along y and one along x, so the
work splits into

* two 1-D problems of O(n) intervals each, solved on the host with binary
  searches over the (ascending) cell bounds -- same tie rules as the reference:
    - overlap (structured.py:370-391, overlap_1d.py:150-248): every pair of
      intervals with ``min(upper) - max(lower) > 0``;
    - centroid location (structured.py:126-157): the cell whose bounds contain the
      target midpoint, outer bounds excluded, a midpoint on an interior bound goes
      to the cell with the smaller index *in the caller's axis order*;
    - linear interpolation (structured.py:242-330,416-438): the located cell and
      its neighbour towards the midpoint, weights ``w`` and ``1 - w``; at the rim
      the neighbour is the cell itself (a 0 / 1 pair, both entries are kept);
* the outer product of the two triplet lists, laid out on the device straight
  into CSR order (rows = target cells in C order) -- the arrays of a
  4096 x 4096 target never exist on the host.

Within a row the entries are ordered by source index (y-major), which is a
canonical order: the reference's own order is whatever ``np.argsort`` (unstable)
leaves, see the note in /root/reference/tests/test_regrid/test_structured.py:41-43.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .grid import StructuredAxis, StructuredGrid2d
from .sparse import MatrixCOO, MatrixCSR

Triplet = Tuple[np.ndarray, np.ndarray, np.ndarray]


def _unflip(axis: StructuredAxis, index: np.ndarray) -> np.ndarray:
    # structured.py:117-121: positions along the ascending bounds -> the caller's axis order
    return axis.size - index - 1 if axis.flipped else index


def _by_target(source: np.ndarray, target: np.ndarray, weights: np.ndarray) -> Triplet:
    order = np.lexsort((source, target))
    return source[order], target[order], weights[order]


def overlap_1d(source: StructuredAxis, target: StructuredAxis, relative: bool) -> Triplet:
    """(source index, target index, overlapping length) of every pair of cells that
    share more than a point; ``relative`` divides by the source cell's length."""
    s_lo, s_hi = source.bounds[:, 0], source.bounds[:, 1]
    t_lo, t_hi = target.bounds[:, 0], target.bounds[:, 1]
    # the sources that can meet target cell j form one run: from the first source whose
    # upper bound lies above the target's lower bound up to the first whose lower bound
    # is not below the target's upper bound
    first = np.searchsorted(s_hi, t_lo, side="right")
    last = np.searchsorted(s_lo, t_hi, side="left")
    count = np.maximum(last - first, 0)
    tgt = np.repeat(np.arange(target.size, dtype=np.int64), count)
    start = np.cumsum(count) - count
    src = np.arange(int(count.sum()), dtype=np.int64) - np.repeat(start, count) + np.repeat(first, count)
    length = np.maximum(0.0, np.minimum(s_hi[src], t_hi[tgt]) - np.maximum(s_lo[src], t_lo[tgt]))
    keep = length > 0.0
    src, tgt, length = src[keep], tgt[keep], length[keep]
    src, tgt = _unflip(source, src), _unflip(target, tgt)
    if relative:
        length = length / source.length[src]
    return _by_target(src, tgt, length)


def _locate(source: StructuredAxis, target: StructuredAxis) -> Tuple[np.ndarray, np.ndarray]:
    """positions (along the ascending bounds) of the source cell holding each target
    midpoint, and of those midpoints; a tie on an interior bound follows
    ``side``: left for an ascending axis, right for a descending one
    (structured.py:36-43), i.e. always the smaller index in the caller's order."""
    side = "right" if source.flipped else "left"
    mid = target.midpoints
    below_lower = np.searchsorted(source.bounds[:, 0], mid, side=side)
    below_upper = np.searchsorted(source.bounds[:, 1], mid, side=side)
    valid = (below_lower == below_upper + 1) & (mid > source.bounds[0, 0]) & (mid < source.bounds[-1, 1])
    return below_upper[valid], np.arange(target.size, dtype=np.int64)[valid]


def locate_1d(source: StructuredAxis, target: StructuredAxis) -> Triplet:
    s_pos, t_pos = _locate(source, target)
    return _by_target(_unflip(source, s_pos), _unflip(target, t_pos), np.ones(s_pos.size))


def linear_1d(source: StructuredAxis, target: StructuredAxis) -> Triplet:
    if source.midpoints.size < 2:
        raise ValueError(
            f"Coordinate {source.name} has size: {source.midpoints.size}. "
            "At least two points are required for interpolation."
        )
    s_pos, t_pos = _locate(source, target)
    s_mid, t_mid = source.midpoints[s_pos], target.midpoints[t_pos]
    # the neighbour on the side of the target midpoint, kept inside the axis
    n_pos = np.clip(s_pos + np.where(t_mid <= s_mid, -1, 1), 0, source.midpoints.size - 1)
    span = source.midpoints[n_pos] - s_mid
    at_rim = n_pos == s_pos
    w = 1.0 - (t_mid - s_mid) / np.where(span == 0, 1.0, span)
    w[at_rim] = 0.0  # the whole weight goes to the (identical) neighbour entry
    src = np.column_stack((_unflip(source, s_pos), _unflip(source, n_pos))).ravel()
    tgt = np.repeat(_unflip(target, t_pos), 2)
    return _by_target(src, tgt, np.column_stack((w, 1.0 - w)).ravel())


def _product_csr(ty: Triplet, tx: Triplet, source: StructuredGrid2d, target: StructuredGrid2d,
                 device) -> MatrixCSR:
    """Outer product of the y and x triplets as a device CSR over the target cells in
    C order; entries of a row ordered by (source y, source x)."""
    (sy, ry, wy), (sx, rx, wx) = ty, tx
    ny_t, nx_t = target.shape
    nx_s = source.shape[1]
    dev = torch.device(device)
    cy = torch.from_numpy(np.bincount(ry, minlength=ny_t)).to(dev)
    cx = torch.from_numpy(np.bincount(rx, minlength=nx_t)).to(dev)
    oy = torch.cumsum(cy, 0) - cy  # first y entry of every target row
    ox = torch.cumsum(cx, 0) - cx
    counts = (cy[:, None] * cx[None, :]).reshape(-1)
    indptr = torch.zeros(ny_t * nx_t + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr[1:])
    nnz = int(indptr[-1].item())
    row = torch.repeat_interleave(torch.arange(ny_t * nx_t, device=dev), counts, output_size=nnz)
    k = torch.arange(nnz, device=dev) - indptr[row]  # position inside the row
    j, i = row // nx_t, row % nx_t
    ey = oy[j] + k // cx[i]
    ex = ox[i] + k % cx[i]
    sy_d, sx_d = torch.from_numpy(sy).to(dev), torch.from_numpy(sx).to(dev)
    wy_d, wx_d = torch.from_numpy(wy).to(dev), torch.from_numpy(wx).to(dev)
    indices = (sy_d[ey] * nx_s + sx_d[ex]).to(torch.int32)
    data = wy_d[ey] * wx_d[ex]
    return MatrixCSR(data, indices, indptr, target.size, source.size, nnz)


def _device(device) -> torch.device:
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def overlap(source: StructuredGrid2d, target: StructuredGrid2d, relative: bool, device=None) -> MatrixCSR:
    """structured.py:523-547 -> MatrixCSR.from_triplet (regridder.py:428-436)."""
    ty = overlap_1d(source.ybounds, target.ybounds, relative)
    tx = overlap_1d(source.xbounds, target.xbounds, relative)
    return _product_csr(ty, tx, source, target, _device(device))


def locate_centroids(source: StructuredGrid2d, target: StructuredGrid2d, device=None) -> MatrixCOO:
    """structured.py:549-574 -> the COO matrix of CentroidLocatorRegridder (regridder.py:386-397)."""
    A = _product_csr(locate_1d(source.ybounds, target.ybounds), locate_1d(source.xbounds, target.xbounds),
                     source, target, _device(device))
    return A.to_coo()


def linear_weights(source: StructuredGrid2d, target: StructuredGrid2d, device=None) -> MatrixCSR:
    """structured.py:576-601: bilinear interpolation between source cell centres."""
    ty = linear_1d(source.ybounds, target.ybounds)
    tx = linear_1d(source.xbounds, target.xbounds)
    return _product_csr(ty, tx, source, target, _device(device))
