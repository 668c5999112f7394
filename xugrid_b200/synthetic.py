"""
Seeded synthetic meshes for tests and benchmarks.

* ``delaunay``          random points in the unit square, scipy/qhull triangulation
* ``jittered_lattice``  nx x ny lattice with jittered nodes; every cell stays
                        convex and counter-clockwise (|jitter| <= 0.2 h), cells
                        are kept as quads or split into two triangles along a
                        random diagonal -- fast enough for tens of millions of faces
* ``disk``              the 384-triangle disk of xugrid.data.disk()
                        (/root/reference/xugrid/data/synthetic.py:27-61), with
                        scipy's Delaunay standing in for matplotlib.tri
"""

from __future__ import annotations

from typing import Tuple

import numpy as np


def delaunay(n_points: int, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    from scipy.spatial import Delaunay

    pts = np.random.default_rng(seed).random((n_points, 2))
    tri = Delaunay(pts)
    return pts, np.ascontiguousarray(tri.simplices.astype(np.int64))


def jittered_lattice(nx: int, ny: int, seed: int, kind: str = "tri", jitter: float = 0.2,
                     quad_fraction: float = 0.5, bounds=(0.0, 1.0, 0.0, 1.0),
                     index_dtype=np.int64) -> Tuple[np.ndarray, np.ndarray]:
    """
    kind = "tri"   -> (2 * nx * ny, 3) triangles
    kind = "quad"  -> (nx * ny, 4) quads
    kind = "mixed" -> quads with probability quad_fraction, else two triangles;
                      (n_face, 4) with -1 fill
    """
    rng = np.random.default_rng(seed)
    x0, x1, y0, y1 = bounds
    hx, hy = (x1 - x0) / nx, (y1 - y0) / ny
    gx = np.linspace(x0, x1, nx + 1)
    gy = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(gx, gy)  # (ny+1, nx+1)
    X = X + rng.uniform(-jitter, jitter, X.shape) * hx
    Y = Y + rng.uniform(-jitter, jitter, Y.shape) * hy
    xy = np.column_stack([X.ravel(), Y.ravel()])
    idx = np.arange((nx + 1) * (ny + 1), dtype=index_dtype).reshape(ny + 1, nx + 1)
    a = idx[:-1, :-1].ravel()
    b = idx[:-1, 1:].ravel()
    c = idx[1:, 1:].ravel()
    d = idx[1:, :-1].ravel()
    n_cell = nx * ny
    if kind == "quad":
        return xy, np.column_stack([a, b, c, d])
    diag = rng.random(n_cell) < 0.5
    # diagonal a-c: (a,b,c) + (a,c,d); diagonal b-d: (a,b,d) + (b,c,d)
    t1 = np.where(diag[:, None], np.column_stack([a, b, c]), np.column_stack([a, b, d]))
    t2 = np.where(diag[:, None], np.column_stack([a, c, d]), np.column_stack([b, c, d]))
    if kind == "tri":
        faces = np.empty((2 * n_cell, 3), dtype=index_dtype)
        faces[0::2] = t1
        faces[1::2] = t2
        return xy, faces
    if kind != "mixed":
        raise ValueError("kind must be 'tri', 'quad' or 'mixed'")
    is_quad = rng.random(n_cell) < quad_fraction
    n_face = int(is_quad.sum() + 2 * (~is_quad).sum())
    faces = np.full((n_face, 4), -1, dtype=index_dtype)
    start = np.cumsum(np.where(is_quad, 1, 2)) - np.where(is_quad, 1, 2)
    q = np.nonzero(is_quad)[0]
    t = np.nonzero(~is_quad)[0]
    faces[start[q]] = np.column_stack([a[q], b[q], c[q], d[q]])
    faces[start[t], :3] = t1[t]
    faces[start[t] + 1, :3] = t2[t]
    return xy, faces


def disk(partitions: int = 6, depth: int = 8) -> Tuple[np.ndarray, np.ndarray]:
    from scipy.spatial import Delaunay

    N = depth + 1
    n_per_level = partitions * np.arange(N)
    n_per_level[0] = 1
    delta_angle = (2 * np.pi) / np.repeat(n_per_level, n_per_level)
    index = np.repeat(np.insert(n_per_level.cumsum()[:-1], 0, 0), n_per_level)
    angles = delta_angle.cumsum()
    angles = angles - angles[index] + 0.5 * np.pi
    radii = np.repeat(np.linspace(0.0, 1.0, N), n_per_level)
    x = np.cos(angles) * radii * 10.0
    y = np.sin(angles) * radii * 10.0
    pts = np.column_stack([x, y])
    tri = Delaunay(pts).simplices.astype(np.int64)
    # drop degenerate hull slivers qhull may add between near-collinear rim points
    p = pts[tri]
    area2 = np.abs(
        (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
        - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    )
    return pts, np.ascontiguousarray(tri[area2 > 1e-9])


def function_z(x, y):
    """xugrid/data/synthetic.py:64-82"""
    r1 = np.sqrt((0.5 - x) ** 2 + (0.5 - y) ** 2)
    theta1 = np.arctan2(0.5 - x, 0.5 - y)
    r2 = np.sqrt((-x - 0.2) ** 2 + (-y - 0.2) ** 2)
    theta2 = np.arctan2(-x - 0.2, -y - 0.2)
    z = -(
        2 * (np.exp((r1 / 10) ** 2) - 1) * 30.0 * np.cos(7.0 * theta1)
        + (np.exp((r2 / 10) ** 2) - 1) * 30.0 * np.cos(11.0 * theta2)
        + 0.7 * (x**2 + y**2)
    )
    zmin = z.min()
    zmax = z.max()
    return (zmax - z) / (zmax - zmin) * 10.0
