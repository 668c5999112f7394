"""
Centroidal Voronoi topology of a mesh, as torch tensor ops (run on the GPU in
the product, on CPU tensors in the tests).

This is the preparation step of ``BarycentricInterpolator``
(/root/reference/xugrid/regrid/unstructured.py:146-201): data lives on the
source FACES, so interpolation happens inside polygons whose vertices are the
centroids of the faces around every source NODE.  The reference builds them in
``xugrid/ugrid/voronoi.py:330-458`` (``voronoi_topology(add_exterior=True,
add_vertices=True, skip_concave=True)``) with numpy / pandas / scipy.sparse on
the host; SURVEY.md section 8f ranks moving it to the device first among the
follow-ups, because at 5 M nodes the host version dominates the wall clock.

What is built (same geometry as the reference, own formulation):

* a node that touches no exterior edge gets the polygon of its adjacent face
  centroids, ordered counter-clockwise by angle around the node;
* an exterior node gets its adjacent face centroids, the orthogonal projections
  of the boundary faces' centroids on the two exterior edges that meet at the
  node, and one corner vertex.  The corner is first the midpoint of the two
  projections (always convex, used to order the polygon by angle around its
  mean), then replaced by the node itself unless that makes the polygon
  smaller (= concave corner).  Deviation: the reference compares the two areas
  with ``>=`` on rounded sums, so on a perfectly straight boundary, where the
  areas are equal, its choice depends on rounding noise; here equal areas
  (relative 1e-12) take the node, which is what exact arithmetic gives.

Vertex numbering: ``[face centroids | kept projections | one corner per
exterior node]``; ``vertex_face`` maps a vertex to the source face whose value
it carries (-1 for corners, whose weight is redistributed to the two
projections by ``redistribute_corner_weights``, the counterpart of
``replace_interpolated_weights``, unstructured.py:17-57).
"""

from __future__ import annotations

from typing import NamedTuple

import torch

X_EPSILON = 1.0e-8  # /root/reference/xugrid/constants.py:45


class VoronoiTopology(NamedTuple):
    vertices: torch.Tensor      # (n_vertex, 2) float64
    faces: torch.Tensor         # (n_polygon, k_max) int64, -1 fill, counter-clockwise
    vertex_face: torch.Tensor   # (n_vertex,) int64 source face per vertex, -1 for corners
    corner_map: torch.Tensor    # (n_corner, 2) int64: the two projection vertices of every corner
    corner_start: int           # vertices >= corner_start are corners


def _lexsort(minor: torch.Tensor, major: torch.Tensor) -> torch.Tensor:
    o1 = torch.argsort(minor, stable=True)
    o2 = torch.argsort(major[o1], stable=True)
    return o1[o2]


def _sort_by_group_then_angle(angle: torch.Tensor, group: torch.Tensor) -> torch.Tensor:
    """Order of ``_lexsort(angle, group)`` (group ids >= 0, angles in [-pi, pi]) with ONE
    sort instead of two: the key is the group id in the high 32 bits and the angle
    quantised to 32 bits in the low ones.  The result is then checked against the
    float64 angles (non-decreasing inside every group); a mesh whose angles the
    quantisation cannot separate falls back to the exact two-pass sort, so the order
    returned is always the exact one."""
    if angle.numel() == 0 or int(group.max().item()) >= (1 << 31):
        return _lexsort(angle, group)
    q = ((angle + torch.pi) * (4294967295.0 / (2.0 * torch.pi))).clamp_(0.0, 4294967295.0).to(torch.int64)
    key = (group.to(torch.int64) << 32) | q
    order = torch.argsort(key, stable=True)
    a, g = angle[order], group[order]
    exact = bool(((a[1:] >= a[:-1]) | (g[1:] != g[:-1])).all())
    return order if exact else _lexsort(angle, group)


def _ragged_to_dense(group: torch.Tensor, values: torch.Tensor, n_group: int) -> torch.Tensor:
    """group must be sorted ascending with ids 0..n_group-1; returns (n_group, max_len), -1 fill."""
    counts = torch.bincount(group, minlength=n_group)
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.arange(group.numel(), device=group.device) - starts[group]
    k = int(counts.max().item()) if group.numel() else 0
    dense = torch.full((n_group, k), -1, dtype=torch.int64, device=group.device)
    dense[group, pos] = values
    return dense


def exterior_edges(faces: torch.Tensor, n_node: int):
    """Directed exterior edges (a -> b in the face's own vertex order) and their face."""
    F, k = faces.shape
    dev = faces.device
    valid = faces >= 0
    length = valid.sum(1)
    col = torch.arange(k, device=dev)[None, :].expand(F, k)
    nxt = torch.where(col + 1 < length[:, None], col + 1, torch.zeros_like(col))
    a = faces[valid]
    b = torch.gather(faces, 1, nxt)[valid]
    f = torch.arange(F, device=dev)[:, None].expand(F, k)[valid]
    ok = a != b
    a, b, f = a[ok], b[ok], f[ok]
    key = torch.minimum(a, b) * n_node + torch.maximum(a, b)
    # an edge is exterior when its key occurs once: one sort and a comparison with both
    # neighbours (torch.unique with inverse and counts sorts too, then makes three more passes)
    sk, order = torch.sort(key)
    once = torch.ones_like(sk, dtype=torch.bool)
    once[1:] &= sk[1:] != sk[:-1]
    once[:-1] &= sk[:-1] != sk[1:]
    ext = torch.zeros_like(once)
    ext[order] = once
    return a[ext], b[ext], f[ext]


def voronoi_topology(faces: torch.Tensor, xy: torch.Tensor, centroids: torch.Tensor) -> VoronoiTopology:
    faces = faces.to(torch.int64)
    dev = faces.device
    F, k = faces.shape
    N = xy.shape[0]
    valid = faces >= 0
    inc_node = faces[valid]
    inc_face = torch.arange(F, device=dev)[:, None].expand(F, k)[valid]

    ea, eb, ef = exterior_edges(faces, N)
    n_edge = ea.numel()
    is_ext = torch.zeros(N, dtype=torch.bool, device=dev)
    is_ext[ea] = True
    is_ext[eb] = True
    ext_nodes = torch.nonzero(is_ext).flatten()
    n_ext = ext_nodes.numel()
    if n_ext != n_edge:
        raise ValueError("the mesh exterior is not a set of simple closed loops "
                         f"({n_edge} exterior edges, {n_ext} exterior nodes)")
    rank = torch.full((N,), -1, dtype=torch.int64, device=dev)
    rank[ext_nodes] = torch.arange(n_ext, device=dev)
    out_edge = torch.full((n_ext,), -1, dtype=torch.int64, device=dev)
    in_edge = torch.full((n_ext,), -1, dtype=torch.int64, device=dev)
    e_ids = torch.arange(n_edge, device=dev)
    out_edge[rank[ea]] = e_ids
    in_edge[rank[eb]] = e_ids
    if bool((out_edge < 0).any()) or bool((in_edge < 0).any()):
        raise ValueError("an exterior node is not met by exactly one incoming and one outgoing exterior edge")

    # ---- interior polygons: face centroids by angle around the node
    interior = ~is_ext[inc_node]
    n_i, f_i = inc_node[interior], inc_face[interior]
    d = centroids[f_i] - xy[n_i]
    order = _sort_by_group_then_angle(torch.atan2(d[:, 1], d[:, 0]), n_i)
    n_i, f_i = n_i[order], f_i[order]
    int_nodes, int_group = torch.unique_consecutive(n_i, return_inverse=True)  # n_i is sorted
    interior_faces = _ragged_to_dense(int_group, f_i, int_nodes.numel())

    # ---- exterior: projections of boundary-face centroids on their exterior edge
    lo, hi = torch.minimum(ea, eb), torch.maximum(ea, eb)
    A, B = xy[lo], xy[hi]
    V = B - A
    U = centroids[ef] - A
    t = (U[:, 1] * V[:, 1] + U[:, 0] * V[:, 0]) / (V[:, 1] * V[:, 1] + V[:, 0] * V[:, 0])
    P = A + t[:, None] * V
    keep = torch.linalg.norm(P - centroids[ef], dim=1) > X_EPSILON * X_EPSILON
    n_keep = int(keep.sum().item())
    pid = torch.where(keep, F + torch.cumsum(keep.to(torch.int64), 0) - 1, ef)  # dropped -> the centroid itself
    corner_start = F + n_keep
    corner_id = corner_start + torch.arange(n_ext, device=dev)
    S = 0.5 * (P[in_edge] + P[out_edge])
    corner_map = torch.stack([pid[in_edge], pid[out_edge]], dim=1)

    # members of every exterior polygon: (node rank, vertex id, coordinates)
    ext_inc = ~interior
    m_rank = [rank[inc_node[ext_inc]], rank[ea[keep]], rank[eb[keep]], torch.arange(n_ext, device=dev)]
    m_vert = [inc_face[ext_inc], pid[keep], pid[keep], corner_id]
    m_xy = [centroids[inc_face[ext_inc]], P[keep], P[keep], S]
    m_rank, m_vert, m_xy = torch.cat(m_rank), torch.cat(m_vert), torch.cat(m_xy)
    cnt = torch.bincount(m_rank, minlength=n_ext).to(torch.float64)
    mean = torch.zeros((n_ext, 2), dtype=torch.float64, device=dev).index_add_(0, m_rank, m_xy) / cnt[:, None]
    dd = m_xy - mean[m_rank]
    order = _lexsort(torch.atan2(dd[:, 1], dd[:, 0]), m_rank)
    m_rank, m_vert, m_xy = m_rank[order], m_vert[order], m_xy[order]

    # corner choice: the node itself unless it makes the polygon smaller
    counts = torch.bincount(m_rank, minlength=n_ext)
    starts = torch.cumsum(counts, 0) - counts
    idx = torch.arange(m_rank.numel(), device=dev)
    nxt = torch.where(idx + 1 < (starts + counts)[m_rank], idx + 1, starts[m_rank])
    is_corner = m_vert >= corner_start
    xy_node = torch.where(is_corner[:, None], xy[ext_nodes][m_rank], m_xy)

    def poly_area(c):
        rel = c - c[starts][m_rank]
        cr = rel[:, 0] * rel[nxt, 1] - rel[:, 1] * rel[nxt, 0]
        return 0.5 * torch.zeros(n_ext, dtype=torch.float64, device=dev).index_add_(0, m_rank, cr).abs()

    area_mid, area_node = poly_area(m_xy), poly_area(xy_node)
    use_node = area_node >= area_mid * (1.0 - 1e-12)
    corner_xy = torch.where(use_node[:, None], xy[ext_nodes], S)

    exterior_faces = _ragged_to_dense(m_rank, m_vert, n_ext)
    kmax = max(interior_faces.shape[1], exterior_faces.shape[1])

    def pad(a):
        if a.shape[1] == kmax:
            return a
        return torch.cat([a, torch.full((a.shape[0], kmax - a.shape[1]), -1, dtype=torch.int64, device=dev)], 1)

    vor_faces = torch.cat([pad(interior_faces), pad(exterior_faces)], 0)
    vertices = torch.cat([centroids, P[keep], corner_xy], 0)
    vertex_face = torch.cat([torch.arange(F, device=dev), ef[keep],
                             torch.full((n_ext,), -1, dtype=torch.int64, device=dev)])
    return VoronoiTopology(vertices, vor_faces, vertex_face, corner_map, corner_start)


def voronoi_topology_device(faces: torch.Tensor, xy: torch.Tensor, centroids: torch.Tensor) -> VoronoiTopology:
    """The same topology built by the library's kernels (csrc/xg_voronoi.cu; CUDA tensors
    in, int32 index tensors out): what ``UnstructuredGrid2d.barycentric`` runs.  The
    tensor-op version above stays as the statement the kernels are tested against."""
    import ctypes

    from . import _native

    L = _native.lib()
    dev = xy.device
    F, k = int(faces.shape[0]), int(faces.shape[1])
    sizes = (ctypes.c_int64 * 4)()
    job = ctypes.c_void_p()
    with torch.cuda.device(dev):
        s = _native.stream_ptr()
        _native.check(L.xg_voronoi_begin(_native.ptr(xy, torch.float64), int(xy.shape[0]), _native.ptr(faces),
                                         faces.element_size(), F, k, _native.ptr(centroids, torch.float64), sizes, s,
                                         ctypes.byref(job)))
        try:
            n_int, n_ext, n_keep, kmax = (int(v) for v in sizes)
            n_vertex = F + n_keep + n_ext
            vertices = torch.empty((n_vertex, 2), dtype=torch.float64, device=dev)
            polygons = torch.empty((n_int + n_ext, kmax), dtype=torch.int32, device=dev)
            vertex_face = torch.empty(n_vertex, dtype=torch.int32, device=dev)
            corner_map = torch.empty((n_ext, 2), dtype=torch.int32, device=dev)
        except BaseException:
            L.xg_voronoi_abort(job)
            raise
        _native.check(L.xg_voronoi_finish(job, _native.ptr(vertices), _native.ptr(polygons), _native.ptr(vertex_face),
                                          _native.ptr(corner_map), s))
    return VoronoiTopology(vertices, polygons, vertex_face, corner_map, F + n_keep)


def weights_to_csr_device(topo: VoronoiTopology, polygon_index: torch.Tensor, weights: torch.Tensor,
                          source_face: torch.Tensor):
    """unstructured.py:176-201 on the device (``xg_bary_csr_count`` / ``_fill``): corner
    weights redistributed (``weights`` is adjusted in place), points outside the source
    mesh (``source_face == -1``) dropped, positive weights as CSR rows.  Returns
    (indptr int64, indices int32, data float64)."""
    import ctypes

    from . import _native

    L = _native.lib()
    dev = weights.device
    n, kv = int(weights.shape[0]), int(weights.shape[1])
    if kv != int(topo.faces.shape[1]):
        raise ValueError("weights and Voronoi polygons disagree on the number of vertices per polygon")
    indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    nnz = ctypes.c_int64()
    with torch.cuda.device(dev):
        s = _native.stream_ptr()
        _native.check(L.xg_bary_csr_count(
            _native.ptr(polygon_index, torch.int32), _native.ptr(weights, torch.float64), n, kv,
            _native.ptr(topo.faces, torch.int32), _native.ptr(topo.vertex_face, torch.int32),
            _native.ptr(topo.corner_map, torch.int32), _native.ptr(topo.vertices, torch.float64),
            int(topo.corner_start), _native.ptr(source_face, torch.int32), _native.ptr(indptr), ctypes.byref(nnz), s))
        indices = torch.empty(nnz.value, dtype=torch.int32, device=dev)
        data = torch.empty(nnz.value, dtype=torch.float64, device=dev)
        _native.check(L.xg_bary_csr_fill(
            _native.ptr(polygon_index, torch.int32), _native.ptr(weights, torch.float64), n, kv,
            _native.ptr(topo.faces, torch.int32), _native.ptr(topo.vertex_face, torch.int32), _native.ptr(indptr),
            _native.ptr(indices), _native.ptr(data), s))
    return indptr, indices, data


def redistribute_corner_weights(topo: VoronoiTopology, conn: torch.Tensor, weights: torch.Tensor) -> None:
    """
    In place.  ``conn`` (n_point, k) are the vertices of the polygon every point
    fell in, ``weights`` its barycentric weights.  The weight of a corner vertex
    goes to the two projection vertices it sits between, inversely proportional
    to distance (unstructured.py:17-57).
    """
    is_corner = (conn >= topo.corner_start) & (weights > 0)
    rows = torch.nonzero(is_corner.any(dim=1)).flatten()
    if rows.numel() == 0:
        return
    col = torch.argmax(is_corner[rows].to(torch.int8), dim=1)
    p = conn[rows, col]
    w = weights[rows, col]
    q = topo.corner_map[p - topo.corner_start, 0]
    r = topo.corner_map[p - topo.corner_start, 1]
    vp, vq, vr = topo.vertices[p], topo.vertices[q], topo.vertices[r]
    p_q = torch.sqrt((vq[:, 0] - vp[:, 0]) ** 2 + (vq[:, 1] - vp[:, 1]) ** 2)
    p_r = torch.sqrt((vr[:, 0] - vp[:, 0]) ** 2 + (vr[:, 1] - vp[:, 1]) ** 2)
    total = p_q + p_r
    weight_q = (p_r / total) * w
    weight_r = (p_q / total) * w
    sub = weights[rows]
    sub[torch.arange(rows.numel(), device=rows.device), col] = 0.0
    c = conn[rows]
    sub = sub + (c == q[:, None]) * weight_q[:, None] + (c == r[:, None]) * weight_r[:, None]
    weights[rows] = sub
