"""CPU test of the Voronoi preparation (torch ops on CPU tensors) against golden
vectors produced by the reference's own voronoi.py / connectivity.py /
replace_interpolated_weights (oracle/make_golden.py::golden_barycentric)."""
import numpy as np
import pytest
import torch

import oracle
from xugrid_b200 import voronoi


def barycentric_triplets_cpu(xy, faces, points):
    """UnstructuredGrid2d.barycentric with the product's Voronoi prep and the
    oracle standing in for the two CUDA queries (no GPU in this test)."""
    t_faces = torch.from_numpy(faces)
    t_xy = torch.from_numpy(xy)
    cent = torch.from_numpy(oracle.face_centroids(xy, faces))
    topo = voronoi.voronoi_topology(t_faces, t_xy, cent)
    vfaces = topo.faces.numpy()
    tree = oracle.CellTree2d(topo.vertices.numpy(), vfaces, -1)
    face_index, weights = tree.compute_barycentric_weights(points)
    conn = topo.faces[torch.from_numpy(face_index).clamp(min=0)]
    w = torch.from_numpy(weights)
    voronoi.redistribute_corner_weights(topo, conn, w)
    outside = oracle.CellTree2d(xy, faces, -1).locate_points(points) == -1
    w[torch.from_numpy(outside)] = 0
    keep = w > 0
    src = topo.vertex_face[conn.clamp(min=0)][keep].numpy()
    tgt = torch.arange(len(points))[:, None].expand_as(w)[keep].numpy()
    wv = w[keep].numpy()
    order = np.lexsort((wv, src, tgt))
    return tgt[order], src[order], wv[order], topo


@pytest.mark.parametrize("name", ["disk", "mixed"])
def test_voronoi_barycentric_matches_reference(golden_dir, name):
    g = np.load(f"{golden_dir}/barycentric_golden.npz")
    xy, faces, points = g[f"{name}_xy"], g[f"{name}_faces"], g[f"{name}_points"]
    tgt, src, w, topo = barycentric_triplets_cpu(xy, faces, points)
    # same Voronoi cells (compared as the sorted list of polygon areas)
    assert topo.vertices.shape[0] == int(g[f"{name}_vor_n_vertex"])
    area = np.sort(oracle.face_area(topo.vertices.numpy(), topo.faces.numpy()))
    assert np.allclose(area, g[f"{name}_vor_area"], rtol=1e-10, atol=1e-14)
    # same interpolation weights
    assert np.array_equal(tgt, g[f"{name}_target"])
    assert np.array_equal(src, g[f"{name}_source"])
    assert np.allclose(w, g[f"{name}_weight"], rtol=1e-10, atol=1e-15)
    # weights of a located point sum to one
    sums = np.bincount(tgt, weights=w, minlength=len(points))
    assert np.allclose(sums[sums > 0], 1.0, rtol=1e-12)


def test_voronoi_polygons_are_ccw_and_tile_the_mesh():
    from xugrid_b200 import synthetic

    xy, faces = synthetic.jittered_lattice(9, 7, 3, "mixed")
    cent = oracle.face_centroids(xy, faces)
    topo = voronoi.voronoi_topology(torch.from_numpy(faces), torch.from_numpy(xy), torch.from_numpy(cent))
    v, f = topo.vertices.numpy(), topo.faces.numpy()
    assert np.array_equal(oracle.counter_clockwise(v, f), f)
    # one polygon per node, every face centroid used by as many polygons as the face has nodes
    assert f.shape[0] == xy.shape[0]
    used = np.bincount(f[f >= 0], minlength=v.shape[0])[: len(faces)]
    assert np.array_equal(used, (faces >= 0).sum(axis=1))


def test_voronoi_rejects_non_manifold_exterior():
    # two triangles touching in one node: that node has four exterior edges
    xy = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]], float)
    faces = np.array([[0, 1, 2], [0, 3, 4]])
    cent = oracle.face_centroids(xy, faces)
    with pytest.raises(ValueError):
        voronoi.voronoi_topology(torch.from_numpy(faces), torch.from_numpy(xy), torch.from_numpy(cent))


def test_single_sort_by_group_then_angle_is_the_exact_lexsort():
    """the one-pass sort (group id | 32-bit angle) returns the order of the exact two-pass
    lexsort, also when the quantised angles collide (fallback) and for equal angles (stable)"""
    rng = np.random.default_rng(0)
    group = torch.from_numpy(rng.integers(0, 500, 20000))
    angle = torch.from_numpy(rng.uniform(-np.pi, np.pi, 20000))
    angle[::97] = angle[1::97][: angle[::97].numel()]          # exact ties
    assert torch.equal(voronoi._sort_by_group_then_angle(angle, group), voronoi._lexsort(angle, group))
    close = torch.tensor([0.1, 0.1 + 1e-12, 0.1 - 1e-12, 2.0, -3.0], dtype=torch.float64)
    g0 = torch.zeros(5, dtype=torch.int64)
    assert torch.equal(voronoi._sort_by_group_then_angle(close, g0), voronoi._lexsort(close, g0))
    empty = torch.zeros(0, dtype=torch.float64)
    assert voronoi._sort_by_group_then_angle(empty, torch.zeros(0, dtype=torch.int64)).numel() == 0
