"""The Voronoi preparation and the CSR assembly of BarycentricInterpolator as dedicated
kernels (csrc/xg_voronoi.cu) against their tensor-op statement (xugrid_b200/voronoi.py, itself
pinned by golden vectors from the reference's voronoi.py in tests/test_voronoi_cpu.py), and the
whole GPU pipeline against those golden vectors at the golden points."""
import numpy as np
import pytest
import torch

import oracle
from xugrid_b200 import synthetic, voronoi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xugrid_b200

    assert torch.cuda.is_available()
    return xugrid_b200


def l_shape(kind, seed):
    """a lattice with one quadrant removed: a concave corner, and nodes no face uses"""
    xy, faces = synthetic.jittered_lattice(14, 12, seed, kind)
    c = oracle.face_centroids(xy, faces)
    return xy, np.ascontiguousarray(faces[~((c[:, 0] > 0.5) & (c[:, 1] > 0.5))])


MESHES = {
    "disk": lambda: synthetic.disk(),
    "tri": lambda: synthetic.jittered_lattice(23, 17, 5, "tri"),
    "mixed": lambda: synthetic.jittered_lattice(19, 21, 6, "mixed"),
    "quad": lambda: synthetic.jittered_lattice(11, 13, 7, "quad"),
    "delaunay": lambda: synthetic.delaunay(3000, 8),
    "l_tri": lambda: l_shape("tri", 9),
    "l_mixed": lambda: l_shape("mixed", 10),
    "straight": lambda: synthetic.jittered_lattice(6, 5, 11, "tri", jitter=0.0),
}


def both_topologies(xb, name, int64_faces=False):
    from xugrid_b200.celltree import face_centroid_device

    xy, faces = MESHES[name]()
    dev = torch.device("cuda", 0)
    dxy = torch.from_numpy(xy).to(dev)
    dfaces = torch.from_numpy(faces).to(dev)
    dfaces = dfaces.to(torch.int64 if int64_faces else torch.int32)
    cen = face_centroid_device(dxy, dfaces)
    return xy, faces, voronoi.voronoi_topology(dfaces, dxy, cen), voronoi.voronoi_topology_device(dfaces, dxy, cen)


@pytest.mark.parametrize("name", list(MESHES))
def test_kernels_build_the_same_topology(xb, name):
    xy, faces, ref, got = both_topologies(xb, name, int64_faces=(name == "mixed"))
    assert got.corner_start == ref.corner_start
    assert got.vertices.shape == ref.vertices.shape
    assert torch.equal(got.vertex_face.long(), ref.vertex_face)
    assert torch.equal(got.corner_map.long(), ref.corner_map)
    # projections and corners bit for bit (same expressions, no contraction)
    assert torch.equal(got.vertices.view(torch.int64), ref.vertices.view(torch.int64))
    assert got.faces.shape == ref.faces.shape
    assert torch.equal(got.faces.long(), ref.faces)


def test_non_manifold_exterior_is_rejected(xb):
    from xugrid_b200 import _native

    xy = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]], float)
    faces = np.array([[0, 1, 2], [0, 3, 4]], dtype=np.int32)
    dev = torch.device("cuda", 0)
    cen = torch.from_numpy(oracle.face_centroids(xy, faces)).to(dev)
    with pytest.raises(_native.XgError):
        voronoi.voronoi_topology_device(torch.from_numpy(faces).to(dev), torch.from_numpy(xy).to(dev), cen)


@pytest.mark.parametrize("name", ["disk", "mixed", "l_tri", "delaunay"])
def test_csr_assembly_kernels_equal_the_tensor_ops(xb, name):
    from xugrid_b200.celltree import CellTree2d

    xy, faces, ref, got = both_topologies(xb, name)
    dev = got.vertices.device
    lo, hi = xy.min(axis=0), xy.max(axis=0)
    pts = np.random.default_rng(3).random((20000, 2)) * (hi - lo) * 1.1 + lo - 0.05 * (hi - lo)
    pts = np.concatenate([pts, xy[:200], 0.5 * (xy[:150] + xy[1:151])])  # + nodes and points near edges
    points = torch.from_numpy(pts).to(dev)
    tree = CellTree2d(got.vertices, got.faces, -1, device=dev)
    polygon_index, weights = tree.compute_barycentric_weights_device(points, None)
    source_face = CellTree2d(xy, faces, -1, device=dev).locate_points_device(points)
    # tensor-op statement (what UnstructuredGrid2d.barycentric did before the kernels existed)
    w = weights.clone()
    conn = ref.faces[polygon_index.long().clamp(min=0)]
    voronoi.redistribute_corner_weights(ref, conn, w)
    w[source_face == -1] = 0
    keep = w > 0
    counts = keep.sum(dim=1)
    indptr_ref = torch.zeros(points.shape[0] + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr_ref[1:])
    indices_ref = ref.vertex_face[conn.clamp(min=0)][keep]
    data_ref = w[keep]
    indptr, indices, data = voronoi.weights_to_csr_device(got, polygon_index, weights, source_face)
    assert torch.equal(indptr, indptr_ref)
    assert torch.equal(indices.long(), indices_ref)
    assert torch.equal(data.view(torch.int64), data_ref.view(torch.int64))
    assert torch.equal(weights.view(torch.int64), w.view(torch.int64))  # adjusted in place, the same way
    assert int((counts > 0).sum()) > 1000


@pytest.mark.parametrize("name", ["disk", "mixed"])
def test_gpu_pipeline_reproduces_the_reference_weights_at_the_golden_points(xb, golden_dir, name):
    """BarycentricInterpolator's weights from the reference's own code (oracle/make_golden.py)
    at the very points they were computed for: 1e-10, not the 1e-6 of a tiny-triangle target"""
    g = np.load(f"{golden_dir}/barycentric_golden.npz")
    xy, faces, points = g[f"{name}_xy"], g[f"{name}_faces"], g[f"{name}_points"]
    dev = torch.device("cuda", 0)
    from xugrid_b200.celltree import CellTree2d, face_centroid_device

    dxy, dfaces = torch.from_numpy(xy).to(dev), torch.from_numpy(faces).to(dev)
    topo = voronoi.voronoi_topology_device(dfaces, dxy, face_centroid_device(dxy, dfaces))
    assert topo.vertices.shape[0] == int(g[f"{name}_vor_n_vertex"])
    area = np.sort(oracle.face_area(topo.vertices.cpu().numpy(), topo.faces.cpu().numpy().astype(np.int64)))
    assert np.allclose(area, g[f"{name}_vor_area"], rtol=1e-10, atol=1e-14)
    pts = torch.from_numpy(points).to(dev)
    polygon_index, weights = CellTree2d(topo.vertices, topo.faces, -1, device=dev).compute_barycentric_weights_device(pts, None)
    source_face = CellTree2d(dxy, dfaces, -1, device=dev).locate_points_device(pts)
    indptr, indices, data = voronoi.weights_to_csr_device(topo, polygon_index, weights, source_face)
    ip, src, w = indptr.cpu().numpy(), indices.cpu().numpy().astype(np.int64), data.cpu().numpy()
    tgt = np.repeat(np.arange(len(points)), np.diff(ip))
    order = np.lexsort((w, src, tgt))
    assert np.array_equal(tgt[order], g[f"{name}_target"])
    assert np.array_equal(src[order], g[f"{name}_source"])
    assert np.allclose(w[order], g[f"{name}_weight"], rtol=1e-10, atol=1e-15)


@pytest.mark.parametrize("name", ["mixed", "delaunay", "l_tri"])
def test_locate_any_face_gives_the_same_mask(xb, name):
    """xg_locate_any (what the outside mask of the barycentric build asks): -1 exactly where
    locate_points says -1; a reported face really holds the point"""
    from xugrid_b200.celltree import CellTree2d

    xy, faces = MESHES[name]()
    dev = torch.device("cuda", 0)
    lo, hi = xy.min(axis=0), xy.max(axis=0)
    pts = np.random.default_rng(7).random((30000, 2)) * (hi - lo) * 1.2 + lo - 0.1 * (hi - lo)
    m = min(200, len(xy) - 1)
    pts = np.concatenate([pts, xy[:300], 0.5 * (xy[:m] + xy[1: m + 1]), [[np.nan, 0.0]]])  # + nodes, edge points, NaN
    tree = CellTree2d(xy, faces, -1, device=dev)
    exact = tree.locate_points_device(pts).cpu().numpy()
    some = tree.locate_points_device(pts, any_face=True).cpu().numpy()
    assert np.array_equal(exact == -1, some == -1)
    assert int((exact == -1).sum()) > 100 and int((exact >= 0).sum()) > 10000
    # where both found a face it is the same one unless the point sits on a shared edge or node
    differ = np.nonzero((exact != some) & (exact >= 0))[0]
    assert len(differ) <= 600
    ref = oracle.CellTree2d(xy, faces, -1)
    for i in differ[:50]:
        # the face reported by the any-variant holds the point too: locate restricted to that face
        one = oracle.CellTree2d(xy, faces[some[i]: some[i] + 1], -1).locate_points(pts[i: i + 1])
        assert one[0] == 0, (i, exact[i], some[i])
