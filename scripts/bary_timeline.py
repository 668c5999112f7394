"""Stage times of the BarycentricInterpolator weight build (config 3)."""
import sys
import numpy as np
import torch
import xugrid_b200 as xb
from xugrid_b200 import synthetic, voronoi
from xugrid_b200.celltree import CellTree2d

ns, nr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2236, 4096)
dev = torch.device("cuda", 0)
sxy, sf = synthetic.jittered_lattice(ns, ns, 20, "tri")
grid = xb.Ugrid2d(sxy[:, 0], sxy[:, 1], -1, sf, device=dev)
c = np.linspace(0.02, 0.98, nr + 1)
mid = 0.5 * (c[:-1] + c[1:])
yy, xx = np.meshgrid(mid, mid, indexing="ij")
points = torch.from_numpy(np.column_stack([xx.ravel(), yy.ravel()])).to(dev)
xy, faces = grid.device_arrays()
torch.cuda.synchronize()

def run():
    marks = []
    def mark(name):
        ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append((name, ev))
    mark("start")
    cen = grid.centroids_device; mark("centroids")
    topo = voronoi.voronoi_topology(faces, xy, cen); mark("voronoi_topology")
    vor_tree = CellTree2d(topo.vertices, topo.faces, -1, device=dev); mark("voronoi index")
    face_index, weights = vor_tree.compute_barycentric_weights_device(points, None); mark("barycentric kernel")
    conn = topo.faces[face_index.long().clamp(min=0)]; mark("conn gather")
    voronoi.redistribute_corner_weights(topo, conn, weights); mark("corner weights")
    tree = CellTree2d(xy, faces, -1, device=dev); mark("source index")
    outside = tree.locate_points_device(points) == -1; mark("locate (outside mask)")
    weights[outside] = 0
    keep = weights > 0
    source_index = topo.vertex_face[conn.clamp(min=0)]
    counts = keep.sum(dim=1)
    indptr = torch.zeros(points.shape[0] + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr[1:])
    indices = source_index[keep].to(torch.int32)
    data = weights[keep]; mark("CSR assembly")
    torch.cuda.synchronize()
    return marks, topo

run()
marks, topo = run()
print("voronoi polygons", tuple(topo.faces.shape), "vertices", tuple(topo.vertices.shape))
prev = marks[0][1]
for name, ev in marks[1:]:
    print(f"  {prev.elapsed_time(ev):9.2f} ms  {name}")
    prev = ev
print(f"  {marks[0][1].elapsed_time(marks[-1][1]):9.2f} ms  total")
