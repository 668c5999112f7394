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
    topo = voronoi.voronoi_topology_device(faces, xy, cen); mark("voronoi_topology (kernels)")
    vor_tree = CellTree2d(topo.vertices, topo.faces, -1, device=dev); mark("voronoi index")
    polygon_index, weights = vor_tree.compute_barycentric_weights_device(points, None); mark("barycentric kernel")
    tree = CellTree2d(xy, faces, -1, device=dev); mark("source index")
    source_face = tree.locate_points_device(points, any_face=True); mark("locate (outside mask)")
    indptr, indices, data = voronoi.weights_to_csr_device(topo, polygon_index, weights, source_face); mark("CSR assembly (kernels)")
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
