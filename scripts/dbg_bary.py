import numpy as np, torch, oracle
import xugrid_b200 as xb
g = np.load("tests/golden/barycentric_golden.npz")
name="mixed"
xy, faces, points = g[f"{name}_xy"], g[f"{name}_faces"], g[f"{name}_points"]
source = xb.Ugrid2d(xy[:, 0], xy[:, 1], -1, faces)
d = 1e-3
tri = np.array([[-d, -d], [2 * d, -d], [-d, 2 * d]])
txy = (points[:, None, :] + tri[None, :, :]).reshape(-1, 2)
tf = np.arange(3 * len(points)).reshape(-1, 3)
target = xb.Ugrid2d(txy[:, 0], txy[:, 1], -1, tf)
r = xb.BarycentricInterpolator(source, target)
data, indices, indptr = r._weights.to_numpy()
tgt = np.repeat(np.arange(len(points)), np.diff(indptr))
order = np.lexsort((data, indices, tgt))
dd = np.abs(data[order]-g[f"{name}_weight"])
bad = np.nonzero(dd > 1e-9)[0]
print(len(bad), dd.max())
for b in bad[:10]:
    t = tgt[order][b]
    print("t", t, points[t], "src", indices[order][b], data[order][b], g[f"{name}_weight"][b])
    sel = tgt[order]==t
    print("   gpu:", indices[order][sel], data[order][sel]); print("   ref:", g[f"{name}_source"][g[f"{name}_target"]==t], g[f"{name}_weight"][g[f"{name}_target"]==t])
