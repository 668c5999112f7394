"""
Generates tests/golden/*.npz by running the UNMODIFIED reference code that is
runnable in the build container (see oracle/ref_loader.py).  Run from the repo
root in a container that has /root/reference:

    python -m oracle.make_golden

The committed .npz files are what travels to the GPU box; this script is the
provenance record.  Test infrastructure only.
"""

from __future__ import annotations

import os

import numpy as np

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_apply_case(seed=1234, n=400, m=300, n_extra=5):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 10, n)
    counts[:4] = [0, 1, 2, 9]
    counts[-1] = 0
    indptr = np.r_[0, np.cumsum(counts)].astype(np.int64)
    nnz = int(indptr[-1])
    indices = rng.integers(0, m, nnz).astype(np.int64)
    data = rng.random(nnz)
    # exact zero weights, an all-zero row, duplicated weights (ties in max_overlap)
    data[rng.random(nnz) < 0.05] = 0.0
    r = 7
    data[indptr[r] : indptr[r + 1]] = 0.0
    r = 9
    data[indptr[r] : indptr[r + 1]] = 0.25
    src = np.empty((n_extra, m))
    src[0] = rng.random(m) + 0.1  # strictly positive
    src[1] = rng.random(m)
    src[1, rng.random(m) < 0.2] = np.nan  # NaNs
    src[2] = rng.integers(0, 4, m).astype(float)  # ties -> mode, zeros
    src[3] = rng.normal(size=m)  # negatives
    src[3, rng.random(m) < 0.1] = 0.0
    src[4] = np.nan  # everything missing
    return dict(indptr=indptr, indices=indices, data=data, source=src, n=n, m=m)


def golden_apply(ns):
    case = make_apply_case()
    A = ns.sparse.MatrixCSR(
        case["data"], case["indices"], case["indptr"], case["n"], case["m"], int(case["indptr"][-1])
    )
    out = dict(case)
    methods = dict(ns.reduce.ABSOLUTE_OVERLAP_METHODS)
    methods.update(ns.reduce.RELATIVE_OVERLAP_METHODS)
    for name, func in methods.items():
        f = ns.make_regrid(func)
        out["out_" + name] = f(case["source"], A, case["n"])
        print("apply", name)
    for p in (0.0, 33.3, 100.0):
        f = ns.make_regrid(ns.reduce.create_percentile_method(p))
        out[f"out_percentile_{p}"] = f(case["source"], A, case["n"])
    # integer source (fixtures use np.arange ints: fixture_regridder.py:96)
    isrc = np.arange(case["m"] * 2, dtype=np.int64).reshape(2, case["m"])
    out["int_source"] = isrc
    out["out_int_mean"] = ns.make_regrid(ns.reduce.mean)(isrc, A, case["n"])
    np.savez_compressed(os.path.join(OUT, "apply_golden.npz"), **out)


def golden_coo(ns):
    rng = np.random.default_rng(77)
    size, m = 250, 180
    keep = rng.random(size) < 0.8
    row = np.arange(size, dtype=np.int64)[keep]
    col = rng.integers(0, m, row.size).astype(np.int64)
    src = rng.random((3, m))
    src[1, ::7] = np.nan
    A = ns.sparse.MatrixCOO(np.ones(row.size), row, col, size, m, row.size)
    out = ns.coo_regrid(src, A, size)
    indptr = ns.sparse.MatrixCOO.from_triplet(row, col, np.ones(row.size), n=size, m=m).to_csr().indptr
    np.savez_compressed(
        os.path.join(OUT, "coo_golden.npz"), row=row, col=col, source=src, size=size, out=out, indptr=indptr
    )


def mixed_mesh(seed=5, nx=9, ny=7):
    """jittered lattice, cells kept as quads or split into two triangles"""
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(nx + 1.0), np.arange(ny + 1.0))
    x = gx + rng.uniform(-0.2, 0.2, gx.shape)
    y = gy + rng.uniform(-0.2, 0.2, gy.shape)
    idx = np.arange((nx + 1) * (ny + 1)).reshape(ny + 1, nx + 1)
    faces = []
    for j in range(ny):
        for i in range(nx):
            a, b, c, d = idx[j, i], idx[j, i + 1], idx[j + 1, i + 1], idx[j + 1, i]
            u = rng.random()
            if u < 0.4:
                faces.append([a, b, c, d])
            elif u < 0.7:
                faces.append([a, b, c, -1])
                faces.append([a, c, d, -1])
            else:
                faces.append([a, b, d, -1])
                faces.append([b, c, d, -1])
    return np.column_stack([x.ravel(), y.ravel()]), np.array(faces, dtype=np.int64)


def golden_connectivity(ns):
    xy, faces = mixed_mesh()
    area = ns.connectivity.area(faces, xy[:, 0], xy[:, 1])
    cent = ns.connectivity.centroids(faces, xy[:, 0], xy[:, 1])
    tri = faces[faces[:, 3] == -1][:, :3]
    tri_area = ns.connectivity.area(tri, xy[:, 0], xy[:, 1])
    tri_cent = ns.connectivity.centroids(tri, xy[:, 0], xy[:, 1])
    # a clockwise copy: reference counterclockwise() (connectivity.py:396-405)
    cw = faces.copy()
    for r in range(0, len(cw), 3):
        n = 4 if cw[r, 3] != -1 else 3
        cw[r, :n] = cw[r, :n][::-1]
    ccw = ns.connectivity.counterclockwise(cw, xy)
    np.savez_compressed(
        os.path.join(OUT, "connectivity_golden.npz"),
        xy=xy, faces=faces, area=area, centroids=cent, tri=tri, tri_area=tri_area,
        tri_centroids=tri_cent, cw=cw, ccw=ccw,
    )


def golden_barycentric(ns):
    """Reference pipeline of UnstructuredGrid2d.barycentric (unstructured.py:146-201) with the
    reference's own voronoi.py / connectivity.py / replace_interpolated_weights; the two
    numba_celltree calls in it (compute_barycentric_weights, locate_points) go through the oracle."""
    import oracle
    from xugrid_b200 import synthetic

    cases = {}
    dxy, df = synthetic.disk()
    gx = np.arange(-10.0, 10.0, 0.5) + 0.25
    yy, xx = np.meshgrid(gx, gx, indexing="ij")
    cases["disk"] = (dxy, df, np.column_stack([xx.ravel(), yy.ravel()]))
    mxy, mf = mixed_mesh(seed=8, nx=12, ny=10)
    rng = np.random.default_rng(4)
    cases["mixed"] = (mxy, mf, np.column_stack([rng.uniform(-0.5, 12.5, 3000), rng.uniform(-0.5, 10.5, 3000)]))
    out = {}
    for name, (xy, faces, points) in cases.items():
        conn = ns.connectivity
        node_face = conn.invert_dense_to_sparse(faces)
        edge_node, face_edge = conn.edge_connectivity(faces)
        edge_face = conn.invert_dense(face_edge)
        centroids = conn.centroids(faces, xy[:, 0], xy[:, 1])
        vertices, vfaces, node_to_face_index, node_to_node_map = ns.voronoi.voronoi_topology(
            node_face, xy, centroids, edge_face_connectivity=edge_face, edge_node_connectivity=edge_node,
            add_exterior=True, add_vertices=True, skip_concave=True)
        vfaces = np.ascontiguousarray(vfaces, dtype=np.int64)
        tree = oracle.CellTree2d(vertices, vfaces, -1)
        face_index, weights = tree.compute_barycentric_weights(points)
        ns.replace_interpolated_weights(vertices, vfaces, face_index, weights, node_to_node_map,
                                        len(vertices) - len(node_to_node_map))
        outside = oracle.CellTree2d(xy, faces, -1).locate_points(points) == -1
        weights[outside] = 0
        keep = weights.ravel() > 0
        source_index = node_to_face_index[vfaces[face_index]].ravel()[keep]
        n_points, n_max = weights.shape
        target_index = np.repeat(np.arange(n_points), n_max)[keep]
        w = weights.ravel()[keep]
        order = np.lexsort((w, source_index, target_index))
        out[f"{name}_xy"], out[f"{name}_faces"], out[f"{name}_points"] = xy, faces, points
        out[f"{name}_target"], out[f"{name}_source"], out[f"{name}_weight"] = target_index[order], source_index[order], w[order]
        out[f"{name}_vor_n_vertex"] = len(vertices)
        out[f"{name}_vor_area"] = np.sort(conn.area(vfaces, vertices[:, 0], vertices[:, 1]))
        print("barycentric", name, len(w), "triplets", vfaces.shape)
    np.savez_compressed(os.path.join(OUT, "barycentric_golden.npz"), **out)


def structured_cases():
    """(name, source kwargs, target kwargs) of raster pairs: ascending / descending axes,
    equidistant midpoints, explicit cell sizes, explicit non-contiguous bounds, midpoints
    that fall exactly on source cell bounds, targets sticking out of the source."""
    rng = np.random.default_rng(2024)
    cases = []
    # the reference's own fixtures a, b, c (tests/fixtures/fixture_regridder.py:94-158)
    a = dict(x=[50.0, 100.0, 150.0], y=[150.0, 100.0, 50.0])
    cases.append(("a_b", a, dict(x=[25.0, 75.0, 125.0, 175.0], y=[175.0, 125.0, 75.0, 25.0])))
    cases.append(("a_c", a, dict(x=[40.0, 90.0, 140.0, 190.0], y=[175.0, 125.0, 75.0, 25.0])))
    cases.append(("b_a", dict(x=[25.0, 75.0, 125.0, 175.0], y=[175.0, 125.0, 75.0, 25.0]), a))
    for i in range(6):
        nsx, nsy, ntx, nty = rng.integers(3, 30, 4)
        sx = np.sort(rng.uniform(0, 100, 1))[0] + np.arange(nsx) * rng.uniform(1, 5)
        sy = np.sort(rng.uniform(0, 100, 1))[0] + np.arange(nsy) * rng.uniform(1, 5)
        tx = sx[0] - 3.0 + np.arange(ntx) * (sx[-1] - sx[0] + 6.0) / ntx
        ty = sy[0] + 1.0 + np.arange(nty) * (sy[-1] - sy[0]) / nty
        if i % 2:
            sy, ty = sy[::-1], ty[::-1]
        if i % 3 == 0:
            sx = sx[::-1]
        if i % 3 == 1:
            tx = tx[::-1]
        cases.append((f"rand{i}", dict(x=sx, y=sy), dict(x=tx, y=ty)))
    # commensurate lattices: every other target midpoint lies exactly on a source bound
    cases.append(("ties", dict(x=np.arange(0.5, 8.5), y=np.arange(7.5, 0.0, -1.0)),
                  dict(x=np.arange(0.0, 9.0, 0.5), y=np.arange(8.0, -0.5, -0.5))))
    # explicit, non-equidistant bounds (fixture e, fixture_regridder.py:160-182) and cell sizes
    e_bounds = np.column_stack(([17.5, 42.5, 92.5], [42.5, 92.5, 117.5]))
    cases.append(("a_e", a, dict(x=[30.0, 67.5, 105.0], y=[175.0, 125.0, 75.0, 25.0], xkw=dict(bounds=e_bounds))))
    edges = np.cumsum(np.r_[0.0, rng.uniform(0.5, 3.0, 12)])
    cases.append(("nonequi", dict(x=0.5 * (edges[1:] + edges[:-1]), y=np.arange(10.0, 0.0, -1.0),
                                  xkw=dict(bounds=np.column_stack((edges[:-1], edges[1:])))),
                  dict(x=np.linspace(1.0, 20.0, 17), y=np.linspace(9.7, 0.2, 23))))
    return cases


def golden_structured():
    st = ref_loader.load_structured()
    out = {}
    names = []
    for name, s_kw, t_kw in structured_cases():
        src, tgt = st.grid(**s_kw), st.grid(**t_kw)
        names.append(name)
        for key, kw in (("s", s_kw), ("t", t_kw)):
            out[f"{name}_{key}x"], out[f"{name}_{key}y"] = np.asarray(kw["x"], float), np.asarray(kw["y"], float)
            for ax in ("x", "y"):
                b = (kw.get(ax + "kw") or {}).get("bounds")
                if b is not None:
                    out[f"{name}_{key}{ax}bounds"] = np.asarray(b, float)
        results = {
            "overlap": src.overlap(tgt, relative=False),
            "relative": src.overlap(tgt, relative=True),
            "centroid": src.locate_centroids(tgt, None),
        }
        try:
            results["linear"] = src.linear_weights(tgt)
        except ValueError:
            pass
        for kind, (si, ti, w) in results.items():
            order = np.lexsort((w, si, ti))
            out[f"{name}_{kind}_source"], out[f"{name}_{kind}_target"] = si[order], ti[order]
            out[f"{name}_{kind}_weight"] = w[order]
        print("structured", name, {k: len(v[2]) for k, v in results.items()})
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "structured_golden.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load()
    golden_structured()
    golden_apply(ns)
    golden_coo(ns)
    golden_connectivity(ns)
    golden_barycentric(ns)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
