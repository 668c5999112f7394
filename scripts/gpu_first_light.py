"""First-light check on a B200: parity against the oracle + rough timings."""
import sys
import time

import numpy as np
import torch

import oracle
import xugrid_b200 as xb
from xugrid_b200 import sparse, synthetic


def canon(ti, si, a):
    o = np.lexsort((si, ti))
    return ti[o], si[o], a[o]


def check_overlap(sxy, sf, txy, tf, label):
    tree = xb.CellTree2d(sxy, sf)
    ti, si, a = tree.intersect_faces(txy, tf)
    ot = oracle.CellTree2d(sxy, sf)
    oti, osi, oa = ot.intersect_faces(txy, tf)
    A, B = canon(ti, si, a), canon(oti, osi, oa)
    same_pairs = A[0].shape == B[0].shape and np.array_equal(A[0], B[0]) and np.array_equal(A[1], B[1])
    print(f"[{label}] nnz gpu={len(a)} oracle={len(oa)} pairs_equal={same_pairs}", end=" ")
    if same_pairs:
        print("area bit-equal:", np.array_equal(A[2], B[2]), "max rel:", np.max(np.abs(A[2] - B[2]) / B[2]) if len(a) else 0)
    else:
        sa = set(zip(A[0].tolist(), A[1].tolist()))
        sb = set(zip(B[0].tolist(), B[1].tolist()))
        print("only gpu:", len(sa - sb), "only oracle:", len(sb - sa))
    print("   stats:", tree.last_stats)
    return same_pairs


def main():
    print(torch.cuda.get_device_name(0))
    ok = True
    # tiny 4-face mesh
    V = np.array([[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1], [1, 2]], float)
    F = np.array([[0, 1, 4, 3], [1, 2, 5, 4], [3, 4, 6, -1], [4, 5, 6, -1]])
    ok &= check_overlap(V, F, V, F, "self 4-face")
    t = xb.CellTree2d(V, F)
    C = np.array([[0.5, 0.5], [1.5, 0.5], [2 / 3, 4 / 3], [4 / 3, 4 / 3]])
    print("locate", t.locate_points(C), t.locate_points(np.array([[-0.01, 1.0], [-0.01, 0.5]]), 0.011))
    xy = np.array([[0, 0], [0.5, 0.5], [1.5, 0.5], [0.5, 1.5], [2, 2]], float)
    print("bary", t.compute_barycentric_weights(xy))
    xy[:, 0] -= 0.01
    print("bary tol", t.compute_barycentric_weights(xy, tolerance=0.01))

    for n in (300, 3000, 50000):
        sxy, sf = synthetic.delaunay(n, 1)
        txy, tf = synthetic.delaunay(n, 2)
        ok &= check_overlap(sxy, sf, txy, tf, f"delaunay {n}")
    sxy, sf = synthetic.jittered_lattice(40, 30, 3, "mixed")
    txy, tf = synthetic.jittered_lattice(37, 33, 4, "tri")
    ok &= check_overlap(sxy, sf, txy, tf, "mixed->tri")
    ok &= check_overlap(txy, tf, sxy, sf, "tri->mixed")
    ok &= check_overlap(sxy, sf, sxy, sf, "mixed self")

    # locate parity
    sxy, sf = synthetic.delaunay(20000, 5)
    pts = np.random.default_rng(6).random((100000, 2)) * 1.1 - 0.05
    g = xb.CellTree2d(sxy, sf).locate_points(pts)
    o = oracle.CellTree2d(sxy, sf).locate_points(pts)
    print("locate equal:", np.array_equal(g, o), "mismatch:", int((g != o).sum()), "outside:", int((o < 0).sum()))

    # apply parity vs golden
    gd = np.load("tests/golden/apply_golden.npz")
    A = sparse.MatrixCSR.from_arrays(gd["data"], gd["indices"], gd["indptr"], int(gd["n"]), int(gd["m"]))
    src = torch.from_numpy(gd["source"]).cuda()
    for name, (code, p) in sparse.METHOD_CODES.items():
        out = sparse.csr_apply(code, p, src, A, A.n).cpu().numpy()
        ref = gd["out_" + name]
        eq = np.array_equal(out, ref, equal_nan=True)
        close = np.allclose(out, ref, rtol=1e-10, atol=0, equal_nan=True)
        print(f"apply {name:26s} bit-equal={eq} close={close}")
        ok &= close

    # timing 1M -> 1M jittered lattice
    for n_side in (707, 2236):
        sxy, sf = synthetic.jittered_lattice(n_side, n_side, 10, "tri")
        txy, tf = synthetic.jittered_lattice(n_side + 3, n_side - 2, 11, "tri")
        dsxy, dsf = torch.from_numpy(sxy).cuda(), torch.from_numpy(sf).cuda()
        dtxy, dtf = torch.from_numpy(txy).cuda(), torch.from_numpy(tf).cuda()
        vals = torch.rand(1, sf.shape[0], dtype=torch.float64, device="cuda")
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tree = xb.CellTree2d(dsxy, dsf)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            indptr, indices, area = tree.intersect_faces_csr(dtxy, dtf)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            A = sparse.MatrixCSR(area, indices, indptr, tf.shape[0], sf.shape[0], int(area.numel()))
            out = sparse.csr_apply(0, 0.0, vals, A, A.n)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            print(f"n_t={tf.shape[0]} index {1e3*(t1-t0):.2f} ms overlap {1e3*(t2-t1):.2f} ms apply {1e3*(t3-t2):.2f} ms "
                  f"-> {tf.shape[0]/(t3-t0):.3e} faces/s nnz={area.numel()} stats={tree.last_stats}")
        ta = oracle.face_area(txy, tf)
        rs = torch.zeros(tf.shape[0], dtype=torch.float64, device="cuda")
        rows = torch.repeat_interleave(torch.arange(tf.shape[0], device="cuda"), indptr[1:] - indptr[:-1])
        rs.index_add_(0, rows, area)
        rel = (rs.cpu().numpy() - ta) / ta
        print("   row-sum vs target area: median rel", np.median(np.abs(rel)), "frac within 1e-9:", np.mean(np.abs(rel) < 1e-9))
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
