"""Raster -> raster weights (xugrid_b200/structured.py, host logic + torch ops, run on CPU
tensors here) against

* the reference's own known answers, /root/reference/tests/test_regrid/test_structured.py
  :51-205 (overlap 1-D / 2-D), :208-262 (centroids), :265-330 (linear weights);
* golden vectors produced by the reference's structured.py itself
  (oracle/make_golden.py::golden_structured -> tests/golden/structured_golden.npz)."""
import os

import numpy as np
import pytest

from xugrid_b200 import structured
from xugrid_b200.grid import StructuredAxis, StructuredGrid2d

# fixtures a..e of /root/reference/tests/fixtures/fixture_regridder.py:94-182
A_X, A_Y = [50.0, 100.0, 150.0], [150.0, 100.0, 50.0]
B_X, B_Y = [25.0, 75.0, 125.0, 175.0], [175.0, 125.0, 75.0, 25.0]
C_X = [40.0, 90.0, 140.0, 190.0]
D_X = [30.0, 55.0, 80.0, 105.0]
E_X = [30.0, 67.5, 105.0]
E_XBOUNDS = np.column_stack(([17.5, 42.5, 92.5], [42.5, 92.5, 117.5]))


def axis(mid, name="x", **kw):
    return StructuredAxis(np.array(mid), name, **kw)


def assert_expected(actual, expected_source, expected_target, expected_weights):
    # the comparison of test_structured.py:29-48: the set of (target, source) pairs and their weights
    s, t, w = actual
    got = sorted(zip(t.tolist(), s.tolist(), np.round(w, 12).tolist()))
    exp = sorted(zip(expected_target, expected_source, np.round(np.asarray(expected_weights, float), 12).tolist()))
    assert got == exp


def test_overlap_1d_known_answers():
    a, b, b_flipped, e = axis(A_X), axis(B_X), axis(B_Y, "y"), axis(E_X, bounds=E_XBOUNDS)
    assert_expected(structured.overlap_1d(a, b, False), [0, 0, 1, 1, 2, 2], [0, 1, 1, 2, 2, 3], [25] * 6)
    assert_expected(structured.overlap_1d(a, b_flipped, False), [0, 0, 1, 1, 2, 2], [2, 3, 1, 2, 0, 1], [25] * 6)
    assert_expected(structured.overlap_1d(a, e, False), [0, 0, 1, 1], [0, 1, 1, 2], [17.5, 32.5, 17.5, 25.0])
    assert_expected(structured.overlap_1d(a, e, True), [0, 0, 1, 1], [0, 1, 1, 2],
                    [17.5 / 50.0, 32.5 / 50.0, 17.5 / 50.0, 25.0 / 50.0])


def test_locate_centroids_1d_known_answers():
    a, b, b_flipped, e = axis(A_X), axis(B_X), axis(B_Y, "y"), axis(E_X, bounds=E_XBOUNDS)
    assert_expected(structured.locate_1d(a, b), [0, 1], [1, 2], [1, 1])
    assert_expected(structured.locate_1d(a, b_flipped), [0, 1], [2, 1], [1, 1])
    assert_expected(structured.locate_1d(a, e), [0, 0, 1], [0, 1, 2], [1, 1, 1])


def test_linear_weights_1d_known_answers():
    a, b, b_flipped, c, d = axis(A_X), axis(B_X), axis(B_Y, "y"), axis(C_X), axis(D_X, size=25.0)
    assert_expected(structured.linear_1d(a, b), [0, 1, 1, 2], [1, 1, 2, 2], [0.5] * 4)
    assert_expected(structured.linear_1d(a, b_flipped), [2, 1, 1, 0], [1, 1, 2, 2], [0.5] * 4)
    assert_expected(structured.linear_1d(a, c), [0, 0, 1, 0, 2, 1], [0, 0, 1, 1, 2, 2], [0.0, 1.0, 0.8, 0.2, 0.8, 0.2])
    assert_expected(structured.linear_1d(a, d), [0, 0, 0, 1, 1, 0, 1, 2], [0, 0, 1, 1, 2, 2, 3, 3],
                    [0.0, 1.0, 0.9, 0.1, 0.6, 0.4, 0.9, 0.1])
    with pytest.raises(ValueError, match="At least two points"):
        structured.linear_1d(axis([1.0], size=1.0), b)


def _triplets(A):
    A = A.to_coo() if hasattr(A, "indptr") else A
    data, row, col = A.to_numpy()
    return col, row, data


def test_overlap_2d_known_answer():
    # test_structured.py:108-205: 36 pairs, every weight 625.0
    a, b = StructuredGrid2d(A_X, A_Y), StructuredGrid2d(B_X, B_Y)
    expected_source = np.repeat(np.arange(9), 4)
    expected_target = [0, 4, 5, 1, 2, 6, 5, 1, 2, 3, 7, 6, 8, 9, 5, 4, 9, 5, 10, 6, 10, 11, 7, 6, 9, 8, 12, 13,
                       10, 14, 13, 9, 10, 11, 14, 15]
    A = structured.overlap(a, b, relative=False, device="cpu")
    assert_expected(_triplets(A), expected_source.tolist(), expected_target, np.full(36, 625.0))
    assert A.n == 16 and A.m == 9 and A.nnz == 36
    assert np.array_equal(a.area, np.full((3, 3), 2500.0))


def test_locate_centroids_2d_known_answer():
    # test_structured.py:249-262
    a, b = StructuredGrid2d(A_X, A_Y), StructuredGrid2d(B_X, B_Y)
    A = structured.locate_centroids(a, b, device="cpu")
    assert_expected(_triplets(A), [0, 1, 3, 4], [5, 6, 9, 10], np.ones(4))


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "structured_golden.npz"))


def _grid(g, name, key):
    kw = {}
    for ax in ("x", "y"):
        b = f"{name}_{key}{ax}bounds"
        if b in g.files:
            kw[ax + "bounds"] = g[b]
    return StructuredGrid2d(g[f"{name}_{key}x"], g[f"{name}_{key}y"], **kw)


def test_against_reference_golden(golden):
    fn = {
        "overlap": lambda s, t: structured.overlap(s, t, False, device="cpu"),
        "relative": lambda s, t: structured.overlap(s, t, True, device="cpu"),
        "centroid": lambda s, t: structured.locate_centroids(s, t, device="cpu"),
        "linear": lambda s, t: structured.linear_weights(s, t, device="cpu"),
    }
    checked = 0
    for name in golden["names"]:
        src, tgt = _grid(golden, name, "s"), _grid(golden, name, "t")
        for kind, f in fn.items():
            if f"{name}_{kind}_weight" not in golden.files:
                continue
            A = f(src, tgt)
            if hasattr(A, "indptr"):
                ip = A.indptr.numpy()
                assert ip[0] == 0 and ip[-1] == A.nnz and np.all(np.diff(ip) >= 0)
            s, t, w = _triplets(A)
            order = np.lexsort((w, s, t))
            assert np.array_equal(t[order], golden[f"{name}_{kind}_target"]), (name, kind)
            assert np.array_equal(s[order], golden[f"{name}_{kind}_source"]), (name, kind)
            # the same products of the same 1-D weights: bit-identical
            assert np.array_equal(w[order], golden[f"{name}_{kind}_weight"]), (name, kind)
            checked += 1
    assert checked >= 40
