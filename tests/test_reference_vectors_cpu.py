"""Pins the ORACLE's intersect_faces / locate_points (the restatement of the third-party
numba_celltree stage, oracle/xg_oracle.c) to the overlap and centroid vectors the reference's
own tests hold: the structured fixtures a..e regridded as QUAD MESHES
(StructuredGrid2d.convert_to(UnstructuredGrid2d), structured.py:489-501) must reproduce
/root/reference/tests/test_regrid/test_structured.py:51-205,249-262 and
tests/fixtures/fixture_regridder.py:240-322.  tests/test_gpu_reference_vectors.py asserts the
same vectors on the CUDA path."""
import numpy as np
import pytest

import oracle
from tests import cases
from xugrid_b200.grid import StructuredGrid2d


def quad_mesh(**kw):
    g = StructuredGrid2d(**kw).to_ugrid2d()
    return g.node_coordinates, g.face_node_connectivity


def overlap(src_kw, tgt_kw):
    sxy, sf = quad_mesh(**src_kw)
    txy, tf = quad_mesh(**tgt_kw)
    ti, si, area = oracle.CellTree2d(sxy, sf).intersect_faces(txy, tf)
    return si, ti, area, oracle.face_area(sxy, sf)


def test_overlap_2d_is_the_reference_answer():
    si, ti, area, _ = overlap(cases.FIX_A, cases.FIX_B)
    assert cases.pair_set(si, ti, area) == cases.pair_set(cases.OVERLAP_2D_SOURCE, cases.OVERLAP_2D_TARGET,
                                                          np.full(36, 625.0))


@pytest.mark.parametrize("name", list(cases.OVERLAP_1D))
def test_overlap_1d_as_quad_strips(name):
    strip = cases.strip_axis_y()
    src = dict(x=cases.FIX_A["x"], **strip)
    if name == "a_b":
        tgt = dict(x=cases.FIX_B["x"], **strip)
    elif name == "a_b_flipped":
        tgt = dict(x=cases.FIX_B["y"], **strip)   # descending midpoints
    else:
        tgt = dict(x=cases.FIX_E_X, xbounds=cases.FIX_E_XBOUNDS, **strip)
    si, ti, area, s_area = overlap(src, tgt)
    if name.endswith("relative"):
        area = area / s_area[si]   # unstructured.py:133-134
    es, et, ew = cases.OVERLAP_1D[name]
    assert cases.pair_set(si, ti, area) == cases.pair_set(es, et, ew)


def test_centroids_2d_is_the_reference_answer():
    sxy, sf = quad_mesh(**cases.FIX_A)
    txy, tf = quad_mesh(**cases.FIX_B)
    tree = oracle.CellTree2d(sxy, sf)
    cent_all = oracle.face_centroids(txy, tf)
    idx = tree.locate_points(cent_all)
    # interior targets: every centroid of b sits on a corner shared by FOUR cells of a.  Which
    # of the four a tree walk reports first depends on the tree layout (unpinned, see the
    # oracle's header); the reference's pairs (test_structured.py:249-262) are the smallest
    # source index among them, which is the rule the CUDA path implements
    for t, s in zip(cases.CENTROID_2D_TARGET, cases.CENTROID_2D_SOURCE):
        qualifying = tree.locate_point_all(*cent_all[t])
        assert idx[t] in qualifying and qualifying.min() == s and len(qualifying) == 4
    # the other twelve centroids lie ON the rim of a or outside it.  The reference's separable
    # path excludes the rim (structured.py:149-153); the tree path it takes for meshes counts a
    # point on an edge as inside (tolerance): nothing outside the rim may be located
    rim_or_outside = np.setdiff1d(np.arange(16), cases.CENTROID_2D_TARGET)
    cent = oracle.face_centroids(txy, tf)[rim_or_outside]
    strictly_outside = (cent[:, 0] < 25) | (cent[:, 0] > 175) | (cent[:, 1] < 25) | (cent[:, 1] > 175)
    assert np.all(idx[rim_or_outside][strictly_outside] == -1)


def test_expected_results_overlap_through_the_oracle():
    # fixture_regridder.py:273-322 with mean over the quad-mesh overlap weights
    si, ti, area, _ = overlap(cases.FIX_A, cases.FIX_B)
    order = np.lexsort((si, ti))
    si, ti, area = si[order], ti[order], area[order]
    indptr = oracle.coo_to_csr(ti, 16)
    out = oracle.csr_apply("mean", cases.FIX_A_DATA.reshape(1, 9), area, si, indptr, 16)
    assert np.array_equal(out.reshape(4, 4), cases.EXPECTED_OVERLAP)
