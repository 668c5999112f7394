"""CPU tests: the C-ABI library loads and exports every declared symbol, the
product refuses to run without a GPU, and the host-side logic behaves like the
reference's."""
import os
import re

import numpy as np
import pytest
import torch

import xugrid_b200 as xb
from xugrid_b200 import _native, build, grid, regridder, synthetic
from xugrid_b200.distributed import max_shard, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "xg_regrid.h")).read()
    declared = set(re.findall(r"\b(xg_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.xg_abi_version() == 1


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    xy, faces = synthetic.disk()
    with pytest.raises(_native.XgError):
        xb.CellTree2d(xy, faces)
    g = xb.Ugrid2d(xy[:, 0], xy[:, 1], -1, faces)
    with pytest.raises(_native.XgError):
        xb.OverlapRegridder(g, g)


def test_ugrid2d_normalises_connectivity():
    # ugrid2d.py:103-109
    xy, faces = synthetic.jittered_lattice(3, 2, 0, "mixed")
    f1 = faces.copy()
    f1[f1 >= 0] += 1
    f1[f1 < 0] = -999
    g = xb.Ugrid2d(xy[:, 0], xy[:, 1], -999, f1, start_index=1)
    assert np.array_equal(g.face_node_connectivity, faces)
    assert g.n_face == len(faces) and g.n_node == len(xy)
    with pytest.raises(TypeError):
        xb.Ugrid2d(xy[:, 0], xy[:, 1], -1, [[0, 1, 2]])


def test_from_structured_bounds_is_ccw_quads():
    # ugrid2d.py:1894-1912,1973-2034
    xb_ = np.array([[0.0, 1.0], [1.0, 2.5], [2.5, 3.0]])
    yb_ = np.array([[8.0, 10.0], [7.0, 8.0]])  # rows descending, each row [lower, upper]
    g = xb.Ugrid2d.from_structured_bounds(xb_, yb_)
    assert g.n_face == 6 and g.face_node_connectivity.shape == (6, 4)
    xyc = g.node_coordinates[g.face_node_connectivity]
    d1 = xyc[:, 1] - xyc[:, 0]
    d2 = xyc[:, 2] - xyc[:, 1]
    assert ((d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]) > 0).all()
    with pytest.raises(ValueError):
        grid.bounds1d_to_vertices(np.array([[0.0, 1.0], [3.0, 2.0], [1.0, 1.5]]))


def test_structured_grid_bounds_inference():
    # structured.py:24-100
    s = xb.StructuredGrid2d(x=np.array([0.5, 1.5, 2.5]), y=np.array([3.5, 2.5]))
    assert s.shape == (2, 3) and s.size == 6 and s.dims == ("y", "x")
    assert np.allclose(s.xbounds.bounds, [[0, 1], [1, 2], [2, 3]])
    assert s.ybounds.flipped and np.allclose(s.ybounds.directional_bounds, [[3, 4], [2, 3]])
    with pytest.raises(ValueError):
        xb.StructuredGrid2d(x=np.array([0.0, 1.0, 3.0]), y=np.array([0.0, 1.0]))
    with pytest.raises(ValueError):
        xb.StructuredGrid2d(x=np.array([0.0, 2.0, 1.0]), y=np.array([0.0, 1.0]))


def test_percentile_method_validation():
    # reduce.py:241-248, tests/test_regrid/test_regridder.py:282-293
    m = xb.OverlapRegridder.create_percentile_method(33.3)
    assert m.percentile == 33.3
    with pytest.raises(ValueError):
        xb.create_percentile_method(-1.0)
    with pytest.raises(ValueError):
        xb.create_percentile_method(100.5)


def test_setup_grid_type_errors():
    # regridder.py:72-80, unstructured.py:77-81
    with pytest.raises(TypeError):
        regridder.setup_grid(np.zeros(3))
    with pytest.raises(TypeError):
        regridder.UnstructuredGrid2d("nope")


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges[:-1], ranges[1:]):
                assert b == c
            assert max(b - a for a, b in ranges) == max_shard(n, world) or n == 0
            assert all(a == min(n, r * max_shard(n, world)) for r, (a, _) in enumerate(ranges))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_recorded_bench_lines_follow_the_contract():
    """the JSON lines bench.py printed on the B200 (kept under profiles/) carry every key the
    bench contract names, for our arm and for the reference arm"""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ours = json.load(open(os.path.join(root, "profiles", "r1_bench_n1.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline",
                "cpu_baseline", "clocks"):
        assert key in ours, key
    assert ours["n_gpus"] == 1 and ours["warmup"] >= 3 and ours["dtype"] == "f64" and ours["vs_baseline"] is None
    assert "workload" in ours["config"] and "model" not in ours["config"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in ours["e2e"], key
    assert ours["e2e"]["h2d_bytes_per_step"] > 0 and ours["e2e"]["value"] < ours["value"]
    roof = ours["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roof, key
    assert roof["bound"] == "hbm" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in ours["cpu_baseline"], key
    assert ours["gpu_launches"] > 0
    assert not set(ours["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ref = json.load(open(os.path.join(root, "profiles", "r1_bench_reference_arm.json")))
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    assert ref["cpu_baseline"]["kind"] in ("port", "reference")


def test_recorded_round2_bench_lines():
    """round 2: the final N = 1 line (with the CPU baseline on the SAME workload) and the N = 8
    line, whose N-rank results were checked bit for bit against one rank inside the run"""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    one = json.load(open(os.path.join(root, "profiles", "r2_bench_n1_v11.json")))
    eight = json.load(open(os.path.join(root, "profiles", "r2_bench_n8_v14.json")))
    for line in (one, eight):
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
            assert key in line, key
        assert line["scaling"] == "strong" and line["warmup"] >= 3 and line["config"]["n_target_faces"] == 10003852
        roof = line["roofline"]
        assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12 and "binding_roof" in roof
        assert line["e2e"]["value"] < line["value"] and line["e2e"]["h2d_bytes_per_step"] > 0
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # measured DRAM traffic of the dominant kernel next to the byte model
    assert one["roofline"]["traffic"] and 0 < one["roofline"]["frac_measured_dram"] < 1
    # the CPU baseline ran the workload the line is labelled with
    assert "10.0M -> 10.0M" in one["cpu_baseline"]["sample"] and one["cpu_baseline"]["cores"] >= 1
    assert eight["n_gpus"] == 8 and eight["value"] > 4.5 * one["value"]
    chk = eight["multi_gpu_check"]
    assert chk["gathered_output_bit_identical_to_single_rank"] is True and chk["e2e_output_bit_identical"] is True
    assert eight["e2e"]["d2h_bytes_per_step"] * 8 >= one["e2e"]["d2h_bytes_per_step"]  # a rank brings its own columns down
    # the reference arm prints the SAME config as our arm (round 1 ran 4 M faces under a 10 M label)
    ref = json.load(open(os.path.join(root, "profiles", "r2_bench_reference_arm_build_container.json")))
    assert ref["impl"] == "reference" and ref["config"] == one["config"] and ref["metric"] == one["metric"]
    assert ref["cpu_baseline"]["cores"] >= 1 and ref["e2e"]["h2d_bytes_per_step"] == 0


def test_ugrid1d_and_network1d_host_logic():
    # the part of xugrid.Ugrid1d / regrid/network.py:4-38 the NetworkGridder touches
    from xugrid_b200.regridder import Network1d

    node_xy = np.array([[0.0, 0.0], [3.0, 4.0], [3.0, 0.0]])
    net = xb.Ugrid1d(node_xy[:, 0], node_xy[:, 1], -1, np.array([[0, 1], [1, 2]]))
    assert net.n_edge == 2 and net.n_node == 3 and net.edge_dimension == net.core_dimension
    assert np.array_equal(net.edge_node_coordinates, node_xy[[[0, 1], [1, 2]]])
    assert np.allclose(net.edge_length, [5.0, 4.0])
    wrapped = Network1d(net)
    assert wrapped.size == 2 and wrapped.shape == (2,) and wrapped.ndim == 1
    data = xb.UgridDataArray(np.array([[1.0, 2.0]]), net)
    assert Network1d(data).ugrid_topology is net and data.dims[-1] == net.edge_dimension
    with pytest.raises(TypeError):
        Network1d(np.zeros(3))
    with pytest.raises(ValueError):
        xb.Ugrid1d(node_xy[:, 0], node_xy[:, 1], -1, np.array([[0, 1, 2]]))
    with pytest.raises(ValueError):
        xb.UgridDataArray(np.zeros(3), net)


def test_pending_csr_settles_or_rebuilds():
    """sparse.PendingCSR (weights whose overlap jobs were enqueued without waiting): exact views
    once the jobs report complete, the rebuild closure when one of them does not"""
    import torch

    from xugrid_b200 import sparse

    class Job:
        def __init__(self, nnz, end, ok):
            self._r, self.phase_ms, self.calls = (nnz, end, ok), {"pack": 0.0}, 0

        def result(self):
            self.calls += 1
            return self._r

    data = torch.arange(10, dtype=torch.float64)
    indices = torch.arange(10, dtype=torch.int32)
    indptr = torch.tensor([0, 2, 5, 7], dtype=torch.int64)
    jobs = [Job(5, 5, True), Job(2, 7, True)]
    rebuilt = []
    p = sparse.PendingCSR(data, indices, indptr, 3, 9, jobs, lambda: rebuilt.append(1))
    A = p.for_apply()
    assert A.nnz == -1 and A.n == 3 and A.m == 9 and A.data is data
    assert p.complete() and p.complete()
    assert [j.calls for j in jobs] == [1, 1]  # asked once
    W = p.finalize()
    assert W.nnz == 7 and W.data.numel() == 7 and W.indices.numel() == 7 and W.indptr is indptr and not rebuilt
    assert p.finalize() is W and len(p.job_phase_ms) == 2
    # one incomplete job: the waiting path is asked for the weights instead
    marker = sparse.MatrixCSR(data[:1], indices[:1], indptr[:2], 1, 9, 1)
    q = sparse.PendingCSR(data, indices, indptr, 3, 9, [Job(5, 5, True), Job(0, 5, False)], lambda: marker)
    assert not q.complete()
    assert q.finalize() is marker
