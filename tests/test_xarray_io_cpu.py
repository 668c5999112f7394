"""xugrid_b200/xarray_io.py without xarray installed: a minimal stand-in for the two
xarray classes records what the adapter builds, so the variable names, dimensions and
attributes of the reference's weights dataset
(/root/reference/xugrid/regrid/regridder.py:264-271,298-361; unstructured.py:217-220;
structured.py:440-456,603-608; xugrid/ugrid/conventions.py:146-177) are checked at the
dict level.  With the real xarray the same code paths run unchanged."""
import sys
import types

import numpy as np
import pytest


class FakeDataArray:
    def __init__(self, data, dims=(), attrs=None, coords=None):
        self.data = np.asarray(data)
        self.dims = tuple(dims)
        assert len(self.dims) == self.data.ndim, (self.dims, self.data.shape)
        self.attrs = dict(attrs or {})
        self.encoding = {}

    def to_numpy(self):
        return self.data

    def item(self):
        return self.data.item()


class FakeDataset:
    def __init__(self, data_vars, coords=None, attrs=None):
        self.data_vars = dict(data_vars)
        self.coords = dict(coords or {})
        self.attrs = dict(attrs or {})

    def __getitem__(self, k):
        return self.data_vars[k] if k in self.data_vars else self.coords[k]

    def __contains__(self, k):
        return k in self.data_vars or k in self.coords


@pytest.fixture
def fake_xarray(monkeypatch):
    mod = types.ModuleType("xarray")
    mod.DataArray, mod.Dataset = FakeDataArray, FakeDataset
    monkeypatch.setitem(sys.modules, "xarray", mod)
    return mod


def weights_dict():
    return {
        "__regrid_data": np.array([0.5, 0.25, 1.0]), "__regrid_indices": np.array([0, 1, 1]),
        "__regrid_indptr": np.array([0, 2, 3]), "__regrid_n": 2, "__regrid_m": 2, "__regrid_nnz": 3,
        "__source_node_x": np.array([0.0, 1.0, 0.0, 1.0]), "__source_node_y": np.array([0.0, 0.0, 1.0, 1.0]),
        "__source_face_node_connectivity": np.array([[0, 1, 3], [0, 3, 2]]), "__source_type": "UnstructuredGrid2d",
        "__target_x": np.array([0.25, 0.75]), "__target_y": np.array([0.5]),
        "__target_xbounds": np.array([[0.0, 0.5], [0.5, 1.0]]), "__target_ybounds": np.array([[0.0, 1.0]]),
        "__target_type": "StructuredGrid2d",
    }


def test_dataset_layout_is_the_references(fake_xarray):
    from xugrid_b200 import xarray_io

    d = weights_dict()
    ds = xarray_io.to_xarray(d)
    # regridder.py:266-268: 1-D weight variables sit on a dimension of their own name
    for key in ("__regrid_data", "__regrid_indices", "__regrid_indptr"):
        assert ds[key].dims == (key,)
    for key in ("__regrid_n", "__regrid_m", "__regrid_nnz"):
        assert ds[key].dims == () and ds[key].item() == d[key]
    # unstructured.py:217-220 / structured.py:603-608
    assert ds["__source_type"].attrs == {"type": "UnstructuredGrid2d"} and ds["__source_type"].item() == -1
    assert ds["__target_type"].attrs == {"type": "StructuredGrid2d"}
    # UGRID topology (conventions.py:158-177)
    topo = ds["__source"].attrs
    assert topo["cf_role"] == "mesh_topology" and topo["topology_dimension"] == 2
    assert topo["node_coordinates"] == "__source_node_x __source_node_y"
    assert topo["face_node_connectivity"] == "__source_face_nodes"
    fn = ds["__source_face_nodes"]
    assert fn.dims == ("__source_nFaces", "__source_nMax_face_nodes")
    assert fn.attrs == {"cf_role": "face_node_connectivity", "start_index": 0, "_FillValue": -1}
    assert ds["__source_node_x"].dims == ("__source_nNodes",)
    # structured.py:440-456
    assert ds["__target_xbounds"].dims == ("__target_x", "__target_xnbounds")
    assert ds["__target_y"].dims == ("__target_y",)
    assert ds.attrs["Conventions"] == "CF-1.9 UGRID-1.0"


def test_round_trip(fake_xarray):
    from xugrid_b200 import xarray_io

    d = weights_dict()
    back = xarray_io.from_xarray(xarray_io.to_xarray(d))
    assert set(back) == set(d)
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(back[k], v), k
        else:
            assert back[k] == v, k


def test_reads_one_based_connectivity_with_another_fill_value(fake_xarray):
    from xugrid_b200 import xarray_io

    ds = xarray_io.to_xarray(weights_dict())
    fn = ds["__source_face_nodes"]
    fn.data = np.array([[1, 2, 4, -999], [1, 4, 3, -999]])
    fn.attrs.update(start_index=1, _FillValue=-999)
    back = xarray_io.from_xarray(ds)
    assert np.array_equal(back["__source_face_node_connectivity"], [[0, 1, 3, -1], [0, 3, 2, -1]])


def test_without_xarray_the_error_says_so(monkeypatch):
    from xugrid_b200 import xarray_io

    monkeypatch.setitem(sys.modules, "xarray", None)
    with pytest.raises(ImportError, match="xarray is needed"):
        xarray_io.to_xarray(weights_dict())
