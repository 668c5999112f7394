"""
``.weights`` <-> ``xarray.Dataset`` in the layout the reference writes
(/root/reference/xugrid/regrid/regridder.py:264-271,298-361), so that weights saved by
xugrid (``regridder.to_dataset().to_netcdf(...)``) load here and the other way round.

xarray is optional: it is not part of the image this library is built in, and nothing
on the compute path needs it.  The regridders' native persistence format is a plain
``dict`` of numpy arrays (regridder.py::BaseRegridder.to_dataset); these two functions
translate between that dict and the reference's dataset:

=========================  ==========================================================
dict key                   dataset variable (reference)
=========================  ==========================================================
``__regrid_data`` ...      1-D variables on a dimension of their own name -- what
                           ``xr.Dataset({name: ndarray})`` makes (regridder.py:266-268);
                           ``__regrid_n/m/nnz`` scalars
``<p>_type``               ``DataArray(-1, attrs={"type": ...})`` (unstructured.py:217-220,
                           structured.py:603-608)
``<p>_node_x/_node_y``     UGRID node coordinates on ``<p>_nNodes``
``<p>_face_node_           ``<p>_face_nodes`` on ``(<p>_nFaces, <p>_nMax_face_nodes)``
connectivity``             with ``cf_role / start_index / _FillValue`` and the mesh
                           topology variable ``<p>`` (xugrid/ugrid/conventions.py:146-177)
``<p>_x/_y/_xbounds/``     coordinates ``<p>_x``, ``<p>_xbounds (<p>_x, <p>_xnbounds)``
``_ybounds``               ... as StructuredGrid1d.to_dataset writes them
                           (structured.py:440-456)
=========================  ==========================================================
"""

from __future__ import annotations

from typing import Dict

import numpy as np

WEIGHT_ARRAYS = ("__regrid_data", "__regrid_indices", "__regrid_indptr", "__regrid_row", "__regrid_col")
WEIGHT_SCALARS = ("__regrid_n", "__regrid_m", "__regrid_nnz")


def _xarray():
    try:
        import xarray as xr
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError(
            "xarray is needed to exchange weights as xarray.Dataset; the dict returned by "
            "regridder.to_dataset() needs no extra package"
        ) from e
    return xr


def topology_attrs(name: str) -> dict:
    # conventions.py:158-177, the entries this layout uses
    return {
        "cf_role": "mesh_topology",
        "long_name": "Topology data of 2D mesh",
        "topology_dimension": 2,
        "node_dimension": f"{name}_nNodes",
        "face_dimension": f"{name}_nFaces",
        "max_face_nodes_dimension": f"{name}_nMax_face_nodes",
        "face_node_connectivity": f"{name}_face_nodes",
        "node_coordinates": f"{name}_node_x {name}_node_y",
    }


def to_xarray(ds: Dict[str, object]):
    """dict from ``regridder.to_dataset()`` -> ``xarray.Dataset`` in the reference's layout."""
    xr = _xarray()
    data_vars, coords = {}, {}
    for key in WEIGHT_ARRAYS:
        if key in ds:
            data_vars[key] = xr.DataArray(np.asarray(ds[key]), dims=(key,))
    for key in WEIGHT_SCALARS:
        data_vars[key] = xr.DataArray(int(ds[key]))
    for p in ("__source", "__target"):
        kind = ds[f"{p}_type"]
        data_vars[f"{p}_type"] = xr.DataArray(-1, attrs={"type": kind})
        if kind == "UnstructuredGrid2d":
            data_vars[p] = xr.DataArray(0, attrs=topology_attrs(p))
            coords[f"{p}_node_x"] = xr.DataArray(np.asarray(ds[f"{p}_node_x"]), dims=(f"{p}_nNodes",),
                                                 attrs={"standard_name": "projection_x_coordinate"})
            coords[f"{p}_node_y"] = xr.DataArray(np.asarray(ds[f"{p}_node_y"]), dims=(f"{p}_nNodes",),
                                                 attrs={"standard_name": "projection_y_coordinate"})
            data_vars[f"{p}_face_nodes"] = xr.DataArray(
                np.asarray(ds[f"{p}_face_node_connectivity"], dtype=np.int64),
                dims=(f"{p}_nFaces", f"{p}_nMax_face_nodes"),
                attrs={"cf_role": "face_node_connectivity", "start_index": 0, "_FillValue": -1},
            )
        elif kind == "StructuredGrid2d":
            for ax in ("x", "y"):
                name = f"{p}_{ax}"
                coords[name] = xr.DataArray(np.asarray(ds[name]), dims=(name,))
                coords[name + "bounds"] = xr.DataArray(np.asarray(ds[name + "bounds"]), dims=(name, name + "nbounds"))
                coords[name + "nbounds"] = xr.DataArray(np.arange(2), dims=(name + "nbounds",))
        else:
            raise ValueError(f"cannot export a {kind} topology")
    return xr.Dataset(data_vars, coords=coords, attrs={"Conventions": "CF-1.9 UGRID-1.0"})


def from_xarray(dataset) -> Dict[str, object]:
    """``xarray.Dataset`` written by the reference (or by ``to_xarray``) -> the dict the
    regridders' ``from_dataset`` / ``from_weights`` read."""
    ds: Dict[str, object] = {}
    for key in WEIGHT_ARRAYS:
        if key in dataset:
            ds[key] = dataset[key].to_numpy()
    for key in WEIGHT_SCALARS:
        ds[key] = int(dataset[key].item())
    for p in ("__source", "__target"):
        if f"{p}_type" not in dataset:
            continue
        kind = dataset[f"{p}_type"].attrs["type"]
        ds[f"{p}_type"] = kind
        if kind == "UnstructuredGrid2d":
            attrs = dataset[p].attrs
            node_x, node_y = attrs["node_coordinates"].split()[:2]
            face_nodes = dataset[attrs["face_node_connectivity"]]
            conn = face_nodes.to_numpy()
            fill = face_nodes.attrs.get("_FillValue", face_nodes.encoding.get("_FillValue", -1))
            start = int(face_nodes.attrs.get("start_index", 0))
            is_fill = (conn == fill) | np.isnan(conn) if conn.dtype.kind == "f" else (conn == fill)
            conn = np.where(is_fill, -1, conn - start if start else conn).astype(np.int64)
            ds[f"{p}_node_x"] = dataset[node_x].to_numpy().astype(np.float64)
            ds[f"{p}_node_y"] = dataset[node_y].to_numpy().astype(np.float64)
            ds[f"{p}_face_node_connectivity"] = conn
        else:
            for ax in ("x", "y"):
                name = f"{p}_{ax}"
                ds[name] = dataset[name].to_numpy()
                ds[name + "bounds"] = dataset[name + "bounds"].to_numpy()
    return ds
