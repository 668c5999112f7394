"""
Loads the pieces of the UNMODIFIED reference that can run in the build
container (numba present; xarray / numba_celltree absent) so that
``oracle/make_golden.py`` can record their outputs as golden vectors.

Test infrastructure only.  Nothing here runs on the GPU box: /root/reference
does not exist there, which is why the outputs are committed under
``tests/golden/`` instead.

Recipe: stub parent packages ``xugrid``, ``xugrid.core``,
``xugrid.regrid``, ``xugrid.ugrid`` in ``sys.modules`` and load the individual
files by path.  ``make_regrid`` (regridder.py:34-69) and
``replace_interpolated_weights`` (unstructured.py:17-57) live in modules that
import xarray at top level, so their source text is extracted at run time into
a temporary file (never committed) and imported from there.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("XUGRID_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "xugrid"))


def _stub(name: str):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    return sys.modules[name]


def _load(dotted: str, relpath: str):
    if dotted in sys.modules and getattr(sys.modules[dotted], "__file__", None):
        return sys.modules[dotted]
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(dotted, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    parent, _, child = dotted.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, mod)
    return mod


_cache = {}


def load():
    """Returns a namespace with the reference's own modules/functions."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "xg_numba_cache"))
    for name in ("xugrid", "xugrid.core", "xugrid.regrid", "xugrid.ugrid"):
        _stub(name)
    ns = types.SimpleNamespace()
    ns.constants = _load("xugrid.constants", "xugrid/constants.py")
    ns.sparse = _load("xugrid.core.sparse", "xugrid/core/sparse.py")
    ns.nanpercentile = _load("xugrid.regrid.nanpercentile", "xugrid/regrid/nanpercentile.py")
    ns.reduce = _load("xugrid.regrid.reduce", "xugrid/regrid/reduce.py")
    ns.connectivity = _load("xugrid.ugrid.connectivity", "xugrid/ugrid/connectivity.py")
    try:
        ns.voronoi = _load("xugrid.ugrid.voronoi", "xugrid/ugrid/voronoi.py")
    except Exception as e:  # pragma: no cover - optional
        ns.voronoi = None
        ns.voronoi_error = e

    # Extract make_regrid / CentroidLocatorRegridder._regrid / replace_interpolated_weights
    def _lines(rel, a, b):
        with open(os.path.join(REFERENCE_ROOT, rel)) as f:
            return "".join(f.readlines()[a - 1 : b])

    tmpdir = tempfile.mkdtemp(prefix="xg_ref_extract_")
    header = (
        "import numba\nimport numpy as np\n"
        "from xugrid.core.sparse import MatrixCOO, MatrixCSR, row_slice\n"
        "FloatArray = np.ndarray\n\n"
    )
    body = _lines("xugrid/regrid/regridder.py", 34, 69)
    coo = _lines("xugrid/regrid/regridder.py", 400, 409)
    # de-indent the static method and drop the @staticmethod line
    coo_lines = [ln[4:] if ln.startswith("    ") else ln for ln in coo.splitlines(True)]
    coo_src = "".join(ln for ln in coo_lines if "@staticmethod" not in ln).replace(
        "def _regrid(", "def coo_regrid("
    )
    riw = _lines("xugrid/regrid/unstructured.py", 17, 57)
    path = os.path.join(tmpdir, "xg_ref_extracted.py")
    with open(path, "w") as f:
        f.write(header + body + "\n\n" + coo_src + "\n\n" + riw + "\n")
    spec = importlib.util.spec_from_file_location("xg_ref_extracted", path)
    ext = importlib.util.module_from_spec(spec)
    sys.modules["xg_ref_extracted"] = ext
    spec.loader.exec_module(ext)
    ns.make_regrid = ext.make_regrid
    ns.coo_regrid = ext.coo_regrid
    ns.replace_interpolated_weights = ext.replace_interpolated_weights
    _cache["ns"] = ns
    return ns


def load_structured():
    """The reference's separable raster path (xugrid/regrid/structured.py,
    overlap_1d.py, utils.py).  structured.py imports xarray, UnstructuredGrid2d and
    Ugrid2d at module level but its 1-D / 2-D weight methods only touch numpy
    attributes, so those three imports are stubbed and grids are assembled with
    ``axis()`` / ``grid()`` below instead of the xarray-reading constructors
    (structured.py:33-83 is restated in ``axis``: same midpoints / bounds /
    flipped / side)."""
    if "st" in _cache:
        return _cache["st"]
    import numpy as np

    load()
    if "xarray" not in sys.modules:
        xr = types.ModuleType("xarray")
        xr.DataArray = type("DataArray", (), {})
        xr.Dataset = type("Dataset", (), {})
        sys.modules["xarray"] = xr
    if "xugrid.regrid.unstructured" not in sys.modules:
        un = types.ModuleType("xugrid.regrid.unstructured")
        un.UnstructuredGrid2d = type("UnstructuredGrid2d", (), {})
        sys.modules["xugrid.regrid.unstructured"] = un
    if "xugrid.ugrid.ugrid2d" not in sys.modules:
        ug = types.ModuleType("xugrid.ugrid.ugrid2d")
        ug.Ugrid2d = type("Ugrid2d", (), {})
        sys.modules["xugrid.ugrid.ugrid2d"] = ug
    _load("xugrid.regrid.utils", "xugrid/regrid/utils.py")
    _load("xugrid.regrid.overlap_1d", "xugrid/regrid/overlap_1d.py")
    st = _load("xugrid.regrid.structured", "xugrid/regrid/structured.py")

    def axis(index, name, bounds=None, size=None):
        g = object.__new__(st.StructuredGrid1d)
        index = np.asarray(index, dtype=float)
        if index.size > 1 and (np.diff(index) < 0).all():
            mid, flipped, side = index[::-1], True, "right"
        else:
            mid, flipped, side = index, False, "left"
        if bounds is None:
            if size is None:
                size = np.full_like(mid, np.diff(mid)[0])
            a = np.abs(np.broadcast_to(np.asarray(size, dtype=float), mid.shape))
            bounds = np.column_stack((mid - 0.5 * a, mid + 0.5 * a))
        g.name, g.midpoints, g.bounds = name, mid, np.asarray(bounds, dtype=float)
        g.flipped, g.side, g.index = flipped, side, index
        return g

    def grid(x, y, xkw=None, ykw=None):
        g = object.__new__(st.StructuredGrid2d)
        g.xbounds = axis(x, "x", **(xkw or {}))
        g.ybounds = axis(y, "y", **(ykw or {}))
        return g

    ns = types.SimpleNamespace(module=st, axis=axis, grid=grid)
    _cache["st"] = ns
    return ns
