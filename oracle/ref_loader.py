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
