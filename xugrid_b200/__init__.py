"""
xugrid_b200 -- B200-native regridding engine behind the xugrid.regrid API
(OverlapRegridder, RelativeOverlapRegridder, CentroidLocatorRegridder,
BarycentricInterpolator, NetworkGridder).  The compute path is hand-written sm_100a CUDA in
libxgregrid.so (C ABI in include/xg_regrid.h); there is no CPU fallback.
"""

from .grid import StructuredGrid2d, Ugrid1d, Ugrid2d, UgridDataArray  # noqa: F401
from .celltree import CellTree2d  # noqa: F401
from .regridder import (  # noqa: F401
    BarycentricInterpolator,
    CentroidLocatorRegridder,
    NetworkGridder,
    OverlapRegridder,
    RelativeOverlapRegridder,
    create_percentile_method,
)
from .sparse import MatrixCOO, MatrixCSR  # noqa: F401

__version__ = "0.1.0"
