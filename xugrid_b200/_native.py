"""
ctypes binding of libxgregrid.so (declared in include/xg_regrid.h).

There is no CPU fallback: if the library is missing or no CUDA device is
usable, every entry point raises.  PyTorch is used only for device memory and
streams; the signatures of the library contain plain pointers and sizes.
"""

from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# XG_LIB_PATH selects a tuning build of the same library (development only)
LIB_PATH = os.environ.get("XG_LIB_PATH") or os.path.join(_HERE, "_lib", "libxgregrid.so")

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double

# name -> (restype, argtypes); must list every symbol of include/xg_regrid.h
SIGNATURES = {
    "xg_abi_version": (_int, []),
    "xg_last_error_string": (ctypes.c_char_p, []),
    "xg_release_cached_memory": (_int, []),
    "xg_index_create": (_int, [_vp, _i64, _vp, _int, _i64, _int, _vp, ctypes.POINTER(_vp)]),
    "xg_index_create_clipped": (_int, [_vp, _i64, _vp, _int, _i64, _int, ctypes.POINTER(_dbl), _vp,
                                       ctypes.POINTER(_vp)]),
    "xg_index_create_for_targets": (_int, [_vp, _i64, _vp, _int, _i64, _int, _vp, _vp, _int, _i64, _i64, _int, _vp,
                                           ctypes.POINTER(_vp)]),
    "xg_index_create_for_target_ids": (_int, [_vp, _i64, _vp, _int, _i64, _int, _vp, _vp, _int, _vp, _i64, _int, _vp,
                                              ctypes.POINTER(_vp)]),
    "xg_target_shard": (_int, [_vp, _i64, _vp, _int, _i64, _int, _int, _int, _vp, ctypes.POINTER(_i64), _vp]),
    "xg_face_range_bounds": (_int, [_vp, _vp, _int, _i64, _i64, _int, ctypes.POINTER(_dbl), _vp]),
    "xg_index_destroy": (_int, [_vp]),
    "xg_index_info": (_int, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(_dbl)]),
    "xg_overlap_begin": (
        _int,
        [_vp, _vp, _i64, _vp, _int, _i64, _int, _i64, _i64, _vp, ctypes.POINTER(_i64), _vp,
         ctypes.POINTER(_vp)],
    ),
    "xg_overlap_begin_ids": (
        _int,
        [_vp, _vp, _i64, _vp, _int, _i64, _int, _vp, _i64, _vp, ctypes.POINTER(_i64), _vp, ctypes.POINTER(_vp)],
    ),
    "xg_overlap_finish": (_int, [_vp, _vp, _vp, _vp, _int, _vp]),
    "xg_overlap_run": (
        _int,
        [_vp, _vp, _i64, _vp, _int, _i64, _int, _vp, _i64, _i64, _int, _vp, _vp, _vp, _i64, _int, _vp,
         ctypes.POINTER(_vp)],
    ),
    "xg_overlap_result": (_int, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(_int)]),
    "xg_overlap_stats": (_int, [_vp, ctypes.POINTER(_i64)]),
    "xg_overlap_timings": (_int, [_vp, ctypes.POINTER(_dbl)]),
    "xg_overlap_abort": (_int, [_vp]),
    "xg_voronoi_begin": (_int, [_vp, _i64, _vp, _int, _i64, _int, _vp, ctypes.POINTER(_i64), _vp, ctypes.POINTER(_vp)]),
    "xg_voronoi_finish": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "xg_voronoi_abort": (_int, [_vp]),
    "xg_bary_csr_count": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, ctypes.POINTER(_i64), _vp]),
    "xg_bary_csr_fill": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xg_locate_points": (_int, [_vp, _vp, _i64, _dbl, _vp, _vp]),
    "xg_locate_any": (_int, [_vp, _vp, _i64, _dbl, _vp, _vp]),
    "xg_barycentric": (_int, [_vp, _vp, _i64, _dbl, _vp, _vp, _vp]),
    "xg_intersect_edges": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, ctypes.POINTER(_i64), _vp]),
    "xg_unpack_u24": (_int, [_vp, _i64, _vp, _vp]),
    "xg_face_area": (_int, [_vp, _vp, _int, _i64, _int, _vp, _vp]),
    "xg_face_centroid": (_int, [_vp, _vp, _int, _i64, _int, _vp, _vp]),
    "xg_csr_apply": (_int, [_int, _dbl, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "xg_csr_apply_peers": (_int, [_int, _dbl, _vp, _vp, _vp, _vp, _i64, _i64, _i64, ctypes.POINTER(_vp), _int, _i64, _vp]),
    "xg_coo_gather": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _vp]),
}

_lib = None


class XgError(RuntimeError):
    pass


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen the library and declare every signature (no GPU needed)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise XgError(
            f"{p} not found: build it with `python -m xugrid_b200.build` "
            "(xugrid_b200 has no CPU fallback)"
        )
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.xg_abi_version() != 1:
        raise XgError("libxgregrid.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def lib() -> ctypes.CDLL:
    L = load_library()
    if not torch.cuda.is_available():
        raise XgError("xugrid_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().xg_last_error_string()
        raise XgError(f"libxgregrid error {rc}: {msg.decode() if msg else ''}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = None) -> Optional[int]:
    """Raw device pointer of a contiguous CUDA tensor; with ``dtype`` the element type
    is checked too (the library reinterprets the bytes it is handed)."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise XgError("expected a contiguous CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        raise XgError(f"expected a {dtype} tensor, received {t.dtype}")
    return t.data_ptr()
