"""
Builds libxgregrid.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

    python -m xugrid_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libxgregrid.so")

SOURCES = ["xg_index.cu", "xg_overlap.cu", "xg_locate.cu", "xg_apply.cu", "xg_voronoi.cu"]

# -fmad=false: fp64 products and sums round separately, like the reference's
# numba/LLVM code on x86, so areas and weighted sums are reproducible bit for bit.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-Wno-deprecated-declarations",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libxgregrid.so cannot be built")
    return exe


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "xg_regrid.h"))
    paths.append(os.path.abspath(__file__))
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _deps_mtime()


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """variant/defines: tuning builds (`--variant NAME -DFOO=1`) go to
    _lib/libxgregrid_NAME.so and are selected with XG_LIB_PATH (development only)."""
    global BUILD, LIB_PATH
    if variant:
        BUILD = os.path.join(HERE, "_build", variant)
        LIB_PATH = os.path.join(LIB_DIR, f"libxgregrid_{variant}.so")
        force = True
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers_mtime = max(
        os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(".cuh")
    )
    headers_mtime = max(headers_mtime, os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "xg_regrid.h")))

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        srcp = os.path.join(CSRC, src)
        if (
            not force
            and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(srcp), headers_mtime)
        ):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", srcp, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant,
                 defines=[a for a in sys.argv if a.startswith("-D")])
    print(path)
