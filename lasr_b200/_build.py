"""Build liblasr_b200.so (the C-ABI product library) in-tree with nvcc for sm_100a.

``python -m lasr_b200._build`` or ``lasr_b200._build.build_library()``; called by
``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB_PATH = os.environ.get("LASR_B200_LIB") or os.path.join(CSRC, "liblasr_b200.so")  # override: A/B builds
SOURCES = ["host_common.cu", "raster.cu", "lbs.cu", "losses.cu", "skin.cu", "meshreg.cu", "chamfer.cu", "texture_io.cu", "staging.cu"]
HEADERS = ["raster_common.cuh", "raster_backward_pair.inc", "host_common.h"]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(INCLUDE, "lasr_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False, defines=()) -> str:
    """Compile every source to an object in parallel (one nvcc each), then link the shared library."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(CSRC, "build", os.path.basename(LIB_PATH) + ".obj")
    os.makedirs(objdir, exist_ok=True)
    common = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC] + [f"-D{d}" for d in defines]
    if verbose:
        common += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        subprocess.check_call(common + ["-c", os.path.join(CSRC, src), "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force=True, verbose="-v" in sys.argv, defines=[a[2:] for a in sys.argv[1:] if a.startswith("-D")]))
