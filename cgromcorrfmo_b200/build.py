"""In-tree build of the native pieces: libcgromcorr.so (CUDA, sm_100a) and the traj2dEcsv CLI driver.

nvcc cross-compiles for sm_100a without a GPU.  The built files sit next to this module (git-ignored, but they travel
to the GPU box with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcgromcorr.so")
CLI = os.path.join(HERE, "traj2dEcsv")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-pthread"]
LIB_SOURCES = ["kernels.cu", "format_kernel.cu", "eigh_kernel.cu", "timecorr_kernel.cu", "parse_kernel.cu", "xtc_kernel.cu", "capi.cu",
               "host_parse.cpp", "xtc_host.cpp"]
LIB_DEPS = LIB_SOURCES + ["kernels.cuh", "cdc_kernel.cuh", "ptx_helpers.cuh", "fmt_g.h", "host_parse.h", "xtc_bits.h", "../../include/cgromcorr.h"]
CLI_SOURCES = ["traj2dEcsv.cpp"]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in deps if os.path.exists(os.path.join(CSRC, d)))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libcgromcorr.so and the CLI if their sources are newer.  Returns the library path."""
    have_sources = all(os.path.exists(os.path.join(CSRC, s)) for s in LIB_SOURCES)
    if not have_sources:
        raise RuntimeError("csrc/ sources missing")
    if force or _stale(LIB, LIB_DEPS):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in LIB_SOURCES] + ["-lcudart"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if os.path.exists(os.path.join(CSRC, CLI_SOURCES[0])) and (force or _stale(CLI, CLI_SOURCES + ["../../include/cgromcorr.h"]) or
                                                               os.path.getmtime(CLI) < os.path.getmtime(LIB)):
        cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-o", CLI] + [os.path.join(CSRC, s) for s in CLI_SOURCES] + \
              ["-L" + HERE, "-lcgromcorr", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
