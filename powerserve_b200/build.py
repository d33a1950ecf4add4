"""Build libps_cuda.so (hand-written sm_100a kernels + the C ABI) in-tree with nvcc.  No torch involved.

    python -m powerserve_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the resulting .so sits next to this file (git-ignored, but it
travels to the GPU box with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libps_cuda.so")
SOURCES = ["ps_cuda.cu"]
HEADERS = ["ps_step.cuh", "ps_tc.cuh", "ps_rw.cuh", "ps_decode.cuh", "ps_kernels.cuh", "ps_math.cuh", os.path.join("..", "..", "include", "ps_cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # never contract a*b+c: FMAs appear only where the reference has them (ps_math.cuh)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-cudart", "shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(os.path.normpath(d)) > t for d in deps if os.path.exists(os.path.normpath(d)))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PS_NVCC_EXTRA", "").split()   # e.g. PS_NVCC_EXTRA=-DPS_ST_DEBUG=1 (ring-protocol assertions in the step kernel)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


SPEC_LIB = os.path.join(HERE, "libps_spec.so")
SPEC_SRC = os.path.join(HERE, "host", "spec_decode.cpp")


def build_spec(force: bool = False) -> str:
    """libps_spec.so: the host-side speculative decoder (include/ps_spec.h), plain C++ over the C ABI of libps_cuda.so."""
    deps = [SPEC_SRC, os.path.join(HERE, "..", "include", "ps_spec.h"), os.path.join(HERE, "..", "include", "ps_cuda.h"), LIB]
    if not force and os.path.exists(SPEC_LIB) and all(os.path.getmtime(SPEC_LIB) >= os.path.getmtime(d) for d in deps if os.path.exists(d)):
        return SPEC_LIB
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", SPEC_SRC, "-o", SPEC_LIB, "-L" + HERE, "-lps_cuda", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return SPEC_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_spec(force="--force" in sys.argv))
