"""In-tree build of the CUDA extension (sm_100a only).

`nvcc` cross-compiles without a GPU, so this runs in the CPU build container; the built
`reslic_tcm_b200/lib/libreslic_b200.so` travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libreslic_b200.so")
CUDA_SOURCES = ["cabi.cu", "gc_fused.cu", "gc_bwd.cu", "eb_fused.cu", "eb_bwd.cu", "stanh_fused.cu", "rans_slots.cu", "rate_reduce.cu", "rate_exchange.cu", "cdf_tables.cpp", "rans.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # IEEE behaviour is part of the parity contract: no --use_fast_math, fp32 division and
    # sqrt stay correctly rounded, denormals are kept.
    "--fmad=true", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return nvcc


def sources():
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "reslic_b200.h"))
    return srcs, deps


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    _, deps = sources()
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA translation unit into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs, _ = sources()
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
