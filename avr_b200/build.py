"""Build recipe of ``libavr_b200.so`` (hand-written sm_100a CUDA + the C-ABI in ``include/avr_b200.h``).

The library is built IN-TREE (``avr_b200/libavr_b200.so``) so that it travels with the repo snapshot
to the GPU box; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libavr_b200.so")
SOURCES = ["abi.cu", "hashgrid.cu", "gemm.cu", "composite.cu", "umma_gemm.cu", "mlp_chain.cu", "collapse.cu", "optim.cu", "criterion.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libavr_b200.so")
    return cand


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, "..", "include", "avr_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """``experiments=True`` (``--experiments``) compiles the timing switches in (``-DAVR_EXPERIMENTS``: environment
    variables that force tile shapes or switch epilogue stages off).  The default library contains none of them."""
    if not force and not experiments and not needs_build():
        return LIB_PATH
    objs = []
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *(["-DAVR_EXPERIMENTS"] if experiments else []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-cudart", "shared"]
    subprocess.check_call(link)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv))
