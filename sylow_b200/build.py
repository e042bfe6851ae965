"""Compile libsylow_b200.so (sm_100a) in-tree with nvcc.  No JIT cache, no torch extension: the C-ABI
library is plain CUDA runtime code so that a Rust/C caller can link it without Python."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libsylow_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC,-pthread", "-Xptxas", "-v",
]


def sources():
    return [os.path.join(CSRC, "sylow_b200.cu")]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(ROOT, "include", "sylow_b200.h"))
    return d


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Build (if stale) and return the path of libsylow_b200.so."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsylow_b200.so")
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(ROOT, "build", "ptxas.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stderr[-4000:])
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True))
