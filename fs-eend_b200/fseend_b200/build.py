"""Build the sm_100a shared library in-tree: fs-eend_b200/lib/libfseend_b200.so (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_DIR = os.path.join(PKG_ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libfseend_b200.so")
SOURCES = ["gemm.cu", "gemm_persist.cu", "gemm_pair.cu", "attn.cu", "attn2.cu", "attn3.cu", "ffn.cu", "ffn_pair.cu", "retention.cu", "p32.cu", "elementwise.cu", "embloss.cu", "spkfuse.cu", "loss.cu", "train_ops.cu", "train_attn.cu",
           "fs_model.cu"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libfseend_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(PKG_ROOT), "include", "fseend_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd[1:1] = os.environ.get("FSEEND_NVCC_FLAGS", "").split()   # e.g. -DFSEEND_WAIT_LIMIT_SPINS=1000000 (debug)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
