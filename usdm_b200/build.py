"""Build recipe for libtvb.so (the C-ABI CUDA library), in-tree, sm_100a only.

    python -m usdm_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting usdm_b200/csrc/libtvb.so is git-ignored but travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libtvb.so")
SOURCES = ["tvb.cu"]
HEADERS = ["common.cuh", "gemm_tc.cuh", "gemmln_tc.cuh", "attn_tc.cuh", "attn4_tc.cuh", "attn5_tc.cuh", "attn6_tc.cuh", "attn7_tc.cuh", "attn8_tc.cuh", "attn9_tc.cuh", "attn10_tc.cuh", "conv_tc.cuh", "glue.cuh", "simt.cuh", "vocoder.cuh", "vocoder_host.inl",
           os.path.join("..", "..", "include", "tvb.h"), os.path.join("..", "..", "include", "tvb_vocoder.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libtvb.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB + ".tmp", *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
