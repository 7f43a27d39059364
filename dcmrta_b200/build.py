"""dcmrta_b200/build.py -- compile the CUDA extension in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SO = PKG / "libdcmrta_b200.so"
SOURCES = [CSRC / "dcm_kernels.cu"]
HEADERS = [CSRC / "dcm_thread.cuh", CSRC / "dcm_soa.h", CSRC / "dcm_layout.h", PKG.parent / "include" / "dcmrta.h"]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-fmad=false",                                   # fp64 event clock must round exactly like the reference
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: cannot build libdcmrta_b200.so")
    return exe


def stale() -> bool:
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if force or stale():
        cmd = [nvcc(), *NVCC_FLAGS, "-o", str(SO), *map(str, SOURCES)]
        # DCM_BUILD_FAST=1 (development only): optimise the kernels of the one translation unit in parallel, 3 min -> under 1.
        # Not the default: register allocation then differs from build to build (k_step 0..64 B of spills, k_obs_tile 76..94 registers).
        if os.environ.get("DCM_BUILD_FAST") == "1":
            cmd[1:1] = ["--split-compile", "0"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
