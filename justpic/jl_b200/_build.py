"""Builds libjustpic_sm100a.so in-tree with nvcc for sm_100a (no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "csrc" / "justpic_sm100a.cu"
DEPS = sorted((HERE / "csrc").glob("*.cu*")) + sorted((HERE / "csrc").glob("*.h")) + [HERE.parent.parent / "include" / "justpic_c.h"]
LIB = HERE / "libjustpic_sm100a.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # no implicit FMA contraction; fma() is explicit (bit-exact contract)
    "-diag-suppress", "550,128",  # "set but never used" from the PREP() prologue shared by all entry points
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libjustpic_sm100a.so")
    return nvcc


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in DEPS if p.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", str(LIB), str(SRC)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
