"""Build libptg_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libptg_b200.so")
SOURCES = ["ptg_capi.cu"]
DEPS = ["ptg_capi.cu", "ptg_kernels.cuh", "ptg_train.cuh", "ptg_device.cuh", "ptg_rng.cuh", "ptg_ziggurat_tables.h",
        os.path.join("..", "..", "include", "ptg_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # reward arithmetic must round like the reference's unfused fp64 expression
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in DEPS + ["build.py"])


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    if not force and not needs_build():
        return SO
    cmd = [find_nvcc(), *NVCC_FLAGS, *(extra or []), "-o", SO, *[os.path.join(HERE, s) for s in SOURCES]]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=HERE)
    return SO


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a != "--force"]     # e.g.  python build.py --force -Xptxas -v
    print(build(force="--force" in sys.argv or bool(extra), verbose=True, extra=extra))
