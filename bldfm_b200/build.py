"""Build libbldfm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m bldfm_b200.build [--force]
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libbldfm_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # rounding is controlled explicitly: no implicit FMA contraction on device or host
    "--fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "bldfm_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cuda_lib = Path(_nvcc()).resolve().parent.parent / "lib64"
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", str(LIB), *map(str, sources()),
           "-I", str(PKG.parent / "include"),
           "-L", str(cuda_lib), "-lcufft", "-ldl",
           "-Xlinker", f"-rpath={cuda_lib}"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=str(CSRC))
    return LIB


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(out)
