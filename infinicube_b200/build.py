"""Builds the sm_100a shared library in-tree with nvcc (no JIT cache: the .so must travel with the repo).

    python -m infinicube_b200.build [--force]

`ICB_NVCC_EXTRA="-DICB_FMHA_WHATIF_BUILD"` (with --force) appends developer flags, e.g. the measurement-only attention
variants that the shipped library does not contain.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libinfinicube_b200.so"
SOURCES = ["api.cu", "gemm_sm100.cu", "fmha_sm100.cu", "dit_ops.cu", "dit_engine.cu", "raster.cu", "conv_sm100.cu", "vae_ops.cu", "t5_ops.cu", "knn.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def _stale(out: Path, deps: list[Path]) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG_DIR.parent / "include" / "infinicube_b200.h"]
    objdir = PKG_DIR / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        s = CSRC / src
        if not s.exists():
            continue
        o = objdir / (s.stem + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("ICB_NVCC_EXTRA", "").split(), "-c", str(s), "-o", str(o)]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    if force or procs or _stale(LIB_PATH, objs):
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs), "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
