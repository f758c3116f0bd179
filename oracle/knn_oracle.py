"""Python front end of the nearest-neighbour oracle — TEST INFRASTRUCTURE ONLY (see knn_oracle.c for what it
restates and how it is pinned).  Only tests/ and __graft_entry__.smoke() may import this module."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "knn_oracle.c"
BUILD_DIR = HERE / "_build"
LIB = BUILD_DIR / "libknn_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    BUILD_DIR.mkdir(exist_ok=True)
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(LIB), str(SRC),
                        "-lm"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.ko_nn1.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]
        _lib.ko_nn1.restype = None
    return _lib


def nn1(queries: np.ndarray, ref: np.ndarray):
    """-> (idx int32 [n], d2 fp32 [n]); ties -> smallest reference index."""
    q = np.ascontiguousarray(queries, dtype=np.float32)
    r = np.ascontiguousarray(ref, dtype=np.float32)
    idx = np.empty(q.shape[0], dtype=np.int32)
    d2 = np.empty(q.shape[0], dtype=np.float32)
    lib().ko_nn1(r.ctypes.data, r.shape[0], r.shape[1], q.ctypes.data, q.shape[0], q.shape[1], idx.ctypes.data, d2.ctypes.data)
    return idx, d2


def semantic_from_points(target_pcs: np.ndarray, ref_pcs: np.ndarray, ref_semantic: np.ndarray) -> np.ndarray:
    """color_util.py:52-60: label of the nearest reference point, int64; empty target -> empty result."""
    if target_pcs.shape[0] == 0:
        return np.zeros((0,), dtype=np.int64)
    idx, _ = nn1(target_pcs, ref_pcs)
    return np.asarray(ref_semantic)[idx].astype(np.int64)
