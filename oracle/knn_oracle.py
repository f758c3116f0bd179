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


def nnk(queries: np.ndarray, ref: np.ndarray, k: int):
    """k nearest neighbours by definition (knn_query_fast with nb_points = k, knn.cu:15-51; the small-cloud branch
    there is literally cdist + topk): d2 = ((qx-px)^2 + (qy-py)^2) + (qz-pz)^2 in fp32 (numpy elementwise ops do not
    contract into FMAs), rows ascending by (d2, reference index).  -> (idx int32 [n, k], d2 fp32 [n, k]); slots beyond
    the number of reference points hold (-1, +inf).  Quadratic memory: small clouds only."""
    q = np.ascontiguousarray(queries, dtype=np.float32)[:, :3]
    r = np.ascontiguousarray(ref, dtype=np.float32)[:, :3]
    dx = q[:, None, 0] - r[None, :, 0]
    dy = q[:, None, 1] - r[None, :, 1]
    dz = q[:, None, 2] - r[None, :, 2]
    d = (dx * dx + dy * dy) + dz * dz
    order = np.argsort(d, axis=1, kind="stable")[:, :k]       # stable: equal distances keep index order
    d2 = np.take_along_axis(d, order, axis=1)
    idx = order.astype(np.int32)
    if r.shape[0] < k:
        pad = k - r.shape[0]
        idx = np.concatenate([idx, np.full((q.shape[0], pad), -1, np.int32)], axis=1)
        d2 = np.concatenate([d2, np.full((q.shape[0], pad), np.inf, np.float32)], axis=1)
    return idx, d2.astype(np.float32)


def color_from_points(target_pcs: np.ndarray, ref_pcs: np.ndarray, ref_colors: np.ndarray, k: int = 8) -> np.ndarray:
    """color_util.py:21-49: inverse-distance weights 1 / (sqrt(d2) + 1e-8), normalised over the k neighbours."""
    if target_pcs.shape[0] == 0:
        return np.zeros((0, 3), np.float32)
    idx, d2 = nnk(target_pcs, ref_pcs, k)
    w = 1.0 / (np.sqrt(d2) + np.float32(1e-8))
    w = w / w.sum(axis=1, keepdims=True)
    return (w[..., None] * np.asarray(ref_colors, np.float32)[idx]).sum(axis=1).astype(np.float32)
