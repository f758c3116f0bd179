"""Mirror of infinicube/voxelgen/utils/color_util.py (`semantic_from_points` :52-60, `color_from_points` :21-49) and
of the extension call they make, `common.knn_query_fast(queries, ref, k)` (infinicube/voxelgen/ext/common/knn.cu:15-50), on the sm_100a
cell-grid search of csrc/knn.cu.  Called by the stage-1 chunk merge
(infinicube/inference/voxel_generation_single_chunk.py:280) and `transform_grid_and_semantic`
(infinicube/voxelgen/utils/extrap_util.py:233-276).  No CPU path."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ..._lib import ICError, check, lib, require_device


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _points(t: torch.Tensor, name: str) -> torch.Tensor:
    if not torch.is_tensor(t) or not t.is_cuda:
        raise ICError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")      # CHECK_IS_FLOAT of the reference extension
    if t.dim() != 2 or t.shape[1] < 3:
        raise ValueError(f"{name} must be (N, >=3), got {tuple(t.shape)}")
    return t if t.stride(1) == 1 else t.contiguous()


class KnnIndex:
    """Counting-sorted cell grid over a reference cloud; build once, query many times."""

    def __init__(self, ref_xyz: torch.Tensor, cell_size: float = 0.0):
        require_device()
        self.ref = _points(ref_xyz, "ref_xyz")
        if self.ref.shape[0] == 0:
            raise ValueError("reference cloud is empty")
        self._h = C.c_void_p()
        with torch.cuda.device(self.ref.device):
            check(lib().ic_knn_build(_p(self.ref), self.ref.shape[0], self.ref.stride(0), float(cell_size),
                                     C.byref(self._h), _stream()), "ic_knn_build")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().ic_knn_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._h = None

    def info(self) -> dict:
        n, c, s, d = C.c_longlong(), C.c_longlong(), C.c_float(), (C.c_int * 3)()
        check(lib().ic_knn_info(self._h, C.byref(n), C.byref(c), C.byref(s), d), "ic_knn_info")
        return {"points": n.value, "cells": c.value, "cell_size": s.value, "dims": tuple(d)}

    def query(self, queries: torch.Tensor, labels: Optional[torch.Tensor] = None, want_dist: bool = True,
              want_idx: bool = True):
        """-> (d2 fp32 [n] | None, idx int32 [n] | None, labels int64 [n] | None)."""
        q = _points(queries, "queries")
        n, dev = q.shape[0], q.device
        if dev != self.ref.device:
            raise ICError("queries and reference live on different devices")
        d2 = torch.empty(n, dtype=torch.float32, device=dev) if want_dist else None
        idx = torch.empty(n, dtype=torch.int32, device=dev) if want_idx else None
        lab_out = None
        if labels is not None:
            if labels.shape[0] != self.ref.shape[0]:
                raise ValueError(f"{labels.shape[0]} labels for {self.ref.shape[0]} reference points")
            if labels.dim() != 1:
                raise ValueError("labels must be one value per reference point")
            labels = labels.to(dev, torch.int64).contiguous()
            lab_out = torch.empty(n, dtype=torch.int64, device=dev)
        if n:
            with torch.cuda.device(dev):
                check(lib().ic_knn_query1(self._h, _p(q), n, q.stride(0), _p(labels), _p(idx), _p(d2), _p(lab_out),
                                          _stream()), "ic_knn_query1")
        return d2, idx, lab_out

    def query_k(self, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """k nearest reference points per query -> (d2 fp32 [n, k], idx int32 [n, k]), rows ascending by
        (squared distance, reference index); fewer than k reference points leave (+inf, -1) in the tail slots."""
        q = _points(queries, "queries")
        n, dev = q.shape[0], q.device
        if dev != self.ref.device:
            raise ICError("queries and reference live on different devices")
        d2 = torch.empty((n, k), dtype=torch.float32, device=dev)
        idx = torch.empty((n, k), dtype=torch.int32, device=dev)
        if n:
            with torch.cuda.device(dev):
                check(lib().ic_knn_query(self._h, _p(q), n, q.stride(0), k, _p(idx), _p(d2), _stream()), "ic_knn_query")
        return d2, idx


def knn_query_fast(queries: torch.Tensor, ref_xyz: torch.Tensor, nb_points: int,
                   cell_size: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(squared distances fp32 [n, nb_points], indices int32 [n, nb_points]) - the reference extension's return
    convention (voxelgen/ext/common/knn.cu:15-51); every row ascending by (distance, reference index)."""
    if not 1 <= int(nb_points) <= 32:
        raise ICError(f"knn_query_fast: nb_points must be in [1, 32], got {nb_points}")
    index = KnnIndex(ref_xyz, cell_size)
    if nb_points == 1:
        d2, idx, _ = index.query(queries)
        return d2[:, None], idx[:, None]
    return index.query_k(queries, int(nb_points))


def color_from_points(target_pcs: torch.Tensor, ref_pcs: torch.Tensor, ref_colors: torch.Tensor, k: int = 8,
                      cell_size: float = 0.0) -> torch.Tensor:
    """Inverse-distance-weighted colour of the k nearest reference points (color_util.py:21-49): weights
    1 / (dist + 1e-8), normalised over the k neighbours."""
    if target_pcs.shape[0] == 0:
        return torch.zeros((0, 3), dtype=torch.float32, device=target_pcs.device)
    dist, idx = knn_query_fast(target_pcs.contiguous(), ref_pcs.contiguous(), k, cell_size)
    dist = dist.sqrt()
    knn_color = ref_colors[idx.long()]
    weight = 1 / (dist + 1e-8)
    weight = weight / weight.sum(dim=1, keepdim=True)
    return (weight.unsqueeze(-1) * knn_color).sum(dim=1)


def semantic_from_points(target_pcs: torch.Tensor, ref_pcs: torch.Tensor, ref_semantic: torch.Tensor,
                         cell_size: float = 0.0) -> torch.Tensor:
    """Label of the nearest reference point for every target point, int64 (color_util.py:52-60).  `cell_size`
    (optional, e.g. the voxel size) only affects speed."""
    if target_pcs.shape[0] == 0:
        return torch.zeros((0), dtype=torch.int64, device=target_pcs.device)
    _, _, lab = KnnIndex(ref_pcs.contiguous(), cell_size).query(target_pcs.contiguous(), ref_semantic, want_dist=False,
                                                                want_idx=False)
    return lab
