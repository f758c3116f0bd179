"""Sparse voxel grid living in HBM — the stand-in for the `fvdb.GridBatch` members the hot path touches
(.ijk, .grid_to_world, .world_to_grid, .ijk_to_index, .total_voxels, .device,
gridbatch_from_points).  Built and traversed by the CUDA kernels in csrc/raster.cu."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from .._lib import ICError, check, lib, require_device


def _f3(v: Sequence[float]):
    a = (C.c_float * 3)(*[float(x) for x in v])
    return a


class VoxelGrid:
    """gridbatch_from_points + per-voxel arg-max labels (reference: infinicube/utils/fvdb_utils.py:141-193)."""

    def __init__(self, points: torch.Tensor, voxel_sizes=(0.1, 0.1, 0.1), origins=(0.05, 0.05, 0.05),
                 semantics: Optional[torch.Tensor] = None, instance: Optional[torch.Tensor] = None):
        require_device()
        if not points.is_cuda:
            raise ICError("VoxelGrid needs CUDA points (there is no CPU path)")
        if points.ndim != 2 or points.shape[1] != 3 or points.shape[0] == 0:
            raise ValueError(f"points must be (N, 3) with N > 0, got {tuple(points.shape)}")
        pts = points.detach().to(torch.float32).contiguous()
        sem = None if semantics is None else semantics.detach().to(device=pts.device, dtype=torch.int32).contiguous()
        ins = None if instance is None else instance.detach().to(device=pts.device, dtype=torch.int32).contiguous()
        self.device = pts.device
        self.voxel_sizes = [float(v) for v in voxel_sizes]
        self.origins = [float(v) for v in origins]
        h = C.c_void_p()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().ic_grid_build(C.c_void_p(pts.data_ptr()), pts.shape[0], _f3(self.voxel_sizes), _f3(self.origins),
                                  None if sem is None else C.c_void_p(sem.data_ptr()),
                                  None if ins is None else C.c_void_p(ins.data_ptr()), C.byref(h), st), "ic_grid_build")
        self._h = h
        self._ijk = self._sem = self._inst = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().ic_grid_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._h = None

    # ---- GridBatch-like surface -------------------------------------------------------------
    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def total_voxels(self) -> int:
        return int(lib().ic_grid_num_voxels(self._h))

    @property
    def num_bricks(self) -> int:
        return int(lib().ic_grid_num_bricks(self._h))

    def info(self) -> dict:
        arrs = [(C.c_int * 3)() for _ in range(4)]
        check(lib().ic_grid_info(self._h, *arrs), "ic_grid_info")
        return {k: np.array(list(a), dtype=np.int32) for k, a in zip(("imin", "imax", "bmin", "bdim"), arrs)}

    def _export(self):
        if self._ijk is None:
            n = self.total_voxels
            ijk = torch.empty((n, 3), dtype=torch.int32, device=self.device)
            sem = torch.empty(n, dtype=torch.int32, device=self.device)
            inst = torch.empty(n, dtype=torch.int32, device=self.device)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(lib().ic_grid_export(self._h, C.c_void_p(ijk.data_ptr()), C.c_void_p(sem.data_ptr()),
                                       C.c_void_p(inst.data_ptr()), st), "ic_grid_export")
            self._ijk, self._sem, self._inst = ijk, sem, inst

    @property
    def ijk(self) -> torch.Tensor:
        """(N_voxel, 3) int32 voxel coordinates in this grid's voxel-index order."""
        self._export()
        return self._ijk

    @property
    def semantics(self) -> torch.Tensor:
        self._export()
        return self._sem

    @property
    def instance(self) -> torch.Tensor:
        self._export()
        return self._inst

    def grid_to_world(self, ijk: torch.Tensor) -> torch.Tensor:
        vs = torch.tensor(self.voxel_sizes, device=ijk.device, dtype=torch.float32)
        org = torch.tensor(self.origins, device=ijk.device, dtype=torch.float32)
        return ijk.to(torch.float32) * vs + org

    def world_to_grid(self, xyz: torch.Tensor) -> torch.Tensor:
        vs = torch.tensor(self.voxel_sizes, device=xyz.device, dtype=torch.float32)
        org = torch.tensor(self.origins, device=xyz.device, dtype=torch.float32)
        return (xyz.to(torch.float32) - org) / vs

    def to(self, device):  # GridBatch.to(): the grid only exists on its CUDA device
        return self
