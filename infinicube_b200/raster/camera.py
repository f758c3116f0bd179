"""Pinhole camera mirroring the part of the reference camera model that is on the hot path
(infinicube/camera/pinhole.py:22-138, infinicube/camera/base.py:207-264, 520-618).  The ray / voxel
intersection runs in csrc/raster.cu; rays are generated in registers and never materialised.

Provenance: `PinholeCamera.__init__`, `cache_torch_and_np_intrinsics`, `from_tensor`, `from_numpy`, the `width` /
`height` / `intrinsics` accessors, `rescale` and `get_intrinsics_matrix` keep the reference's attribute names, argument
order and arithmetic (`camera/pinhole.py:22-104`) because callers read those attributes directly; they are trivial
accessors and follow the reference closely rather than being re-derived.  Everything that touches rays or voxels is new."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Union

import numpy as np
import torch

from .._lib import check, lib
from .grid import VoxelGrid


def _torch(x, device=None, dtype=torch.float32) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=dtype)


class PinholeCamera:
    def __init__(self, fx, fy, cx, cy, w, h, dtype=torch.float32, device=None):
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy
        self.w, self.h = int(w), int(h)
        self.dtype = dtype
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.device = device
        self.cache_torch_and_np_intrinsics()

    def cache_torch_and_np_intrinsics(self):
        # K and torch.inverse(K) in fp32, as the reference caches them (pinhole.py:37-43)
        self.intrinsics_matrix_torch = self.get_intrinsics_matrix()
        self.intrinsics_matrix_inv_torch = self.get_inv_intrinsics_matrix()
        self.intrinsics_matrix_np = self.intrinsics_matrix_torch.cpu().numpy()
        self.intrinsics_matrix_inv_np = self.intrinsics_matrix_inv_torch.cpu().numpy()

    @staticmethod
    def from_tensor(x: torch.Tensor):
        return PinholeCamera(x[0], x[1], x[2], x[3], x[4], x[5])

    @staticmethod
    def from_numpy(x: np.ndarray, device=None):
        return PinholeCamera(x[0], x[1], x[2], x[3], x[4], x[5], device=device)

    @property
    def width(self) -> int:
        return self.w

    @property
    def height(self) -> int:
        return self.h

    @property
    def intrinsics(self) -> np.ndarray:
        return np.array([self.fx, self.fy, self.cx, self.cy, self.w, self.h])

    def rescale(self, ratio_h: float, ratio_w: float = None):
        if ratio_w is None:
            ratio_w = ratio_h
        self.w = int(self.w * ratio_w)  # truncation, pinhole.py:69-70
        self.h = int(self.h * ratio_h)
        self.fx, self.fy = self.fx * ratio_w, self.fy * ratio_h
        self.cx, self.cy = self.cx * ratio_w, self.cy * ratio_h
        self.cache_torch_and_np_intrinsics()

    def get_intrinsics_matrix(self) -> torch.Tensor:
        return torch.tensor([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]], device=self.device,
                            dtype=self.dtype)

    def get_inv_intrinsics_matrix(self) -> torch.Tensor:
        k = torch.tensor([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]], dtype=self.dtype)
        return torch.inverse(k).to(self.device)  # inverted on the host: no cuSOLVER call on the path

    # ---- transforms (camera/base.py:229-264) ------------------------------------------------
    @staticmethod
    def transform_points(points, tfm):
        assert isinstance(points, type(tfm)), (
            f"points and tfm must be the same type, but got {type(points)} and {type(tfm)}")
        if isinstance(points, torch.Tensor):
            return (tfm[:3, :3] @ points.T + tfm[:3, 3].unsqueeze(-1)).T
        return (tfm[:3, :3] @ points.T + tfm[:3, 3].reshape(-1, 1)).T

    # ---- voxel rendering (camera/base.py:520-618) --------------------------------------------
    def render_voxel_buffers(self, camera_poses, voxel_grid: VoxelGrid, attr0: Optional[torch.Tensor] = None,
                             attr1: Optional[torch.Tensor] = None, background0: int = 0, background1: int = 0):
        """One fused launch: (zdepth fp32, attr0 image int32, attr1 image int32), each (N, H, W)."""
        poses = _torch(camera_poses, voxel_grid.device).reshape(-1, 16).contiguous()
        n = poses.shape[0]
        dev = voxel_grid.device
        depth = torch.empty((n, self.h, self.w), dtype=torch.float32, device=dev)
        a0 = torch.empty((n, self.h, self.w), dtype=torch.int32, device=dev)
        a1 = torch.empty((n, self.h, self.w), dtype=torch.int32, device=dev)
        kinv = (C.c_float * 9)(*[float(v) for v in self.intrinsics_matrix_inv_np.reshape(-1)])
        if attr0 is not None:
            attr0 = attr0.to(device=dev, dtype=torch.int32).contiguous()
        if attr1 is not None:
            attr1 = attr1.to(device=dev, dtype=torch.int32).contiguous()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().ic_raster_render(voxel_grid.handle, kinv, C.c_void_p(poses.data_ptr()), n, self.w, self.h,
                                     None if attr0 is None else C.c_void_p(attr0.data_ptr()),
                                     None if attr1 is None else C.c_void_p(attr1.data_ptr()), int(background0),
                                     int(background1), C.c_void_p(depth.data_ptr()), C.c_void_p(a0.data_ptr()),
                                     C.c_void_p(a1.data_ptr()), st), "ic_raster_render")
        return depth, a0, a1

    def get_zdepth_map_from_voxel(self, camera_poses, voxel_grid: VoxelGrid) -> torch.Tensor:
        single = len(camera_poses.shape) == 2
        depth, _, _ = self.render_voxel_buffers(camera_poses, voxel_grid)
        return depth[0] if single else depth

    def get_semantic_map_from_voxel(self, camera_poses, voxel_grid: VoxelGrid, voxel_semantic: torch.Tensor,
                                    background_semantic: int = 0) -> torch.Tensor:
        single = len(camera_poses.shape) == 2
        _, sem, _ = self.render_voxel_buffers(camera_poses, voxel_grid, attr0=voxel_semantic,
                                              background0=background_semantic)
        sem = sem.to(voxel_semantic.dtype)
        return sem[0] if single else sem
