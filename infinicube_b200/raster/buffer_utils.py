"""Coordinate guidance buffer (mirrors infinicube/utils/buffer_utils.py:180-265 and
infinicube/utils/depth_utils.py:402-466).  Unprojection and normalisation are csrc/raster.cu kernels; the
100k-point quantile sample is drawn on the device by default and, with reference_sampling=True, by the
reference's own torch.randperm call so that a seeded run replays its sample exactly."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from .._lib import ICError, check, lib, require_device
from .camera import PinholeCamera


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def unproject_to_first_camera(depth_buffer: torch.Tensor, camera_model: PinholeCamera,
                              camera_poses: torch.Tensor) -> torch.Tensor:
    """X_cam0 for every pixel, [N,H,W,3] fp32; pixels with depth == 0 carry the 1e7 sentinel."""
    require_device()
    if not depth_buffer.is_cuda:
        raise ICError("depth_buffer must be a CUDA tensor")
    dev = depth_buffer.device
    depth = depth_buffer.to(torch.float32).contiguous()
    n, h, w = depth.shape
    poses = camera_poses.detach().to("cpu", torch.float32)
    c2c0 = torch.einsum("ij,bjk->bik", torch.inverse(poses[0]), poses).contiguous().to(dev)  # buffer_utils.py:212-217
    kinv = camera_model.intrinsics_matrix_inv_torch.to(dev, torch.float32).contiguous()
    xyz = torch.empty((n, h, w, 3), dtype=torch.float32, device=dev)
    check(lib().ic_coord_unproject(_p(depth), _p(c2c0), _p(kinv), n, h, w, _p(xyz), _stream()), "ic_coord_unproject")
    return xyz


QUANTILE_SAMPLE = 100000  # buffer_utils.py:239-244


def global_quantiles(xyz: torch.Tensor, percentile: float = 0.05, reference_sampling: bool = False,
                     depth: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """mins / ranges from <= 100k randomly sampled valid points (buffer_utils.py:232-249).

    reference_sampling=True replays the reference call for call - `torch.nonzero` over every pixel and a CPU
    `torch.randperm(n_valid)` - so that a seeded run draws exactly the reference's sample (the golden test); at 93
    x 480 x 832 that costs 0.1-0.8 s, almost all of it the 37 M-element CPU permutation.  The default draws the sample
    on the device: uniformly random pixel indices (device generator seeded from torch's global CPU generator, so
    `torch.manual_seed` still makes a run reproducible), invalid pixels rejected, first 100k kept - the same
    distribution (a uniform sample of the valid points; with replacement, which is immaterial at 100k of ~3e7) at
    the cost of a few small kernels."""
    flat = xyz.reshape(-1, 3)
    if reference_sampling:
        valid_idx = torch.nonzero(flat[:, 2] < 1e6, as_tuple=False)[:, 0]
        if valid_idx.numel() == 0:
            return None, None
        perm = torch.randperm(valid_idx.shape[0])[:QUANTILE_SAMPLE].to(flat.device)  # CPU generator, like the reference
        sample = flat[valid_idx[perm]]
    else:
        n = flat.shape[0]
        valid = (depth.reshape(-1) != 0) if depth is not None else (flat[:, 2] < 1e6)
        n_valid = int(torch.count_nonzero(valid))
        if n_valid == 0:
            return None, None
        if n_valid <= QUANTILE_SAMPLE:
            sample = flat[torch.nonzero(valid, as_tuple=False)[:, 0]]
        else:
            gen = torch.Generator(device=flat.device)
            gen.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,))))
            want = int(QUANTILE_SAMPLE * (n / n_valid) * 1.25) + 4096
            picked = []
            have = 0
            while have < QUANTILE_SAMPLE:
                idx = torch.randint(0, n, (want,), device=flat.device, generator=gen)
                idx = idx[valid[idx]]
                picked.append(idx)
                have += idx.numel()
            sample = flat[torch.cat(picked)[:QUANTILE_SAMPLE]]
    mins = torch.quantile(sample, percentile, dim=0)
    maxs = torch.quantile(sample, 1 - percentile, dim=0)
    ranges = torch.clamp(maxs - mins, min=1e-7)
    return mins.contiguous(), ranges.contiguous()


def coordinate_buffer(depth_buffer: torch.Tensor, camera_model: PinholeCamera, camera_poses: torch.Tensor,
                      percentile: float = 0.05, want_f32: bool = True, want_u8: bool = False,
                      reference_sampling: bool = False):
    depth = depth_buffer.to(torch.float32).contiguous()
    xyz = unproject_to_first_camera(depth, camera_model, camera_poses)
    mins, ranges = global_quantiles(xyz, percentile, reference_sampling, depth)
    n, h, w = depth.shape
    dev = depth.device
    if mins is None:  # no valid points: reference returns points*0.5 then sets misses to 1 -> all ones
        f = torch.ones((n, h, w, 3), dtype=torch.float32, device=dev) if want_f32 else None
        u = torch.full((n, h, w, 3), 255, dtype=torch.uint8, device=dev) if want_u8 else None
        return f, u
    f = torch.empty((n, h, w, 3), dtype=torch.float32, device=dev) if want_f32 else None
    u = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    check(lib().ic_coord_normalize(_p(xyz), _p(depth), n * h * w, _p(mins), _p(ranges), _p(f), _p(u), _stream()),
          "ic_coord_normalize")
    return f, u


def generate_coordinate_buffer_from_memory_global_norm(depth_buffer: torch.Tensor, camera_model: PinholeCamera,
                                                       camera_poses: torch.Tensor, percentile: float = 0.05,
                                                       reference_sampling: bool = False) -> torch.Tensor:
    """[N,H,W,3] fp32 in [0,1]; infinitely-far pixels are 1 (same signature as the reference, plus the opt-in exact
    replay of its quantile sample)."""
    f, _ = coordinate_buffer(depth_buffer, camera_model, camera_poses, percentile, want_f32=True, want_u8=False,
                             reference_sampling=reference_sampling)
    return f
