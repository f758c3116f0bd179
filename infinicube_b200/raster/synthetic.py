"""Deterministic synthetic voxel worlds and camera paths: the workload of the
rasteriser benchmarks and parity tests.  Pure numpy; no GPU."""
from __future__ import annotations

import numpy as np

ROAD, BUILDING, POLE, CAR = 18, 14, 10, 1
DEFAULT_INTRINSICS = np.array([890.5, 770.6, 406.9, 240.4, 832, 480], dtype=np.float64)


def synthetic_scene(S: int, voxel_size: float = 0.2, seed: int = 7):
    """Cubic index extent S: ground slab at k = S/8 (ROAD), S/8 hollow boxes (BUILDING), S/16 poles (POLE),
    S/16 solid 24x10x8-voxel cars (CAR, instance ids 1..).  Returns voxel-centre points [N,3] fp32 in the
    grid frame (origin = voxel_size/2), semantics int32, instance int32, and the ijk array."""
    rng = np.random.RandomState(seed)
    occ = np.zeros((S, S, S), dtype=np.uint8)  # label+1, index [i, j, k]
    inst = np.zeros((S, S, S), dtype=np.int32)
    gk = S // 8
    occ[:, :, gk] = ROAD + 1
    for _ in range(S // 8):
        w, d = rng.randint(S // 16 + 2, S // 4 + 3, size=2)
        hgt = rng.randint(S // 8, S // 2 + 1)
        i0, j0 = rng.randint(0, S - w), rng.randint(0, S - d)
        k1 = min(S - 1, gk + hgt)
        box = occ[i0:i0 + w, j0:j0 + d, gk + 1:k1 + 1]
        box[0, :, :] = box[-1, :, :] = BUILDING + 1
        box[:, 0, :] = box[:, -1, :] = BUILDING + 1
        box[:, :, -1] = BUILDING + 1
    for _ in range(S // 16):
        i0, j0 = rng.randint(0, S - 2, size=2)
        hgt = rng.randint(S // 16 + 2, S // 4 + 2)
        occ[i0:i0 + 2, j0:j0 + 2, gk + 1:min(S, gk + 1 + hgt)] = POLE + 1
    cl, cw, ch = min(24, S // 3), min(10, S // 6), min(8, S // 8)
    for c in range(S // 16):
        i0, j0 = rng.randint(0, S - cl), rng.randint(0, S - cw)
        occ[i0:i0 + cl, j0:j0 + cw, gk + 1:gk + 1 + ch] = CAR + 1
        inst[i0:i0 + cl, j0:j0 + cw, gk + 1:gk + 1 + ch] = c + 1
    ijk = np.argwhere(occ > 0).astype(np.int32)
    sem = occ[ijk[:, 0], ijk[:, 1], ijk[:, 2]].astype(np.int32) - 1
    ins = inst[ijk[:, 0], ijk[:, 1], ijk[:, 2]].astype(np.int32)
    ins[sem != CAR] = 0
    pts = (ijk.astype(np.float32) * np.float32(voxel_size) + np.float32(voxel_size / 2)).astype(np.float32)
    return pts, sem, ins, ijk


def flu_to_opencv_rotation() -> np.ndarray:
    """Camera axes (x right, y down, z forward) expressed in a FLU body frame (x forward, y left, z up):
    columns are the camera axes in FLU (infinicube/camera/base.py:74-115 convention)."""
    return np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


def synthetic_poses(S: int, n: int = 93, voxel_size: float = 0.2, seed: int = 8) -> np.ndarray:
    """n camera->grid poses on a straight line along +x at z = ground + 1.6 m, y centred, yaw jitter +-2 deg."""
    rng = np.random.RandomState(seed)
    ext = S * voxel_size
    xs = np.linspace(0.1 * ext, 0.6 * ext, n)
    z = (S // 8 + 1) * voxel_size + 1.6
    y = 0.5 * ext
    base = flu_to_opencv_rotation()
    poses = np.zeros((n, 4, 4), dtype=np.float64)
    for i in range(n):
        yaw = np.deg2rad(rng.uniform(-2.0, 2.0))
        c, s = np.cos(yaw), np.sin(yaw)
        rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        poses[i, :3, :3] = rz @ base
        poses[i, :3, 3] = [xs[i], y, z]
        poses[i, 3, 3] = 1.0
    return poses.astype(np.float32)


def raster_algorithmic_bytes(n_vox: int, n_cam: int, H: int, W: int) -> int:
    """each frame reads every voxel record once (20 B) and writes 12 B per pixel."""
    return n_cam * (20 * n_vox + 12 * H * W)
