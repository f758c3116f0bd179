"""Python front end of the rasteriser oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The traversal lives in raster_oracle.c (compiled here with gcc, FMA contraction off);
the in-repo post-processing of the reference (camera intrinsics, palette, coordinate buffer) is
restated below in numpy / torch-CPU, each function citing the reference lines it follows, and is
pinned against tests/golden/ fixtures generated from the reference's own code by
oracle/gen_golden_from_reference.py.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
SRC = HERE / "raster_oracle.c"
BUILD_DIR = HERE / "_build"
LIB = BUILD_DIR / "libraster_oracle.so"

_lib = None


def build(force: bool = False) -> Path:
    BUILD_DIR.mkdir(exist_ok=True)
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(LIB), str(SRC), "-lm"]
        subprocess.run(cmd, check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        vp, ci, ll = C.c_void_p, C.c_int, C.c_longlong
        L.ro_voxelize.restype = vp
        L.ro_voxelize.argtypes = [vp, ll, vp, vp, vp, vp]
        L.ro_free.argtypes = [vp]
        L.ro_num_voxels.restype = ll
        L.ro_num_voxels.argtypes = [vp]
        L.ro_num_bricks.restype = ll
        L.ro_num_bricks.argtypes = [vp]
        L.ro_grid_info.argtypes = [vp, vp, vp, vp, vp]
        L.ro_export.argtypes = [vp, vp, vp, vp]
        L.ro_render.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
        L.ro_render_flat.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
        L.ro_render_rows.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, vp, vp]
        L.ro_semantic_rgb.argtypes = [vp, vp, ll, vp, vp, vp, ci, vp]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleGrid:
    """points_to_fvdb restated (infinicube/utils/fvdb_utils.py:71-216)."""

    def __init__(self, points: np.ndarray, voxel_sizes, origins, sem: Optional[np.ndarray] = None,
                 inst: Optional[np.ndarray] = None):
        pts = np.ascontiguousarray(points, dtype=np.float32)
        self.vs = np.asarray(voxel_sizes, dtype=np.float32)
        self.org = np.asarray(origins, dtype=np.float32)
        s = None if sem is None else np.ascontiguousarray(sem, dtype=np.int32)
        i = None if inst is None else np.ascontiguousarray(inst, dtype=np.int32)
        self._h = lib().ro_voxelize(_p(pts), pts.shape[0], _p(self.vs), _p(self.org), _p(s), _p(i))
        if not self._h:
            raise ValueError("empty point set")

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().ro_free(self._h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._h = None

    @property
    def total_voxels(self) -> int:
        return int(lib().ro_num_voxels(self._h))

    def info(self):
        a = [np.zeros(3, dtype=np.int32) for _ in range(4)]
        lib().ro_grid_info(self._h, *[_p(x) for x in a])
        return dict(imin=a[0], imax=a[1], bmin=a[2], bdim=a[3])

    def export(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        n = self.total_voxels
        ijk = np.zeros((n, 3), dtype=np.int32)
        sem = np.zeros(n, dtype=np.int32)
        inst = np.zeros(n, dtype=np.int32)
        lib().ro_export(self._h, _p(ijk), _p(sem), _p(inst))
        return ijk, sem, inst

    def render(self, kinv: np.ndarray, poses: np.ndarray, W: int, H: int, flat: bool = False):
        """get_zdepth_map_from_voxel + 2 x get_semantic_map_from_voxel (camera/base.py:520-618)."""
        kinv = np.ascontiguousarray(kinv, dtype=np.float32).reshape(9)
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 16)
        n = poses.shape[0]
        depth = np.zeros((n, H, W), dtype=np.float32)
        sem = np.zeros((n, H, W), dtype=np.int32)
        inst = np.zeros((n, H, W), dtype=np.int32)
        fn = lib().ro_render_flat if flat else lib().ro_render
        fn(self._h, _p(kinv), _p(poses), n, W, H, _p(depth), _p(sem), _p(inst))
        return depth, sem, inst

    def render_rows(self, kinv: np.ndarray, pose: np.ndarray, W: int, H: int, v0: int, v1: int, depth, sem, inst):
        kinv = np.ascontiguousarray(kinv, dtype=np.float32).reshape(9)
        pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(16)
        lib().ro_render_rows(self._h, _p(kinv), _p(pose), W, H, v0, v1, _p(depth), _p(sem), _p(inst))


# --------------------------------------------------------------------------------------------------
# camera (infinicube/camera/pinhole.py:23-138)
# --------------------------------------------------------------------------------------------------
def intrinsics_matrix(intr) -> torch.Tensor:
    fx, fy, cx, cy = [float(v) for v in intr[:4]]
    return torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)


def inv_intrinsics_matrix(intr) -> np.ndarray:
    """torch.inverse of the fp32 K, exactly as PinholeCamera caches it (pinhole.py:38-43,107-108)."""
    return torch.inverse(intrinsics_matrix(intr)).numpy()


def camera_rays(intr) -> np.ndarray:
    """PinholeCamera._get_rays_impl (pinhole.py:110-138) with the oracle's fixed operation order."""
    kinv = inv_intrinsics_matrix(intr)
    w, h = int(intr[4]), int(intr[5])
    u = np.arange(w, dtype=np.float32)[None, :].repeat(h, 0)
    v = np.arange(h, dtype=np.float32)[:, None].repeat(w, 1)
    rc = [(kinv[a, 0] * u + kinv[a, 1] * v) + kinv[a, 2] for a in range(3)]
    n = np.sqrt((rc[0] * rc[0] + rc[1] * rc[1]) + rc[2] * rc[2])
    return np.stack([c / n for c in rc], axis=-1).astype(np.float32)


# --------------------------------------------------------------------------------------------------
# semantic palette (infinicube/utils/semantic_utils.py:22-101)
# --------------------------------------------------------------------------------------------------
WAYMO_CATEGORY_NAMES = [
    "UNDEFINED", "CAR", "TRUCK", "BUS", "OTHER_VEHICLE", "MOTORCYCLIST", "BICYCLIST", "PEDESTRIAN", "SIGN",
    "TRAFFIC_LIGHT", "POLE", "CONSTRUCTION_CONE", "BICYCLE", "MOTORCYCLE", "BUILDING", "VEGETATION", "TREE_TRUNK",
    "CURB", "ROAD", "LANE_MARKER", "OTHER_GROUND", "WALKABLE", "SIDEWALK",
]
_VIS_TYPES = {
    0: ["SIGN", "TRAFFIC_LIGHT", "CONSTRUCTION_CONE"],
    1: ["MOTORCYCLIST", "BICYCLIST", "PEDESTRIAN", "BICYCLE", "MOTORCYCLE"],
    2: ["WALKABLE", "SIDEWALK"],
    3: ["CAR", "TRUCK", "BUS", "OTHER_VEHICLE"],
    4: ["VEGETATION", "TREE_TRUNK"],
    5: ["CURB", "LANE_MARKER"],
    6: ["BUILDING"],
    7: ["ROAD", "OTHER_GROUND"],
    8: ["UNDEFINED"],
    9: ["POLE"],
}
# ColorBrewer values used by the reference through pycg.color.get_cmap_array (= matplotlib listed colormaps)
_SET2 = ["66c2a5", "fc8d62", "8da0cb", "e78ac3", "a6d854", "ffd92f", "e5c494", "b3b3b3"]
_SET3_9, _SET3_10 = "bc80bd", "ccebc5"
_SET1_2 = "4daf4a"
_PAIRED_1 = "1f78b4"


def _hex(c: str) -> np.ndarray:
    # matplotlib stores ListedColormap colours as float64 r/255; the reference casts to float32
    return np.array([int(c[i:i + 2], 16) / 255.0 for i in (0, 2, 4)], dtype=np.float64)


def waymo_mapping_and_palette() -> Tuple[np.ndarray, np.ndarray]:
    mapping = np.zeros(23, dtype=np.int32)
    for idx, names in _VIS_TYPES.items():
        for n in names:
            mapping[WAYMO_CATEGORY_NAMES.index(n)] = idx
    pal = np.zeros((10, 3), dtype=np.float32)
    pal[:8] = np.stack([_hex(c) for c in _SET2])
    pal[3] = _hex(_SET3_9)
    pal[4] = _hex(_SET1_2)
    pal[8] = _hex(_PAIRED_1)
    pal[9] = _hex(_SET3_10)
    return mapping, pal


def semantic_palette_u8() -> np.ndarray:
    """label -> uint8 RGB: (semantic_to_color(l) * 255).astype(uint8), truncation
    (semantic_utils.py:88-101, guidance_buffer_generation.py:693-695)."""
    mapping, pal = waymo_mapping_and_palette()
    return (pal[mapping] * 255).astype(np.uint8)


def semantic_rgb(sem: np.ndarray, inst: np.ndarray, inst_ids: np.ndarray, inst_colors_u8: np.ndarray) -> np.ndarray:
    """generate_rgb_semantic_buffer (semantic_utils.py:104-131) with injected instance colours."""
    pal = np.ascontiguousarray(semantic_palette_u8())
    s = np.ascontiguousarray(sem, dtype=np.int32).reshape(-1)
    i = np.ascontiguousarray(inst, dtype=np.int32).reshape(-1)
    ids = np.ascontiguousarray(inst_ids, dtype=np.int32)
    cols = np.ascontiguousarray(inst_colors_u8, dtype=np.uint8)
    out = np.zeros((s.shape[0], 3), dtype=np.uint8)
    lib().ro_semantic_rgb(_p(s), _p(i), s.shape[0], _p(pal), _p(ids), _p(cols), ids.shape[0], _p(out))
    return out.reshape(*sem.shape, 3)


# --------------------------------------------------------------------------------------------------
# coordinate buffer (infinicube/utils/buffer_utils.py:180-265, utils/depth_utils.py:402-466)
# --------------------------------------------------------------------------------------------------
def unproject_to_cam0(depth: np.ndarray, intr, poses: np.ndarray) -> np.ndarray:
    """X_cam0 = T_0^-1 T_i [depth * K^-1 (u,v,1); 1], fixed fp32 operation order; misses -> 1e7."""
    kinv = inv_intrinsics_matrix(intr)
    P = torch.from_numpy(np.asarray(poses, dtype=np.float32))
    c2c0 = torch.einsum("ij,bjk->bik", torch.inverse(P[0]), P).numpy()  # buffer_utils.py:212-217
    n, h, w = depth.shape
    u = np.arange(w, dtype=np.float32)[None, None, :]
    v = np.arange(h, dtype=np.float32)[None, :, None]
    cp = [depth * ((kinv[a, 0] * u + kinv[a, 1] * v) + kinv[a, 2]) for a in range(3)]
    out = np.empty((n, h, w, 3), dtype=np.float32)
    for a in range(3):
        T = c2c0[:, a, :][:, None, None, :]
        out[..., a] = ((T[..., 0] * cp[0] + T[..., 1] * cp[1]) + T[..., 2] * cp[2]) + T[..., 3]
    out[depth == 0] = 1e7
    return out


def sample_quantiles(xyz: np.ndarray, percentile: float = 0.05, seed: Optional[int] = None):
    """buffer_utils.py:232-249: quantiles over <= 100k randperm-sampled valid points."""
    flat = torch.from_numpy(xyz.reshape(-1, 3))
    valid = flat[flat[:, 2] < 1e6]
    if seed is not None:
        torch.manual_seed(seed)
    sample = valid[torch.randperm(valid.shape[0])[:100000]]
    mins = torch.quantile(sample, percentile, dim=0)
    maxs = torch.quantile(sample, 1 - percentile, dim=0)
    ranges = torch.clamp(maxs - mins, min=1e-7)
    return mins.numpy(), ranges.numpy()


def coordinate_buffer(xyz: np.ndarray, depth: np.ndarray, mins: np.ndarray, ranges: np.ndarray):
    """buffer_utils.py:251-262 then (x*255).astype(uint8) (guidance_buffer_generation.py:710)."""
    q = (xyz - mins) / ranges * np.float32(2.0) - np.float32(1.0)
    q = np.clip(q, -1.0, 1.0).astype(np.float32)
    c = ((q + np.float32(1.0)) / np.float32(2.0)).astype(np.float32)
    c[depth == 0] = 1.0
    return c, (c * 255).astype(np.uint8)
