"""CPU restatement of the CAD-mesh voxelisation — TEST INFRASTRUCTURE ONLY (imported by tests/ only).

Follows infinicube/utils/fvdb_utils.py:219-296 (scale the mesh to each box's lwh, gridbatch_from_mesh at 0.1 m with
origin 0.05, return voxel centres).  PARITY UNPINNED for the triangle/voxel rule itself: it lives in the absent
fvdb==0.2.0 wheel; restated here as "the closed voxel cube overlaps the closed triangle" (exact separating-axis
test in fp64), written independently of the CUDA kernel: this version clips nothing and tests every voxel of the
triangle's bounding box with vectorised numpy."""
from __future__ import annotations

import numpy as np


def _overlap(c: np.ndarray, h: float, tri: np.ndarray) -> np.ndarray:
    """c [N,3] cube centres, tri [3,3] -> bool [N]"""
    v = tri[None, :, :] - c[:, None, :]            # [N, 3 verts, 3]
    e = np.stack([tri[1] - tri[0], tri[2] - tri[1], tri[0] - tri[2]])  # [3,3]
    ok = np.ones(c.shape[0], dtype=bool)
    eye = np.eye(3)
    for j in range(3):
        for i in range(3):
            ax = np.cross(eye[i], e[j])
            pr = v @ ax                               # [N,3]
            rad = h * np.abs(ax).sum()
            ok &= ~((pr.min(1) > rad) | (pr.max(1) < -rad))
    for a in range(3):
        ok &= ~((v[:, :, a].min(1) > h) | (v[:, :, a].max(1) < -h))
    n = np.cross(e[0], e[1])
    d = v[:, 0, :] @ n
    r = h * np.abs(n).sum()
    ok &= ~((d > r) | (d < -r))
    return ok


def voxelize_mesh(vertices: np.ndarray, faces: np.ndarray, voxel_size: float = 0.1, origin: float = 0.05) -> np.ndarray:
    v = np.asarray(vertices, dtype=np.float64)
    out = set()
    for f in np.asarray(faces):
        tri = v[f]
        lo = np.floor((tri.min(0) - origin) / voxel_size).astype(int) - 1
        hi = np.ceil((tri.max(0) - origin) / voxel_size).astype(int) + 1
        g = np.stack(np.meshgrid(*[np.arange(lo[a], hi[a] + 1) for a in range(3)], indexing="ij"), -1).reshape(-1, 3)
        c = origin + g * voxel_size
        m = _overlap(c, 0.5 * voxel_size, tri)
        out.update(map(tuple, g[m]))
    arr = np.array(sorted(out, key=lambda t: (t[2], t[1], t[0])), dtype=np.int32).reshape(-1, 3)
    return arr
