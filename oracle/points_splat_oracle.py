"""TEST / MEASUREMENT INFRASTRUCTURE - not part of the product path.

CPU restatement of the reference's own points-to-depth rasteriser,
`CameraBase.get_zdepth_map_from_points_torch` (infinicube/camera/base.py:386-447), the nearest reference-AUTHORED CPU
path to the voxel ray march (whose real implementation lives in the absent fVDB wheel):
every point is transformed into the camera (`transform_points_torch`, base.py:229-238), projected with the pinhole
intrinsics (`ray2pixel_torch`, pinhole.py:157-172), rounded to the nearest pixel and z-buffered with
`scatter_reduce_(amin)`; pixels nobody hit read 0.

Pinned: `tests/test_oracle_raster.py::test_points_splat_matches_reference_output` compares it with outputs of the
reference function itself (tests/golden/points_splat.npz, written by `python oracle/points_splat_oracle.py --golden`
in the build container where /root/reference exists).  `tools/raster_sweep.py` times it on the GPU box's host cores
next to the GPU ray march.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch


def zdepth_map_from_points(intrinsics6, camera_poses: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """intrinsics6 = [fx, fy, cx, cy, w, h]; camera_poses (N,4,4) camera->world; points (M,3) world -> (N,H,W)."""
    fx, fy, cx, cy, w, h = [float(v) for v in intrinsics6]
    w, h = int(w), int(h)
    K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)
    poses = camera_poses.to(torch.float32)
    pts = points.to(torch.float32)
    out = []
    for cam_to_world in poses:
        tfm = torch.inverse(cam_to_world)
        pc = (tfm[:3, :3] @ pts.T + tfm[:3, 3].unsqueeze(-1)).T          # base.py:237-238
        rays_norm = pc / pc[:, 2:3]                                        # pinhole.py:167
        uv = torch.einsum("ij,nj->ni", K, rays_norm)[:, :2]                # pinhole.py:168-170
        depth = pc[:, 2]
        u = torch.round(uv[:, 0]).long()
        v = torch.round(uv[:, 1]).long()
        ok = (depth > 0) & (u >= 0) & (u < w) & (v >= 0) & (v < h)
        idx = v[ok] * w + u[ok]
        img = torch.full((h * w,), float("inf"), dtype=torch.float32)
        img = img.scatter_reduce_(0, idx, depth[ok], "amin")
        img[~torch.isfinite(img)] = 0
        out.append(img.view(h, w))
    return torch.stack(out, 0)


def _golden():
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from gen_golden_from_reference import REF, _install_stubs
    sys.path.insert(0, str(REF))
    _install_stubs()
    from infinicube.camera.pinhole import PinholeCamera
    g = torch.Generator().manual_seed(21)
    intr = np.array([60.0, 55.0, 31.5, 19.5, 64, 40], dtype=np.float64)
    cam = PinholeCamera.from_numpy(intr, device=torch.device("cpu"))
    poses = torch.eye(4).repeat(3, 1, 1)
    for i in range(3):
        q, _ = torch.linalg.qr(torch.eye(3) + 0.2 * torch.randn(3, 3, generator=g))
        poses[i, :3, :3] = q
        poses[i, :3, 3] = torch.randn(3, generator=g)
    pts = torch.randn(6000, 3, generator=g) * torch.tensor([4.0, 3.0, 1.0]) + torch.tensor([0.0, 0.0, 8.0])
    pts = torch.round(pts / 0.2) * 0.2 + 0.1
    ref = cam.get_zdepth_map_from_points_torch(poses, pts)
    out = Path(__file__).resolve().parent.parent / "tests" / "golden" / "points_splat.npz"
    np.savez_compressed(out, intr=intr, poses=poses.numpy(), points=pts.numpy(), depth=ref.numpy())
    mine = zdepth_map_from_points(intr, poses, pts)
    print("wrote", out, "hit fraction", float((ref > 0).float().mean()), "restatement equal:", bool(torch.equal(mine, ref)))


if __name__ == "__main__":
    if "--golden" in sys.argv:
        _golden()
