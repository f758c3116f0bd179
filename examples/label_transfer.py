"""Stage-1 chunk merge on the GPU: move a voxel world rigidly, re-voxelise it and carry the semantic labels over with
the exact nearest-neighbour search (the reference's `transform_grid_and_semantic` recipe, extrap_util.py:233-276, with
`semantic_from_points` from this package).  Needs a B200.

    python examples/label_transfer.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from infinicube_b200.raster import VoxelGrid, synthetic as syn  # noqa: E402
from infinicube_b200.voxelgen import semantic_from_points  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    vs = 0.2
    pts, sem, _, _ = syn.synthetic_scene(128, voxel_size=vs)
    xyz = torch.from_numpy(pts).to(dev)
    labels = torch.from_numpy(sem).to(dev).long()
    # rigid motion of the chunk (double precision like the reference, then back to fp32)
    a = 0.1
    T = torch.tensor([[np.cos(a), -np.sin(a), 0, 1.3], [np.sin(a), np.cos(a), 0, -0.4], [0, 0, 1, 0.05], [0, 0, 0, 1]],
                     dtype=torch.float64, device=dev)
    moved = (xyz.double() @ T[:3, :3].T + T[:3, 3]).float()
    # re-voxelise at the same voxel size (origin = half a voxel, as fvdb.gridbatch_from_points is called there)
    grid = VoxelGrid(moved, [vs] * 3, [vs / 2] * 3)
    centres = grid.grid_to_world(grid.ijk)          # voxel centres of the new grid, world space
    new_labels = semantic_from_points(centres, moved, labels, cell_size=2 * vs)
    print(f"{xyz.shape[0]} voxels moved -> {centres.shape[0]} voxels; label histogram",
          torch.bincount(new_labels, minlength=int(labels.max()) + 1).tolist())


if __name__ == "__main__":
    main()
