"""Rasteriser over several GPUs: cameras are independent units, so the poses are split across the ranks
(93 -> 12/12/12/12/12/11/11/11 on eight), every rank keeps the whole voxel grid and renders its share with the
one-launch kernel, and the three images per camera are gathered.  No collective touches the data path of the
render itself; the gather is a plain `all_gather` of equal-sized (padded) blocks."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_cameras(n_cam: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first camera, count) of `rank`: contiguous blocks, the first n_cam % world_size ranks hold one more."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_cam, world_size)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def gather_camera_shards(local: Sequence[torch.Tensor], n_cam: int, world_size: int, rank: int,
                         group=None) -> List[torch.Tensor]:
    """all-gather per-camera tensors `[count, ...]` into `[n_cam, ...]` in camera order.  Shards differ by at most one
    camera, so every rank pads to the largest count and the padding rows are dropped after the gather."""
    if world_size == 1:
        return [t for t in local]
    import torch.distributed as dist
    max_count = -(-n_cam // world_size)
    out = []
    for t in local:
        start, count = shard_cameras(n_cam, world_size, rank)
        if t.shape[0] != count:
            raise ValueError(f"rank {rank} holds {t.shape[0]} cameras, its shard is {count}")
        block = t if count == max_count else torch.cat([t, t.new_zeros((max_count - count,) + tuple(t.shape[1:]))])
        parts = [torch.empty_like(block) for _ in range(world_size)]
        dist.all_gather(parts, block.contiguous(), group=group)
        out.append(torch.cat([parts[r][: shard_cameras(n_cam, world_size, r)[1]] for r in range(world_size)]))
    return out


def render_voxel_buffers_sharded(camera, camera_poses, voxel_grid, world_size: int, rank: int, gather: bool = True,
                                 group=None, attr0: Optional[torch.Tensor] = None, attr1: Optional[torch.Tensor] = None,
                                 background0: int = 0, background1: int = 0):
    """`PinholeCamera.render_voxel_buffers` with the cameras split over `world_size` ranks.  Returns (depth,
    attr0 image, attr1 image) for all cameras (`gather=True`, identical on every rank) or for this rank's cameras."""
    poses = torch.as_tensor(camera_poses)
    n_cam = poses.shape[0]
    start, count = shard_cameras(n_cam, world_size, rank)
    local = camera.render_voxel_buffers(poses[start:start + count], voxel_grid, attr0=attr0, attr1=attr1,
                                        background0=background0, background1=background1) if count else None
    if local is None:   # more ranks than cameras: an empty shard still takes part in the gather
        dev = voxel_grid.device
        local = (torch.empty((0, camera.h, camera.w), dtype=torch.float32, device=dev),
                 torch.empty((0, camera.h, camera.w), dtype=torch.int32, device=dev),
                 torch.empty((0, camera.h, camera.w), dtype=torch.int32, device=dev))
    if not gather:
        return local
    return tuple(gather_camera_shards(local, n_cam, world_size, rank, group))
