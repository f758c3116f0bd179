"""Stage-3 entry: reads what stage 2 wrote - the generated video, poses, intrinsics, depth x 100 and
instance buffers, dynamic object info - into the dictionary `scene_gaussian_generation.py` feeds its reconstruction
model.  Mirror of `get_data_dict_from_folder`, `_determine_key_frame_indices`, `_compute_sky_and_foreground_masks`
(depth-buffer branch) and `_create_gsm_input_masks`
(infinicube/inference/scene_gaussian_generation.py:258-372, 375-404, 407-432, 435-477).

Host-side file and mask plumbing only; the GSM reconstruction and the sky-segmentation network it may call are out
of scope (the reference itself falls back to the depth buffer when segmentation is unavailable, :411-424).  The
point of this module is that the folder our stage 2 writes is consumed unchanged by a stage-3 reader.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from ..utils_io import get_sample, read_video_file

DEPTH_SCALE_FACTOR = 100.0          # scene_gaussian_generation.py:66
DYNAMIC_INSTANCE_ID_START = 10000   # "dynamic objects are counting from 10000" (:324)


def determine_key_frame_indices(args, data_folder: Path, total_frames: int) -> List[int]:
    """key_frame_indices.json > meta.json > command-line arguments (:375-404)."""
    data_folder = Path(data_folder)
    if (data_folder / "key_frame_indices.json").exists():
        return [int(i) for i in json.load(open(data_folder / "key_frame_indices.json"))]
    if (data_folder / "meta.json").exists():
        meta = json.load(open(data_folder / "meta.json"))
        end_frame = int(meta["active_frame_proportion"] * total_frames)
        return list(range(int(meta["start_frame_index"]), end_frame, int(meta["use_frame_interval"])))
    start = int(args.start_frame_index)
    end = min(start + int(args.active_frame_proportion * total_frames), total_frames)
    return list(range(start, end, args.use_frame_interval))


def compute_sky_and_foreground_masks(data_dict: Dict, sky_segmenter: Optional[Callable] = None) -> Dict:
    """foreground_mask_from_seg / _from_grid (:407-432).  `sky_segmenter(video [N,H,W,3] float) -> bool [N,H,W]` is
    optional; without one the depth buffer alone decides, exactly the reference's own fallback."""
    sky_from_seg = None
    if sky_segmenter is not None:
        try:
            sky_from_seg = torch.as_tensor(np.asarray(sky_segmenter(data_dict["video_array"]))).bool()
        except Exception as e:  # noqa: BLE001 - the reference catches everything here too
            print(f"Sky segmentation failed: {e}. Using depth buffer only.")
    if sky_from_seg is None:
        sky_from_seg = torch.zeros_like(data_dict["depth_buffers"], dtype=torch.bool)
    sky_from_grid = data_dict["depth_buffers"] == 0          # a ray that hit no voxel
    data_dict["foreground_mask_from_seg"] = ~sky_from_seg
    data_dict["foreground_mask_from_grid"] = ~sky_from_grid
    return data_dict


def create_gsm_input_masks(data_dict: Dict, args) -> Dict:
    """4-channel input mask (:435-477): 0 = foreground from segmentation, 1 = non-dynamic, 2 = ones, 3 = foreground
    from the depth grid; before the last `enable_pixel_branch_last_n_frame` frames channel 0 := channel 3."""
    shape = (*data_dict["video_array"].shape[:3], 4)
    m = torch.ones(shape).to(data_dict["video_array"])
    m[..., 0] = data_dict["foreground_mask_from_seg"]
    m[..., 3] = data_dict["foreground_mask_from_grid"]
    m[..., 1] = 1.0
    n_frames = int(getattr(args, "enable_pixel_branch_last_n_frame", 0))
    if n_frames > 0:
        m[:-n_frames, ..., 0] = m[:-n_frames, ..., 3]
    else:
        m[..., 0] = m[..., 3]
    data_dict["gsm_images_input_mask"] = m
    return data_dict


def get_data_dict_from_folder(args, resolution: str = "480p", sky_segmenter: Optional[Callable] = None) -> Dict:
    """args: data_folder, start_frame_index, active_frame_proportion, use_frame_interval,
    enable_pixel_branch_last_n_frame (the reference's argparse names, :200-255)."""
    data_folder = Path(args.data_folder)
    data_dict: Dict = {}

    # 1) voxel world of pass 0 (neutral .npz written by our stage 2, or a tensor .pt)
    for name in ("voxel.npz", "voxel.pt"):
        if (data_folder / name).exists():
            from .guidance_buffer_generation import read_voxel_file
            pts, sem = read_voxel_file(data_folder / name)
            data_dict["scene_grid"], data_dict["scene_semantic"] = pts, sem
            break

    # 2) poses
    pose_sample = get_sample(data_folder / "pose.tar")
    pose_keys = sorted(k for k in pose_sample if "pose.front.npy" in k)
    poses = torch.from_numpy(np.stack([pose_sample[k] for k in pose_keys])).float()

    # 3) the generated video, [N, H, W, 3] in [0, 1]
    video_array = torch.from_numpy(read_video_file(data_folder / f"video_{resolution}_front.mp4")) / 255.0

    # 4) intrinsics [fx, fy, cx, cy, w, h]
    intrinsics = torch.from_numpy(get_sample(data_folder / "intrinsic.tar")["intrinsic.front.npy"]).float()

    # 5) depth buffers: 16-bit PNG of depth x 100 -> metres
    depth_sample = get_sample(data_folder / f"voxel_depth_100_{resolution}_front.tar")
    depth_keys = sorted(k for k in depth_sample if ".voxel_depth_100.front.png" in k)
    depth_buffers = torch.from_numpy(np.stack([depth_sample[k] / DEPTH_SCALE_FACTOR for k in depth_keys])).float()

    # 6) instance buffers -> dynamic masks
    inst_sample = get_sample(data_folder / f"instance_buffer_{resolution}_front.tar")
    inst_keys = sorted(k for k in inst_sample if ".instance_buffer.front.png" in k)
    instance_buffers = torch.from_numpy(np.stack([inst_sample[k] for k in inst_keys]).astype(np.int32))
    dynamic_masks = instance_buffers >= DYNAMIC_INSTANCE_ID_START
    non_dynamic_masks = ~dynamic_masks

    # 7) dynamic object info, one json per frame
    info_file = data_folder / "dynamic_object_info.tar"
    info_sample = get_sample(info_file)
    info_keys = sorted(k for k in info_sample if ".dynamic_object_info.json" in k)
    dynamic_object_infos = [info_sample[k] for k in info_keys]

    n_video = video_array.shape[0]
    key = determine_key_frame_indices(args, data_folder, n_video)
    if key and max(key) >= min(n_video, poses.shape[0], depth_buffers.shape[0], instance_buffers.shape[0]):
        raise IndexError(f"key frame {max(key)} beyond the {n_video} video / {poses.shape[0]} pose / "
                         f"{depth_buffers.shape[0]} depth frames in {data_folder}")
    data_dict.update({
        "key_frame_indices": key,
        "poses": poses[key],
        "intrinsics": intrinsics.expand(len(key), -1),
        "video_array": video_array[key],
        "depth_buffers": depth_buffers[key],
        "dynamic_masks": dynamic_masks[key],
        "non_dynamic_masks": non_dynamic_masks[key],
        "dynamic_object_infos": [dynamic_object_infos[k] for k in key] if dynamic_object_infos else [],
        "original_dynamic_object_info_file": info_file,
    })
    data_dict = compute_sky_and_foreground_masks(data_dict, sky_segmenter)
    return create_gsm_input_masks(data_dict, args)
