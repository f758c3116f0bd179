"""Stage-2 orchestration with the reference's surface: rasterise -> colour -> coordinate buffer -> tars / mp4s ->
video (mirrors generate_guidance_buffer_and_save, infinicube/inference/guidance_buffer_generation.py:591-791).
All per-pixel work runs in the CUDA library; buffers stay on the GPU until the files are written."""
from __future__ import annotations

import traceback
from pathlib import Path

import numpy as np
import torch

from ..raster import generate_infinicube_buffer_from_fvdb_grid
from ..raster.buffer_utils import coordinate_buffer
from ..raster.semantic_utils import semantic_rgb_u8
from ..utils_io import encode_png, vis_depth, write_to_tar, write_video_file


def generate_guidance_buffer_and_save(clip, output_folder, resolution, camera_model, camera_poses, fvdb_grid, fvdb_semantic,
                                      static_object_info, dynamic_object_info, video_prompt, disable_video_generation,
                                      video_checkpoint_path, use_wan_1pt3b, cad_model_location=None, rng=None):
    output_folder = Path(output_folder)
    grid_to_world = torch.eye(4, device="cuda")
    depth, semantic, instance = generate_infinicube_buffer_from_fvdb_grid(
        camera_model=camera_model, camera_poses_in_world=camera_poses, fvdb_scene_grid_or_points=fvdb_grid,
        fvdb_scene_semantic=fvdb_semantic, fvdb_grid_to_world=grid_to_world, static_object_info=static_object_info,
        dynamic_object_info=dynamic_object_info, cad_model_for_dynamic_objects=True, cad_model_for_static_object=True,
        cad_model_location=cad_model_location, enlarge_lwh_factor=1.2)
    n = depth.shape[0]
    # guidance images on the GPU (uint8), then one D->H copy each for the files
    sem_rgb = semantic_rgb_u8(semantic, instance, rng=rng)
    _, coord_u8 = coordinate_buffer(depth, camera_model, camera_poses.detach().cpu(), percentile=0.05, want_f32=False,
                                    want_u8=True)
    depth_np = depth.cpu().numpy()
    inst_np = instance.cpu().numpy().astype(np.uint16)
    poses_np = camera_poses.detach().cpu().numpy()
    depth_sample, instance_sample, pose_sample, depth_vis_frames = {}, {}, {}, []
    for i in range(n):
        depth_sample[f"{i:06d}.voxel_depth_100.front.png"] = encode_png((depth_np[i] * 100).astype(np.uint16))
        instance_sample[f"{i:06d}.instance_buffer.front.png"] = encode_png(inst_np[i])
        depth_vis_frames.append(vis_depth(depth_np[i]))
        pose_sample[f"{i:06d}.pose.front.npy"] = poses_np[i]
    write_to_tar(depth_sample, output_folder / f"voxel_depth_100_{resolution}_front.tar", __key__=clip)
    write_to_tar(instance_sample, output_folder / f"instance_buffer_{resolution}_front.tar", __key__=clip)
    write_to_tar(pose_sample, output_folder / "pose.tar", __key__=clip)
    write_to_tar({"intrinsic.front.npy": camera_model.intrinsics}, output_folder / "intrinsic.tar", __key__=clip)
    write_to_tar(dict(dynamic_object_info or {}), output_folder / "dynamic_object_info.tar", __key__=clip)
    sem_frames = sem_rgb.cpu().numpy()
    coord_frames = coord_u8.cpu().numpy()
    write_video_file(sem_frames, output_folder / f"semantic_buffer_video_{resolution}_front.mp4", fps=10)
    write_video_file(depth_vis_frames, output_folder / f"depth_vis_video_{resolution}_front.mp4", fps=10)
    write_video_file(coord_frames, output_folder / f"coordinate_buffer_video_{resolution}_front.mp4", fps=10)
    print(f"Saved guidance buffer to {output_folder}")

    if not disable_video_generation:
        try:
            from ..videogen import WanVideoGenerator
            if not hasattr(generate_guidance_buffer_and_save, "_video_generator"):
                generate_guidance_buffer_and_save._video_generator = WanVideoGenerator(
                    checkpoint_path=video_checkpoint_path, device="cuda:0", torch_dtype=torch.bfloat16, buffer_channels=16,
                    enable_vram_management=True, use_wan_1pt3b=use_wan_1pt3b)
            generator = generate_guidance_buffer_and_save._video_generator
            out = output_folder / f"video_{resolution}_front.mp4"
            generator.generate(semantic_buffer=sem_frames[:93], coordinate_buffer=coord_frames[:93], prompt=video_prompt,
                               seed=0, tiled=True, output_path=str(out), fps=10, quality=8)
        except Exception as e:  # noqa: BLE001 - the reference logs and continues (guidance_buffer_generation.py:786-791)
            print(f"Failed to generate video: {e}")
            print("Continuing without video generation...")
            traceback.print_exc()
    return depth, semantic, instance
