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
                # one cached generator per process, exactly like the reference (:755-768).  INFINICUBE_B200_WORLD_SIZE=N
                # makes that one object drive N GPUs: it spawns and owns ranks 1..N-1 (videogen/multiproc.py), so this
                # script needs neither torchrun nor any change to use the whole box.
                import os
                world = int(os.environ.get("INFINICUBE_B200_WORLD_SIZE", "1"))
                generate_guidance_buffer_and_save._video_generator = WanVideoGenerator(
                    checkpoint_path=video_checkpoint_path, device="cuda:0", torch_dtype=torch.bfloat16, buffer_channels=16,
                    enable_vram_management=True, use_wan_1pt3b=use_wan_1pt3b, world_size=world)
            generator = generate_guidance_buffer_and_save._video_generator
            out = output_folder / f"video_{resolution}_front.mp4"
            generator.generate(semantic_buffer=sem_frames[:93], coordinate_buffer=coord_frames[:93], prompt=video_prompt,
                               seed=0, tiled=True, output_path=str(out), fps=10, quality=8)
        except Exception as e:  # noqa: BLE001 - the reference logs and continues (guidance_buffer_generation.py:786-791)
            print(f"Failed to generate video: {e}")
            print("Continuing without video generation...")
            traceback.print_exc()
    return depth, semantic, instance


# --------------------------------------------------------------------------------------------------
# voxel-world wire format
# --------------------------------------------------------------------------------------------------
def save_voxel_npz(path, ijk, semantics, voxel_size=0.2, origin=0.1) -> Path:
    """Neutral voxel-world file: {ijk int32 (N,3), semantics int64 (N,), voxel_size f64 (3,), origin f64 (3,)}.
    Replaces the pickled fvdb.GridBatch written by stage 1 (voxel_world_generation.py:849-856), which can only be
    unpickled where fVDB is installed."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    ijk = np.ascontiguousarray(torch.as_tensor(ijk).cpu().numpy().astype(np.int32)).reshape(-1, 3)
    sem = np.ascontiguousarray(torch.as_tensor(semantics).cpu().numpy().astype(np.int64)).reshape(-1)
    if sem.shape[0] != ijk.shape[0]:
        raise ValueError("semantics and ijk differ in length")
    vs = np.broadcast_to(np.asarray(voxel_size, np.float64), (3,)).copy()
    og = np.broadcast_to(np.asarray(origin, np.float64), (3,)).copy()
    with open(path, "wb") as f:
        np.savez(f, ijk=ijk, semantics=sem, voxel_size=vs, origin=og)
    return path


def convert_fvdb_voxel_to_npz(pt_path, npz_path=None) -> Path:
    """Run where fVDB exists: `{step}.pt` (pickled GridBatch + semantics) -> neutral `.npz` next to it."""
    pt_path = Path(pt_path)
    voxel = torch.load(pt_path, weights_only=False)
    grid = voxel["points"]
    vs = grid.voxel_sizes[0].cpu().numpy().astype(np.float64)
    og = grid.origins[0].cpu().numpy().astype(np.float64)
    return save_voxel_npz(npz_path or pt_path.with_suffix(".npz"), grid.ijk.jdata, voxel["semantics"], vs, og)


def _voxel_step(p: Path) -> int:
    return int(p.stem)


def select_voxel_file(voxel_root, clip, extrap_voxel_time=None) -> Path:
    """File choice of load_voxel (guidance_buffer_generation.py:446-456): highest-numbered step unless one is given;
    the neutral `.npz` wins over a `.pt` of the same step."""
    root = Path(voxel_root).resolve() / clip
    if extrap_voxel_time is None:
        files = [p for p in list(root.glob("*.npz")) + list(root.glob("*.pt")) if p.stem.isdigit()]
        if not files:
            raise FileNotFoundError(f"no voxel files under {root}")
        top = max(_voxel_step(p) for p in files)
        cands = [p for p in files if _voxel_step(p) == top]
    else:
        cands = [root / f"{extrap_voxel_time}.npz", root / f"{extrap_voxel_time}.pt"]
    cands = sorted((p for p in cands if p.exists()), key=lambda p: p.suffix != ".npz")
    if not cands:
        raise FileNotFoundError(f"Voxel {root / str(extrap_voxel_time)}.[npz|pt] does not exist")
    return cands[0]


def read_voxel_file(voxel_path):
    """-> (voxel centres in world space (N,3) fp32, semantics (N,) int64), both on the host."""
    voxel_path = Path(voxel_path)
    if voxel_path.suffix == ".npz":
        with np.load(voxel_path) as z:
            ijk, sem = z["ijk"], z["semantics"]
            vs, og = z["voxel_size"].astype(np.float32), z["origin"].astype(np.float32)
        pts = torch.from_numpy(ijk.astype(np.float32)) * torch.from_numpy(vs) + torch.from_numpy(og)
        semantics = torch.from_numpy(sem)
    else:
        try:
            voxel = torch.load(voxel_path, weights_only=False)
        except (ModuleNotFoundError, AttributeError) as e:
            raise RuntimeError(f"{voxel_path} pickles an fvdb.GridBatch and fVDB is not installed here; convert it "
                               f"with convert_fvdb_voxel_to_npz where fVDB exists") from e
        p = voxel["points"]
        if isinstance(p, dict):
            vs = torch.as_tensor(p["voxel_size"], dtype=torch.float32).expand(3)
            og = torch.as_tensor(p["origin"], dtype=torch.float32).expand(3)
            pts = torch.as_tensor(p["ijk"]).float() * vs + og
        elif torch.is_tensor(p):
            pts = p.float()
        else:   # a live GridBatch (fVDB present): voxel centres in world space
            pts = p.grid_to_world(p.ijk.float()).jdata.float()
        semantics = torch.as_tensor(voxel["semantics"])
    if semantics.shape[0] != pts.shape[0]:
        raise ValueError(f"{voxel_path}: {pts.shape[0]} voxels but {semantics.shape[0]} semantics")
    return pts.cpu(), semantics.long().cpu()


def load_voxel(voxel_root, clip, extrap_voxel_time=None):
    """Mirror of load_voxel (guidance_buffer_generation.py:431-462): returns (scene, semantics int64 cuda, path).
    `scene` is the (N,3) fp32 tensor of voxel centres in world space - the tensor form
    generate_infinicube_buffer_from_fvdb_grid accepts (utils/fvdb_utils.py:492-497).  Reads the neutral `.npz`; a
    `.pt` holding plain tensors ({"points": (N,3) float | {"ijk","voxel_size","origin"}, "semantics"}) is accepted
    too; a `.pt` that pickles an fvdb.GridBatch needs fVDB and is converted with convert_fvdb_voxel_to_npz."""
    voxel_path = select_voxel_file(voxel_root, clip, extrap_voxel_time)
    pts, semantics = read_voxel_file(voxel_path)
    return pts.cuda(), semantics.cuda(), voxel_path
