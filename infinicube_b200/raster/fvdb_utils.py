"""Guidance-buffer generation with the reference's surface
(infinicube/utils/fvdb_utils.py:71-216, 299-618).  Grid construction, label arg-max and the fused
depth / semantic / instance ray march run in csrc/raster.cu."""
from __future__ import annotations

from typing import Dict, Optional, Union

import numpy as np
import torch

from .camera import PinholeCamera
from .grid import VoxelGrid
from .semantic_utils import WAYMO_CATEGORY_NAMES

_CAR_LIKE = [WAYMO_CATEGORY_NAMES.index(n) for n in ("CAR", "TRUCK", "BUS", "OTHER_VEHICLE")]


def keep_car_only_in_object_info(object_info_all_frames: Optional[Dict]) -> Dict:
    """infinicube/data_process/waymo_utils.py:271-302"""
    out = {}
    for key, frame in (object_info_all_frames or {}).items():
        if key.endswith(".json"):
            out[key] = {gid: info for gid, info in frame.items() if info["object_type"] == "car"}
        else:
            out[key] = frame
    return out


def points_to_fvdb(points: torch.Tensor, points_to_world, attrs: Optional[Dict[str, torch.Tensor]] = None,
                   voxel_sizes=(0.1, 0.1, 0.1), origins=(0.05, 0.05, 0.05), bound_min=None, bound_max=None,
                   cc_removal=False, cc_removal_min=-1, extra_meshes=None):
    """points -> (VoxelGrid, per-voxel attributes); `semantics` and `instance` reduce by arg-max count
    with ties to the smallest label (fvdb_utils.py:105-109,174-191)."""
    if cc_removal:
        raise NotImplementedError("connected-component removal is not on the guidance-buffer path")
    xyz, att = points, attrs
    if bound_min is not None or bound_max is not None:
        lo = bound_min if bound_min is not None else [-1e7] * 3
        hi = bound_max if bound_max is not None else [1e7] * 3
        m = ((points[:, 0] >= lo[0]) & (points[:, 0] < hi[0]) & (points[:, 1] >= lo[1]) & (points[:, 1] < hi[1])
             & (points[:, 2] >= lo[2]) & (points[:, 2] < hi[2]))
        xyz = points[m]
        att = {k: v[m] for k, v in attrs.items()} if attrs is not None else None
    unknown = set(att or {}) - {"semantics", "instance"}
    if unknown:
        raise NotImplementedError(f"Reduce strategy for {sorted(unknown)} is not implemented.")
    grid = VoxelGrid(xyz, voxel_sizes, origins, (att or {}).get("semantics"), (att or {}).get("instance"))
    out = {}
    if att is not None:
        if "semantics" in att:
            out["semantics"] = grid.semantics.to(att["semantics"].dtype)
        if "instance" in att:
            out["instance"] = grid.instance.to(att["instance"].dtype)
    out["grid_to_world"] = points_to_world
    return grid, out


def pack_instance_boxes(static_object_info: Dict, enlarge_lwh_factor: float = 1.0) -> torch.Tensor:
    """Frame-0 static boxes -> the [n_boxes, 16] fp32 table ic_instance_from_boxes reads: rows 0-2 of
    torch.inverse(object_to_world) (fp32, like fvdb_utils.py:356), lwh / 2 * factor (fvdb_utils.py:370), and
    object_id_int bit-cast into the last float.  Dict order is preserved: the later box wins (fvdb_utils.py:378)."""
    boxes = (static_object_info or {}).get("000000.static_object_info.json", {})
    tab = torch.zeros(len(boxes), 16, dtype=torch.float32)
    for b, data in enumerate(boxes.values()):
        o2w = torch.tensor(data["object_to_world"], dtype=torch.float32)
        tab[b, :12] = torch.inverse(o2w)[:3].reshape(-1)
        tab[b, 12:15] = torch.tensor(data["object_lwh"], dtype=torch.float32) / 2.0 * enlarge_lwh_factor
        tab[b, 15] = torch.tensor([int(data["object_id_int"])], dtype=torch.int32).view(torch.float32)[0]
    return tab


def get_instance_id_for_fvdb_scene_points(points_in_world: torch.Tensor, semantic: torch.Tensor,
                                          static_object_info: Dict, enlarge_lwh_factor: float = 1.0) -> torch.Tensor:
    """Car-class points inside an (enlarged) frame-0 static box take its object_id_int (fvdb_utils.py:299-385).
    One kernel over all points and boxes (csrc/raster.cu: instance_from_boxes_kernel) instead of a Python loop of
    full-tensor passes per box."""
    import ctypes as C

    from .._lib import check, lib
    n = points_in_world.shape[0]
    instance_id = torch.zeros(n, dtype=torch.int32, device=points_in_world.device)
    tab = pack_instance_boxes(static_object_info, enlarge_lwh_factor)
    if n == 0 or tab.shape[0] == 0:
        return instance_id
    if not points_in_world.is_cuda:
        raise TypeError("get_instance_id_for_fvdb_scene_points needs CUDA tensors (there is no CPU path)")
    pts = points_in_world.to(torch.float32).contiguous()
    sem = semantic.to(torch.int32).contiguous()
    tab = tab.to(pts.device)
    mask = 0
    for c in _CAR_LIKE:
        mask |= 1 << c
    check(lib().ic_instance_from_boxes(C.c_void_p(pts.data_ptr()), n, C.c_void_p(sem.data_ptr()),
                                       C.c_void_p(tab.data_ptr()), tab.shape[0], mask,
                                       C.c_void_p(instance_id.data_ptr()),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)),
          "ic_instance_from_boxes")
    return instance_id


def generate_infinicube_buffer_from_fvdb_grid(
    camera_model: PinholeCamera,
    camera_poses_in_world: torch.Tensor,
    fvdb_scene_grid_or_points: Union[VoxelGrid, torch.Tensor],
    fvdb_scene_semantic: torch.Tensor,
    fvdb_grid_to_world: torch.Tensor,
    static_object_info: Dict,
    dynamic_object_info: Dict = None,
    dynamic_object_points_canonical_data: Dict = None,
    cad_model_for_static_object: bool = False,
    cad_model_for_dynamic_objects: bool = False,
    cad_model_location=None,
    voxel_sizes=(0.2, 0.2, 0.2),
    enlarge_lwh_factor=1.2,
):
    """Returns (depth, semantic, instance) — same order as the reference's return statement
    (fvdb_utils.py:618), each (N, H, W) or (H, W)."""
    from .mesh import generate_object_points_canonical_from_cad_model
    single = camera_poses_in_world.ndim == 2
    poses = camera_poses_in_world.unsqueeze(0) if single else camera_poses_in_world
    n = poses.shape[0]
    static_info = keep_car_only_in_object_info(static_object_info)
    dyn_info = keep_car_only_in_object_info(dynamic_object_info)
    static_canon = (generate_object_points_canonical_from_cad_model(static_info, cad_model_location)
                    if cad_model_for_static_object else {})
    dyn_canon = (generate_object_points_canonical_from_cad_model(dyn_info, cad_model_location)
                 if cad_model_for_dynamic_objects else (dynamic_object_points_canonical_data or {}))
    canon = {**static_canon, **dyn_canon}

    if isinstance(fvdb_scene_grid_or_points, VoxelGrid):
        g0 = fvdb_scene_grid_or_points
        scene_points = g0.grid_to_world(g0.ijk)
    else:
        scene_points = fvdb_scene_grid_or_points
    dev = scene_points.device
    sem = fvdb_scene_semantic.to(dev)
    if cad_model_for_static_object:  # CAD cars replace every car-like voxel of the static scene (fvdb_utils.py:500-508)
        keep = torch.ones_like(sem, dtype=torch.bool)
        for c in _CAR_LIKE:
            keep &= sem != c
        scene_points, sem = scene_points[keep], sem[keep]
    g2w = fvdb_grid_to_world.to(dev, torch.float32)
    scene_points_w = PinholeCamera.transform_points(scene_points.to(torch.float32), g2w)
    inst = get_instance_id_for_fvdb_scene_points(scene_points_w, sem, static_object_info, enlarge_lwh_factor)
    origins = [v / 2 for v in voxel_sizes]

    def frame_objects(i):
        objs = dict(dyn_info.get(f"{i:06d}.dynamic_object_info.json", {}))
        if cad_model_for_static_object:
            objs.update(static_info.get(f"{i:06d}.static_object_info.json", {}))
        return objs

    if not any(len(frame_objects(i)) for i in range(n)):
        # static scene: one grid, every camera in one fused launch
        grid = VoxelGrid(scene_points_w, voxel_sizes, origins, sem, inst)
        depth, s_img, i_img = camera_model.render_voxel_buffers(poses, grid)
    else:
        ds, ss, iss = [], [], []
        for i in range(n):
            pts, sems, insts = [scene_points_w], [sem], [inst]
            for gid, data in frame_objects(i).items():
                o2w = np.array(data["object_to_world"])
                p = PinholeCamera.transform_points(np.asarray(canon[gid + "_xyz"]), o2w)
                p = torch.from_numpy(p).to(scene_points_w)
                pts.append(p)
                sems.append(torch.full((p.shape[0],), int(canon[gid + "_semantic"]), device=dev, dtype=sem.dtype))
                insts.append(torch.full((p.shape[0],), int(data["object_id_int"]), device=dev, dtype=inst.dtype))
            grid = VoxelGrid(torch.cat(pts), voxel_sizes, origins, torch.cat(sems), torch.cat(insts))
            d, s, ii = camera_model.render_voxel_buffers(poses[i:i + 1], grid)
            ds.append(d)
            ss.append(s)
            iss.append(ii)
        depth, s_img, i_img = torch.cat(ds), torch.cat(ss), torch.cat(iss)
    if single:
        return depth[0], s_img[0], i_img[0]
    return depth, s_img, i_img
