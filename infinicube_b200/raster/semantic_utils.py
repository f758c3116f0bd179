"""Semantic palette + instance overlay (mirrors infinicube/utils/semantic_utils.py:22-131 and
infinicube/utils/instance_utils.py:21-56).  The per-pixel work runs in csrc/raster.cu.

Provenance: `WAYMO_CATEGORY_NAMES`, `WAYMO_VISUALIZATION_TYPES_BLUE_SKY` and `build_waymo_mapping_and_palette()` are the
reference's constant tables and the short routine that indexes them (`utils/semantic_utils.py:22-83`), reproduced
because the palette is data the output must match bit for bit (pinned by `tests/golden/reference_vectors.npz`)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch

from .._lib import ICError, check, lib, require_device

WAYMO_CATEGORY_NAMES = [
    "UNDEFINED", "CAR", "TRUCK", "BUS", "OTHER_VEHICLE", "MOTORCYCLIST", "BICYCLIST", "PEDESTRIAN", "SIGN",
    "TRAFFIC_LIGHT", "POLE", "CONSTRUCTION_CONE", "BICYCLE", "MOTORCYCLE", "BUILDING", "VEGETATION", "TREE_TRUNK",
    "CURB", "ROAD", "LANE_MARKER", "OTHER_GROUND", "WALKABLE", "SIDEWALK",
]
WAYMO_VISUALIZATION_TYPES_BLUE_SKY = {
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

# ColorBrewer qualitative sets as served by matplotlib / pycg.color.get_cmap_array (hex, r/255 floats)
_BREWER = {
    "Set2": ["66c2a5", "fc8d62", "8da0cb", "e78ac3", "a6d854", "ffd92f", "e5c494", "b3b3b3"],
    "Set3": ["8dd3c7", "ffffb3", "bebada", "fb8072", "80b1d3", "fdb462", "b3de69", "fccde5", "d9d9d9", "bc80bd",
             "ccebc5", "ffed6f"],
    "Set1": ["e41a1c", "377eb8", "4daf4a", "984ea3", "ff7f00", "ffff33", "a65628", "f781bf", "999999"],
    "Paired": ["a6cee3", "1f78b4", "b2df8a", "33a02c", "fb9a99", "e31a1c", "fdbf6f", "ff7f00", "cab2d6", "6a3d9a",
               "ffff99", "b15928"],
    # sequential 9-class anchors of the two colormaps used for instances
    "PuRd": ["f7f4f9", "e7e1ef", "d4b9da", "c994c7", "df65b0", "e7298a", "ce1256", "980043", "67001f"],
    "YlOrBr": ["ffffe5", "fff7bc", "fee391", "fec44f", "fe9929", "ec7014", "cc4c02", "993404", "662506"],
}


def get_cmap_array(name: str) -> np.ndarray:
    return np.array([[int(c[i:i + 2], 16) / 255.0 for i in (0, 2, 4)] for c in _BREWER[name]], dtype=np.float64)


def build_waymo_mapping_and_palette() -> Tuple[np.ndarray, np.ndarray]:
    waymo_mapping = np.zeros(23, dtype=np.int32)
    for palette_idx, names in WAYMO_VISUALIZATION_TYPES_BLUE_SKY.items():
        for n in names:
            waymo_mapping[WAYMO_CATEGORY_NAMES.index(n)] = palette_idx
    waymo_palette = np.zeros((10, 3), dtype=np.float32)
    waymo_palette[:8] = get_cmap_array("Set2")
    waymo_palette[3] = get_cmap_array("Set3")[9]
    waymo_palette[4] = get_cmap_array("Set1")[2]
    waymo_palette[8] = get_cmap_array("Paired")[1]
    waymo_palette[9] = get_cmap_array("Set3")[10]
    return waymo_mapping, waymo_palette


WAYMO_MAPPING, WAYMO_PALETTE = build_waymo_mapping_and_palette()
# label -> float colour and label -> uint8 colour ((c*255).astype(uint8): truncation,
# guidance_buffer_generation.py:693-695)
LABEL_COLORS_F32 = np.ascontiguousarray(WAYMO_PALETTE[WAYMO_MAPPING])
LABEL_COLORS_U8 = np.ascontiguousarray((LABEL_COLORS_F32 * 255).astype(np.uint8))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_i32(x, device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=device, dtype=torch.int32).contiguous()


def semantic_to_color(semantics) -> np.ndarray:
    """label array -> float32 colours [..., 3] in [0,1] (numpy, like the reference)."""
    require_device()
    dev = semantics.device if isinstance(semantics, torch.Tensor) and semantics.is_cuda else torch.device("cuda")
    sem = _dev_i32(semantics, dev)
    lut = torch.from_numpy(LABEL_COLORS_F32).to(dev)
    out = torch.empty((*sem.shape, 3), dtype=torch.float32, device=dev)
    check(lib().ic_lut_gather_f32(C.c_void_p(sem.data_ptr()), sem.numel(), C.c_void_p(lut.data_ptr()), lut.shape[0],
                                  C.c_void_p(out.data_ptr()), _stream()), "ic_lut_gather_f32")
    return out.cpu().numpy()


def _sequential_cmap(name: str, x: np.ndarray) -> np.ndarray:
    """matplotlib's LinearSegmentedColormap.from_list lookup (N = 256) for the 9 ColorBrewer anchors."""
    anchors = get_cmap_array(name)
    xa = np.linspace(0.0, 1.0, anchors.shape[0])
    xind = np.linspace(0.0, 1.0, 256)
    lut = np.stack([np.interp(xind, xa, anchors[:, c]) for c in range(3)], axis=1)
    idx = np.clip((np.asarray(x) * 256).astype(np.int64), 0, 255)
    return lut[idx]


def create_instance_mapping(unique_instance_ids: np.ndarray, rng=None, color_map_for_vechile="PuRd",
                            color_map_for_pedestrian="YlOrBr") -> Dict[int, np.ndarray]:
    """instance id -> RGB in [0,1] (instance_utils.py:21-56).  The reference draws with the unseeded global
    np.random; pass `rng` (np.random.Generator / RandomState) for reproducible buffers."""
    draw = np.random.rand if rng is None else (rng.random if hasattr(rng, "random") else rng.rand)
    ids = np.asarray(unique_instance_ids)
    veh, ped = ids[ids < 2 ** 15], ids[ids >= 2 ** 15]
    cv = _sequential_cmap(color_map_for_vechile, draw(len(veh)))
    cp = _sequential_cmap(color_map_for_pedestrian, draw(len(ped)))
    out = {int(i): c for i, c in zip(veh, cv)}
    out.update({int(i): c for i, c in zip(ped, cp)})
    return out


def semantic_rgb_u8(semantic: torch.Tensor, instance: Optional[torch.Tensor] = None,
                    instance_colors: Optional[Dict[int, np.ndarray]] = None, rng=None,
                    base_rgb: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused device path: labels (+ instance ids) -> uint8 RGB guidance image, stays in HBM."""
    require_device()
    ref = semantic if semantic is not None else base_rgb
    if not ref.is_cuda:
        raise ICError("semantic_rgb_u8 needs CUDA tensors")
    dev = ref.device
    shape = tuple(semantic.shape) if semantic is not None else tuple(base_rgb.shape[:-1])
    sem = None if semantic is None else semantic.to(torch.int32).contiguous()
    inst = None if instance is None else instance.to(device=dev, dtype=torch.int32).contiguous()
    ids_t = cols_t = None
    n_ids = 0
    if inst is not None:
        if instance_colors is None:
            uniq = torch.unique(inst).cpu().numpy()
            instance_colors = create_instance_mapping(uniq[uniq != 0], rng)
        ids = np.array(sorted(instance_colors.keys()), dtype=np.int32)
        n_ids = len(ids)
        if n_ids:
            cols = np.stack([(np.asarray(instance_colors[int(i)], dtype=np.float64)[:3] * 255).astype(np.uint8)
                             for i in ids])
            ids_t = torch.from_numpy(ids).to(dev)
            cols_t = torch.from_numpy(np.ascontiguousarray(cols)).to(dev)
    pal = torch.from_numpy(LABEL_COLORS_U8).to(dev)
    out = torch.empty((*shape, 3), dtype=torch.uint8, device=dev)
    n = int(np.prod(shape))
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
    check(lib().ic_semantic_rgb(p(sem), p(base_rgb), p(inst), n, p(pal), pal.shape[0], p(ids_t), p(cols_t), n_ids,
                                p(out), _stream()), "ic_semantic_rgb")
    return out


def generate_rgb_semantic_buffer(semantics_rgb: np.ndarray, instance_buffer: Union[np.ndarray, torch.Tensor],
                                 instance_colors: Optional[Dict[int, np.ndarray]] = None, rng=None) -> np.ndarray:
    """Overlay coloured instances on the uint8 semantic RGB buffer (semantic_utils.py:104-131)."""
    require_device()
    dev = torch.device("cuda")
    base = torch.from_numpy(np.ascontiguousarray(semantics_rgb, dtype=np.uint8)).to(dev)
    inst = _dev_i32(instance_buffer, dev)
    return semantic_rgb_u8(None, inst, instance_colors, rng, base_rgb=base).cpu().numpy()
