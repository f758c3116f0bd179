"""CAD-model voxelisation for inserted cars (mirrors generate_object_points_canonical_from_cad_model,
infinicube/utils/fvdb_utils.py:219-296: trimesh.load + per-object scale to lwh + fvdb.gridbatch_from_mesh at
0.1 m, origin 0.05).  The triangle / voxel overlap runs in csrc/raster.cu (ic_mesh_voxelize_mask)."""
from __future__ import annotations

import ctypes as C
import os
import struct
from pathlib import Path
from typing import Dict, Tuple

import numpy as np
import torch

from .._lib import check, lib, require_device
from .semantic_utils import WAYMO_CATEGORY_NAMES

_PLY_TYPES = {"char": "b", "uchar": "B", "short": "h", "ushort": "H", "int": "i", "uint": "I", "float": "f",
              "double": "d", "int8": "b", "uint8": "B", "int16": "h", "uint16": "H", "int32": "i", "uint32": "I",
              "float32": "f", "float64": "d"}


def load_ply(path) -> Tuple[np.ndarray, np.ndarray]:
    """Minimal PLY reader (binary little-endian or ascii): vertices float64 [V,3], triangles int64 [F,3]
    (polygons are fan-triangulated, like trimesh does on load)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode("ascii", "replace").splitlines()
    fmt = [l.split()[1] for l in header if l.startswith("format")][0]
    elements, cur = [], None
    for l in header:
        t = l.split()
        if not t:
            continue
        if t[0] == "element":
            cur = {"name": t[1], "count": int(t[2]), "props": []}
            elements.append(cur)
        elif t[0] == "property" and cur is not None:
            if t[1] == "list":
                cur["props"].append(("list", t[2], t[3], t[4]))
            else:
                cur["props"].append((t[1], t[2]))
    verts, faces = None, []
    if fmt == "ascii":
        lines = data[end:].decode("ascii").split("\n")
        li = 0
        for el in elements:
            rows = lines[li:li + el["count"]]
            li += el["count"]
            if el["name"] == "vertex":
                names = [p[1] for p in el["props"]]
                arr = np.array([[float(v) for v in r.split()] for r in rows], dtype=np.float64)
                verts = arr[:, [names.index("x"), names.index("y"), names.index("z")]]
            elif el["name"] == "face":
                for r in rows:
                    idx = [int(v) for v in r.split()]
                    faces.append(idx[1:1 + idx[0]])
    else:
        if fmt != "binary_little_endian":
            raise ValueError(f"unsupported PLY format {fmt}")
        off = end
        for el in elements:
            if el["name"] == "vertex":
                dt = np.dtype([(p[1], "<" + _PLY_TYPES[p[0]]) for p in el["props"]])
                arr = np.frombuffer(data, dtype=dt, count=el["count"], offset=off)
                off += dt.itemsize * el["count"]
                verts = np.stack([arr["x"], arr["y"], arr["z"]], axis=1).astype(np.float64)
            elif el["name"] == "face":
                (_, ct, it, _), = [p for p in el["props"] if p[0] == "list"]
                cs, isz = struct.calcsize(_PLY_TYPES[ct]), struct.calcsize(_PLY_TYPES[it])
                for _ in range(el["count"]):
                    n = struct.unpack_from("<" + _PLY_TYPES[ct], data, off)[0]
                    off += cs
                    faces.append(struct.unpack_from("<" + str(n) + _PLY_TYPES[it], data, off))
                    off += isz * n
            else:
                raise ValueError(f"unsupported PLY element {el['name']}")
    tris = []
    for fc in faces:
        for k in range(1, len(fc) - 1):
            tris.append((fc[0], fc[k], fc[k + 1]))
    return verts, np.asarray(tris, dtype=np.int64).reshape(-1, 3)


def voxelize_mesh(vertices: np.ndarray, faces: np.ndarray, voxel_size: float = 0.1, origin: float = 0.05,
                  device="cuda") -> np.ndarray:
    """Voxel coordinates ijk int32 [N,3] (sorted by k, j, i) of all voxels overlapped by the mesh surface."""
    require_device()
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    f = np.ascontiguousarray(faces, dtype=np.int32)
    lo = np.floor((v.min(0) - origin) / voxel_size - 0.5).astype(np.int64) - 1
    hi = np.ceil((v.max(0) - origin) / voxel_size + 0.5).astype(np.int64) + 1
    dims = (hi - lo + 1).astype(np.int32)
    nbits = int(dims[0]) * int(dims[1]) * int(dims[2])
    vd = torch.from_numpy(v).to(device)
    fd = torch.from_numpy(f).to(device)
    mask = torch.empty((nbits + 31) // 32, dtype=torch.int32, device=device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib().ic_mesh_voxelize_mask(C.c_void_p(vd.data_ptr()), v.shape[0], C.c_void_p(fd.data_ptr()), f.shape[0],
                                      float(voxel_size), float(origin), (C.c_int * 3)(*[int(x) for x in lo]),
                                      (C.c_int * 3)(*[int(x) for x in dims]), C.c_void_p(mask.data_ptr()), st),
          "ic_mesh_voxelize_mask")
    bits = np.unpackbits(mask.cpu().numpy().view(np.uint8), bitorder="little")[:nbits]
    lin = np.nonzero(bits)[0]
    i = lin % dims[0]
    j = (lin // dims[0]) % dims[1]
    k = lin // (int(dims[0]) * int(dims[1]))
    return (np.stack([i, j, k], axis=1) + lo[None, :]).astype(np.int32)


def default_cad_model_location() -> Path:
    """The reference ships the model as infinicube/assets/car.ply; it is not copied into this repo."""
    env = os.environ.get("INFINICUBE_CAD_MODEL")
    if env:
        return Path(env)
    try:
        import infinicube  # type: ignore
        return Path(infinicube.__file__).parent / "assets" / "car.ply"
    except Exception:  # noqa: BLE001
        return Path("infinicube/assets/car.ply")


def generate_object_points_canonical_from_cad_model(car_object_info: Dict, cad_model_location=None) -> Dict:
    """{gid_xyz: (M,3) float32 voxel centres of the CAD car scaled to the object's lwh, gid_semantic: CAR}."""
    lwh, seen = {}, []
    for key, frame in (car_object_info or {}).items():
        if key.endswith(".json"):
            for gid, data in frame.items():
                if gid not in lwh:
                    lwh[gid] = data["object_lwh"]
                    seen.append(gid)
    if not seen:
        return {}
    path = Path(cad_model_location) if cad_model_location is not None else default_cad_model_location()
    if not path.exists():
        raise FileNotFoundError(f"CAD model {path} not found (set INFINICUBE_CAD_MODEL or pass cad_model_location)")
    verts, faces = load_ply(path)
    mesh_lwh = verts.max(0) - verts.min(0)
    out = {}
    car = WAYMO_CATEGORY_NAMES.index("CAR")
    for gid in seen:
        scale = np.asarray(lwh[gid], dtype=np.float64) / mesh_lwh
        ijk = voxelize_mesh(verts * scale[None, :], faces, 0.1, 0.05)
        out[gid + "_xyz"] = (ijk.astype(np.float32) * np.float32(0.1) + np.float32(0.05)).astype(np.float32)
        out[gid + "_semantic"] = car
    return out
