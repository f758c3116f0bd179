"""Generates tests/golden/r4_instance_ids.npz by running the REFERENCE'S OWN
`get_instance_id_for_fvdb_scene_points` (infinicube/utils/fvdb_utils.py:299-385) on seeded inputs.

    python oracle/gen_golden_r4.py        # build container only: the GPU box has no /root/reference

The function is plain torch, but its module imports fvdb / torch_scatter / trimesh at the top; empty stub modules
stand in for them (nothing of theirs is called on this path).  Inputs cover: the four car-like classes and
non-car classes, rotated + translated boxes, the x1.2 enlargement (points between the tight and the enlarged box),
overlapping boxes (the LATER box in dict order wins), points far from every box, and an empty box dict.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from gen_golden_from_reference import REF, _install_stubs  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "r4_instance_ids.npz"


def make_case(seed: int):
    rs = np.random.RandomState(seed)
    boxes = {}
    n_box = 6
    for b in range(n_box):
        yaw = rs.uniform(-np.pi, np.pi)
        c, s = np.cos(yaw), np.sin(yaw)
        T = np.eye(4)
        T[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        T[:3, 3] = [rs.uniform(-20, 20), rs.uniform(-20, 20), rs.uniform(0, 2)]
        if b == n_box - 1:  # overlaps box 0: the later box must win inside the intersection
            T[:3, 3] = boxes["gid0"]["object_to_world"][:3, 3] + np.array([1.0, 0.3, 0.0])
        boxes[f"gid{b}"] = {"object_to_world": T, "object_lwh": np.array([rs.uniform(3.5, 6), rs.uniform(1.6, 2.4),
                                                                          rs.uniform(1.4, 2.2)]),
                            "object_is_moving": False, "object_type": "car", "object_id_int": int(100 + 7 * b)}
    pts, sem = [], []
    for b, d in enumerate(boxes.values()):
        # points in the box frame up to 1.5 x the half extents: inside, in the x1.2 shell, and outside
        loc = (rs.uniform(-1.5, 1.5, size=(400, 3)) * d["object_lwh"] / 2)
        w = loc @ d["object_to_world"][:3, :3].T + d["object_to_world"][:3, 3]
        pts.append(w)
        sem.append(rs.choice([1, 2, 3, 4, 14, 18, 0], size=400, p=[0.4, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1]))
    pts.append(rs.uniform(-40, 40, size=(500, 3)))
    sem.append(rs.choice([1, 14, 18], size=500))
    pts = np.round(np.concatenate(pts) / 0.2) * 0.2 + 0.1      # voxel centres of a 0.2 m grid
    return pts.astype(np.float32), np.concatenate(sem).astype(np.int64), boxes


def main():
    sys.path.insert(0, str(REF))
    _install_stubs()
    for name, attrs in [("fvdb", dict(GridBatch=object, JaggedTensor=object)), ("torch_scatter", {}), ("trimesh", {}),
                        ("infinicube.data_process", {}),
                        ("infinicube.data_process.waymo_utils", dict(keep_car_only_in_object_info=lambda x: x))]:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    from infinicube.utils.fvdb_utils import get_instance_id_for_fvdb_scene_points as ref_fn
    from infinicube.utils.semantic_utils import WAYMO_CATEGORY_NAMES
    out = {"car_like": np.array([WAYMO_CATEGORY_NAMES.index(n) for n in ("CAR", "TRUCK", "BUS", "OTHER_VEHICLE")])}
    for ci, (seed, factor) in enumerate([(1, 1.0), (2, 1.2), (3, 1.2)]):
        pts, sem, boxes = make_case(seed)
        info = {"000000.static_object_info.json": {k: {**v, "object_to_world": v["object_to_world"].tolist(),
                                                      "object_lwh": v["object_lwh"].tolist()} for k, v in boxes.items()}}
        ids = ref_fn(torch.from_numpy(pts), torch.from_numpy(sem), info, enlarge_lwh_factor=factor)
        out[f"c{ci}_points"] = pts
        out[f"c{ci}_sem"] = sem
        out[f"c{ci}_factor"] = np.array(factor)
        out[f"c{ci}_o2w"] = np.stack([b["object_to_world"] for b in boxes.values()])
        out[f"c{ci}_lwh"] = np.stack([b["object_lwh"] for b in boxes.values()])
        out[f"c{ci}_id"] = np.array([b["object_id_int"] for b in boxes.values()])
        out[f"c{ci}_out"] = ids.numpy()
        print(f"case {ci}: {len(pts)} points, {(ids > 0).sum().item()} labelled, ids {sorted(set(ids.tolist()))}")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
