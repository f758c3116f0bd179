"""Generates tests/golden/reference_vectors.npz by running the REFERENCE'S OWN code (imported from
/root/reference with stub modules for its missing heavy dependencies) on small seeded inputs.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden_from_reference.py

What can be imported this way: infinicube.camera.{pinhole,base},
infinicube.utils.{depth_utils,buffer_utils,semantic_utils,instance_utils}, and the input validation of
infinicube.videogen.inference (with a fake `diffsynth`).  The DiT arithmetic (diffsynth) and the
ray/voxel traversal (fvdb) cannot: those parts of the oracle stay "parity unpinned".
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "reference_vectors.npz"

# ColorBrewer tables matplotlib serves for these names (r/255 floats), needed by the pycg stub
_BREWER = {
    "Set2": ["66c2a5", "fc8d62", "8da0cb", "e78ac3", "a6d854", "ffd92f", "e5c494", "b3b3b3"],
    "Set3": ["8dd3c7", "ffffb3", "bebada", "fb8072", "80b1d3", "fdb462", "b3de69", "fccde5", "d9d9d9", "bc80bd",
             "ccebc5", "ffed6f"],
    "Set1": ["e41a1c", "377eb8", "4daf4a", "984ea3", "ff7f00", "ffff33", "a65628", "f781bf", "999999"],
    "Paired": ["a6cee3", "1f78b4", "b2df8a", "33a02c", "fb9a99", "e31a1c", "fdbf6f", "ff7f00", "cab2d6", "6a3d9a",
               "ffff99", "b15928"],
}


def _install_stubs():
    root = types.ModuleType("infinicube")
    root.__path__ = [str(REF / "infinicube")]
    sys.modules["infinicube"] = root

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("shapely")
    mod("shapely.geometry", Polygon=object)
    mod("decord", VideoReader=object, cpu=lambda *a, **k: None)
    mod("webdataset", WebDataset=object, non_empty=lambda *a, **k: None)
    mpl = mod("matplotlib", colormaps={})
    mod("matplotlib.pyplot")
    mod("matplotlib.cm")
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    mpl.cm = sys.modules["matplotlib.cm"]
    pycg = mod("pycg")
    color = mod("pycg.color", get_cmap_array=lambda name: np.array(
        [[int(c[i:i + 2], 16) / 255.0 for i in (0, 2, 4)] for c in _BREWER[name]]))
    pycg.color = color
    for name in ("imageio", "imageio.v3", "mediapy", "termcolor", "loguru", "tqdm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                mod(name, colored=lambda s, *a, **k: s, logger=None, tqdm=lambda x, *a, **k: x)
    # fake diffsynth for the videogen wrapper's validation paths
    ds = mod("diffsynth", load_state_dict=lambda p: {}, save_video=lambda *a, **k: None)
    mod("diffsynth.pipelines")

    class _Pipe:
        buffer_embedder = None
        dit = types.SimpleNamespace(load_state_dict=lambda *a, **k: None)

        @staticmethod
        def from_pretrained(**kw):
            return _Pipe()

        def initialize_buffer_embedder(self, **kw):
            pass

        def enable_vram_management(self):
            pass

        def __call__(self, **kw):
            return ["frame"] * kw["num_frames"]

    mod("diffsynth.pipelines.wan_video_new", ModelConfig=lambda **kw: kw, WanVideoPipeline=_Pipe)
    return ds


def main():
    sys.path.insert(0, str(REF))
    _install_stubs()
    from infinicube.camera.pinhole import PinholeCamera
    from infinicube.utils.buffer_utils import generate_coordinate_buffer_from_memory_global_norm
    from infinicube.utils.depth_utils import unproject_depth_torch
    from infinicube.utils.semantic_utils import WAYMO_MAPPING, WAYMO_PALETTE, semantic_to_color

    out = {}
    cpu = torch.device("cpu")
    # ---- camera: intrinsics, rescale (pinhole.py:62-75), rays (pinhole.py:110-138), posed rays (base.py:207-226)
    intr0 = np.array([2059.6, 2059.6, 941.0, 640.3, 1920, 1280], dtype=np.float64)
    cam = PinholeCamera.from_numpy(intr0, device=cpu)
    cam.rescale(30 / 1280, 52 / 1920)
    out["cam_intr_in"] = intr0
    out["cam_intr_rescaled"] = cam.intrinsics
    out["cam_K"] = cam.get_intrinsics_matrix().numpy()
    out["cam_Kinv"] = cam.intrinsics_matrix_inv_torch.numpy()
    out["cam_rays"] = cam.get_rays().numpy()
    g = torch.Generator().manual_seed(11)
    poses = torch.eye(4).repeat(3, 1, 1)
    for i in range(3):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        poses[i, :3, :3] = q
        poses[i, :3, 3] = torch.randn(3, generator=g) * 3
    ro, rd = cam.get_rays_posed(poses)
    out["poses"] = poses.numpy()
    out["rays_o"] = ro.numpy()
    out["rays_d"] = rd.numpy()
    pts = torch.randn(17, 3, generator=g)
    out["tp_points"] = pts.numpy()
    out["tp_out"] = PinholeCamera.transform_points(pts, poses[1]).numpy()

    # ---- palette (semantic_utils.py:22-101) and uint8 truncation (guidance_buffer_generation.py:693-695)
    labels = np.arange(23)
    out["waymo_mapping"] = WAYMO_MAPPING
    out["waymo_palette"] = WAYMO_PALETTE
    colors = semantic_to_color(labels)
    out["label_colors_f32"] = colors
    out["label_colors_u8"] = (colors * 255).astype(np.uint8)

    # ---- coordinate buffer (buffer_utils.py:180-265, depth_utils.py:402-466)
    cam2 = PinholeCamera(40.0, 38.0, 25.5, 14.25, 52, 30, device=cpu)
    depth = torch.rand(3, 30, 52, generator=g) * 30 + 1
    depth[torch.rand(3, 30, 52, generator=g) < 0.2] = 0.0
    out["cb_depth"] = depth.numpy()
    out["cb_intr"] = cam2.intrinsics
    K = cam2.get_intrinsics_matrix().unsqueeze(0).repeat(3, 1, 1)
    c2c0 = torch.einsum("ij,bjk->bik", torch.inverse(poses[0]), poses)
    out["cb_xyz"] = unproject_depth_torch(depth.unsqueeze(1), c2c0, K).numpy()
    torch.manual_seed(1234)
    cb = generate_coordinate_buffer_from_memory_global_norm(depth, cam2, poses)
    out["cb_seed"] = np.array(1234)
    out["cb_out_f32"] = cb.numpy()
    out["cb_out_u8"] = (cb.numpy() * 255).astype(np.uint8)  # guidance_buffer_generation.py:710

    # ---- videogen wrapper validation (videogen/inference.py:130-162,194-198)
    from infinicube.videogen.inference import WanVideoGenerator
    gen = WanVideoGenerator("dummy.safetensors", device="cpu", use_wan_1pt3b=True)
    ok = np.zeros((5, 16, 16, 3), dtype=np.uint8)
    errs = []
    for name, a, b in [
        ("float_dtype", ok.astype(np.float32), ok.astype(np.float32)),
        ("shape_mismatch", ok, ok[:4]),
        ("bad_channels", np.zeros((5, 16, 16, 4), np.uint8), np.zeros((5, 16, 16, 4), np.uint8)),
        ("ok", ok, ok),
    ]:
        try:
            r = gen.generate(a, b)
            errs.append(f"{name}:ok:{len(r)}")
        except Exception as e:  # noqa: BLE001
            errs.append(f"{name}:{type(e).__name__}")
    out["videogen_validation"] = np.array(errs)

    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
