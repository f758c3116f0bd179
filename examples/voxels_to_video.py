"""The whole hot path on one B200, GPU-resident: synthetic voxel world -> depth / semantic / instance buffers ->
uint8 guidance images -> Wan2.1 (synthetic weights) -> frames.  Mirrors what
infinicube/inference/guidance_buffer_generation.py:591-791 does with the reference stack.

    python examples/voxels_to_video.py --size 128 --frames 93 [--out video.mp4]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from infinicube_b200.raster import PinholeCamera, generate_infinicube_buffer_from_fvdb_grid, synthetic as syn  # noqa: E402
from infinicube_b200.raster.buffer_utils import coordinate_buffer  # noqa: E402
from infinicube_b200.raster.semantic_utils import semantic_rgb_u8  # noqa: E402
from infinicube_b200.videogen import WanVideoGenerator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--frames", type=int, default=93)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    pts, sem, inst, _ = syn.synthetic_scene(a.size)
    cam = PinholeCamera.from_numpy(syn.DEFAULT_INTRINSICS, device=dev)
    poses = torch.from_numpy(syn.synthetic_poses(a.size, n=a.frames)).to(dev)
    t0 = time.perf_counter()
    depth, s_img, i_img = generate_infinicube_buffer_from_fvdb_grid(
        cam, poses, torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev).long(), torch.eye(4),
        static_object_info={}, dynamic_object_info={}, dynamic_object_points_canonical_data={})
    rng = np.random.RandomState(0)
    sem_rgb = semantic_rgb_u8(s_img, i_img, rng=rng)                                  # uint8 [N,H,W,3] on the GPU
    torch.manual_seed(0)
    _, coord_u8 = coordinate_buffer(depth, cam, poses.cpu(), want_f32=False, want_u8=True)
    torch.cuda.synchronize()
    print(f"guidance buffers for {a.frames} frames: {time.perf_counter() - t0:.3f} s")
    gen = WanVideoGenerator("synthetic.safetensors", device="cuda:0", use_wan_1pt3b=True, synthetic_weights=True)
    t0 = time.perf_counter()
    frames = gen.generate_device(sem_rgb, coord_u8, seed=0, tiled=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"video: {tuple(frames.shape)} uint8 on {frames.device} in {dt:.1f} s ({a.frames / dt:.2f} frames/s)")
    if a.out:
        from PIL import Image
        from infinicube_b200.videogen.inference import save_video
        save_video([Image.fromarray(f) for f in frames.cpu().numpy()], a.out, fps=10)


if __name__ == "__main__":
    main()
