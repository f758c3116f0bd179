"""Rasteriser throughput sweep (BASELINE configs[4]): S^3 synthetic worlds, 93 cameras, 480x832.
Writes gpurun_out/raster_sweep.json with ms, algorithmic GB/s (SURVEY §8d formula) and fraction of measured HBM."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from infinicube_b200.raster import PinholeCamera, VoxelGrid, synthetic as syn  # noqa: E402
from infinicube_b200.raster.buffer_utils import coordinate_buffer  # noqa: E402
from infinicube_b200.raster.semantic_utils import semantic_rgb_u8  # noqa: E402


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cpu_baseline(sizes=(64, 256), n_cam=None):
    """CPU legs, timed on THIS host next to the GPU numbers:
    (i) "port": the oracle's two-level DDA (oracle/raster_oracle.c) on ALL host cores - one camera per task on a
        thread pool (ctypes releases the GIL) - over a bounded sample of the 93 cameras, extrapolated linearly;
    (ii) "reference": the reference-authored CPU path BASELINE.md names, CameraBase.get_zdepth_map_from_points_torch
        (camera/base.py:386-447), as restated and pinned in oracle/points_splat_oracle.py, fed the voxel centres
        (depth only - it cannot produce the semantic / instance images), torch on all cores, same camera sample.
    tools/ may execute the oracle only as a measured baseline, never inside the product."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    from oracle import points_splat_oracle as ps
    from oracle import raster_oracle as ro
    cores = os.cpu_count() or 1
    n_cam = n_cam or max(2, min(93, cores))
    torch.set_num_threads(cores)
    out = []
    for S in sizes:
        vs = 0.2
        pts, sem, inst, _ = syn.synthetic_scene(S, voxel_size=vs)
        t0 = time.perf_counter()
        og = ro.OracleGrid(pts, [vs] * 3, [vs / 2] * 3, sem, inst)
        build_s = time.perf_counter() - t0
        poses = syn.synthetic_poses(S, n=93, voxel_size=vs)
        pick = np.linspace(0, 92, n_cam).round().astype(int)
        kinv = ro.inv_intrinsics_matrix(syn.DEFAULT_INTRINSICS)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(lambda c: og.render(kinv, poses[c:c + 1], 832, 480), pick))
        render_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        d = ps.zdepth_map_from_points(syn.DEFAULT_INTRINSICS, torch.from_numpy(poses[pick[:max(2, n_cam // 4)]]),
                                      torch.from_numpy(pts))
        splat_s = (time.perf_counter() - t0) / max(2, n_cam // 4) * n_cam
        out.append({"S": S, "n_vox": int(len(pts)), "host_cpus": cores, "sample": f"{n_cam} of 93 cameras, 480x832",
                    "port": {"kind": "port", "what": "oracle two-level DDA, depth + semantic + instance", "cores": cores,
                             "grid_build_s": build_s, "render_s_sample": render_s,
                             "render_s_93cams_extrapolated": render_s * 93 / n_cam,
                             "mrays_per_s": n_cam * 480 * 832 / render_s / 1e6},
                    "reference": {"kind": "reference-authored (restated)", "cores": cores,
                                  "what": "get_zdepth_map_from_points_torch over the voxel centres, depth only",
                                  "splat_s_sample": splat_s, "splat_s_93cams_extrapolated": splat_s * 93 / n_cam,
                                  "hit_fraction": float((d > 0).float().mean())}})
        print(json.dumps(out[-1]), flush=True)
    return out


def main():
    if "--cpu-only" in sys.argv:
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "raster_cpu_baseline.json").write_text(json.dumps({"cpu_baseline": cpu_baseline()}, indent=1))
        return
    dev = torch.device("cuda:0")
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    out = []
    for S in (64, 128, 256, 512):
        vs = 0.2
        pts, sem, inst, _ = syn.synthetic_scene(S, voxel_size=vs)
        p_d, s_d, i_d = torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev), torch.from_numpy(inst).to(dev)
        t0 = time.perf_counter()
        grid = VoxelGrid(p_d, [vs] * 3, [vs / 2] * 3, s_d, i_d)
        torch.cuda.synchronize()
        build_ms_first = (time.perf_counter() - t0) * 1e3
        build_ms = ev_time(lambda: VoxelGrid(p_d, [vs] * 3, [vs / 2] * 3, s_d, i_d), iters=3, warm=1)
        cam = PinholeCamera.from_numpy(syn.DEFAULT_INTRINSICS, device=dev)
        poses = torch.from_numpy(syn.synthetic_poses(S, n=93, voxel_size=vs)).to(dev)
        ms = ev_time(lambda: cam.render_voxel_buffers(poses, grid))
        d, s, i = cam.render_voxel_buffers(poses, grid)
        rgb_ms = ev_time(lambda: semantic_rgb_u8(s, i, {k: np.array([0.5, 0.2, 0.7]) for k in range(1, S // 16 + 1)}))
        torch.manual_seed(0)
        coord_ms = ev_time(lambda: coordinate_buffer(d, cam, poses.cpu(), want_f32=False, want_u8=True), iters=2, warm=1)
        nbytes = syn.raster_algorithmic_bytes(grid.total_voxels, 93, 480, 832)
        out.append({"S": S, "n_vox": grid.total_voxels, "n_bricks": grid.num_bricks, "build_ms": build_ms,
                    "build_ms_first_call": build_ms_first, "render_ms_93cams": ms,
                    "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6,
                    "hbm_frac": nbytes / ms / 1e6 / peaks["hbm_gbs"], "hit_fraction": float((s > 0).float().mean()),
                    "mrays_per_s": 93 * 480 * 832 / ms / 1e3, "semantic_rgb_ms": rgb_ms, "coord_buffer_ms": coord_ms})
        print(json.dumps(out[-1]), flush=True)
        del grid
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    cpu = cpu_baseline() if "--no-cpu" not in sys.argv else None
    if cpu:   # GPU / CPU ratios on the same box, same workload
        by_s = {r["S"]: r for r in out}
        for c in cpu:
            g = by_s.get(c["S"])
            if g:
                c["gpu_render_ms_93cams"] = g["render_ms_93cams"]
                c["speedup_vs_port_all_cores"] = c["port"]["render_s_93cams_extrapolated"] * 1e3 / g["render_ms_93cams"]
                c["speedup_vs_reference_splat"] = c["reference"]["splat_s_93cams_extrapolated"] * 1e3 / g["render_ms_93cams"]
    (ROOT / "gpurun_out" / "raster_sweep.json").write_text(json.dumps({
        "peak_hbm_gbs": peaks["hbm_gbs"], "gpu": torch.cuda.get_device_name(0), "sweep": out, "cpu_baseline": cpu}, indent=1))


if __name__ == "__main__":
    main()
