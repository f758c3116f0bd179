#!/usr/bin/env python
"""Benchmark of the hot path BASELINE.json names: denoised frames/sec, Wan2.1-1.3B, 93 x 480 x 832.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores

A *step* is one denoising step of the 50-step loop: two full 30-layer DiT forwards (prompt / negative prompt),
the CFG combine and the flow-match Euler update, on the full 37 440-token latent (configs[1]).
`value` = 93 frames / (50 x seconds per step), inputs resident in HBM.  `e2e` = ONE full
`WanVideoGenerator.generate()` call on host uint8 guidance buffers (the call guidance_buffer_generation.py makes):
H2D of both buffer videos, tiled VAE encode x2, the 50-step CFG loop, tiled VAE decode and the D2H of the frames are
all inside the timed region; e2e.value = 93 / call seconds.  With N > 1 the prompt / negative-prompt forwards run on
two rank groups and the token axis is sharded by latent frame inside each group (one (K || V^T) all-gather per
layer); total work is fixed => "strong" scaling.  `cpu_baseline` (N = 1 only) and `--impl reference` time the fp32
oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAMES, HEIGHT, WIDTH = 93, 480, 832
LAT = (16, 24, 60, 104)
NUM_INFERENCE_STEPS = 50
METRIC = "denoised frames/sec Wan2.1-1.3B 93x480p"
METRIC_14B = "denoised frames/sec Wan2.1-14B 93x480p"
UNIT = "frames/s"


def claim_stdout() -> int:
    """The JSON line must be the only thing on stdout: hand fd 1 to stderr for the whole run (NCCL prints its
    version banner to stdout, libraries may print warnings) and keep the real stdout for emit()."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def emit(fd: int, line: dict) -> None:
    os.write(fd, (json.dumps(line) + "\n").encode())


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_burst": d.get("bf16_tflops", 1590.0), "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU arm (oracle port) — the only place besides tests/smoke that executes oracle/
# ----------------------------------------------------------------------------------------------------
def cpu_block_sample(n_threads: int, repeats: int):
    """Times one fp32 DiT block forward of Wan2.1-1.3B at N = 2048 tokens (configs[0] size) on the host and
    extrapolates to the full video by algorithmic FLOPs (same cores, same code)."""
    import torch
    from oracle import wan_dit_oracle as o
    torch.set_num_threads(n_threads)
    cfg = o.WanConfig(num_layers=1)
    sd = o.make_weights(cfg, seed=1234)
    g = torch.Generator().manual_seed(0)
    f, h, w = 8, 16, 16
    x = torch.randn(f * h * w, cfg.dim, generator=g)
    ctx = torch.randn(cfg.text_len, cfg.dim, generator=g)
    t_mod = torch.randn(6, cfg.dim, generator=g) * 0.1
    ang = o.rope_angles(f, h, w, cfg.head_dim)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            o.dit_block(x, ctx, t_mod, sd, 0, cfg, ang)
            times.append(time.perf_counter() - t0)
    n_s = f * h * w
    full = o.WanConfig.wan_1_3b()
    fl_video = o.dit_flops_per_forward(full, LAT[1] * (LAT[2] // 2) * (LAT[3] // 2)) * 2 * NUM_INFERENCE_STEPS
    D, Fd, L = cfg.dim, cfg.ffn_dim, cfg.text_len
    fl_sample = 8 * n_s * D * D + 4 * n_s * n_s * D + 4 * n_s * D * D + 4 * n_s * L * D + 4 * n_s * D * Fd
    return times, fl_sample, fl_video


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    times, fl_sample, fl_video = cpu_block_sample(cores, args.warmup + args.steps)
    timed = times[args.warmup:]
    sec_per_step = sum(timed) / len(timed)
    video_s = sec_per_step * fl_video / fl_sample
    value = FRAMES / video_s
    sample = ("one fp32 DiT block forward (Wan2.1-1.3B dims) at N=2048 tokens per step; whole video extrapolated by "
              "algorithmic FLOPs (x%.0f)" % (fl_video / fl_sample))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Wan2.1-1.3B full DiT, 50 steps x 2 CFG forwards, 93x480x832 (37440 tokens), synthetic "
                               "guidance; CPU arm runs a bounded sample per step", "sample_tokens": 2048},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(args.stdout_fd, line)


def raster_record(dev):
    """Voxel -> guidance-buffer rasteriser at the S = 256 point of BASELINE configs[4]: 93 cameras, 480 x 832, one fused
    launch for depth + semantic + instance, CUDA-event timed; achieved GB/s by the SURVEY §8(d) byte model
    (93 x (20 B x voxels + 12 B x pixels)) against the measured HBM peak, plus the two guidance images."""
    import numpy as np
    import torch
    from infinicube_b200.raster import PinholeCamera, VoxelGrid, synthetic as syn
    from infinicube_b200.raster.buffer_utils import coordinate_buffer
    from infinicube_b200.raster.semantic_utils import semantic_rgb_u8
    S, vs = 256, 0.2
    pts, sem, inst, _ = syn.synthetic_scene(S, voxel_size=vs)
    grid = VoxelGrid(torch.from_numpy(pts).to(dev), [vs] * 3, [vs / 2] * 3, torch.from_numpy(sem).to(dev),
                     torch.from_numpy(inst).to(dev))
    cam = PinholeCamera.from_numpy(syn.DEFAULT_INTRINSICS, device=dev)
    poses = torch.from_numpy(syn.synthetic_poses(S, n=FRAMES, voxel_size=vs)).to(dev)

    def ev(fn, iters=5, warm=2):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    ms = ev(lambda: cam.render_voxel_buffers(poses, grid))
    d, s_img, i_img = cam.render_voxel_buffers(poses, grid)
    rgb_ms = ev(lambda: semantic_rgb_u8(s_img, i_img, rng=np.random.RandomState(0)), iters=3, warm=1)
    coord_ms = ev(lambda: coordinate_buffer(d, cam, poses.cpu(), want_f32=False, want_u8=True), iters=3, warm=1)
    nbytes = syn.raster_algorithmic_bytes(grid.total_voxels, FRAMES, HEIGHT, WIDTH)
    peaks = load_peaks()
    rec = {"workload": "256^3 synthetic voxel world, 93 cameras, 480x832 (configs[4] point)", "n_voxels": int(grid.total_voxels),
           "render_ms_93cams": ms, "mrays_per_s": FRAMES * HEIGHT * WIDTH / ms / 1e3, "algorithmic_bytes": nbytes,
           "achieved_gbs": nbytes / ms / 1e6, "hbm_peak_gbs": peaks["hbm"], "hbm_frac": nbytes / ms / 1e6 / peaks["hbm"],
           "semantic_rgb_ms": rgb_ms, "coordinate_buffer_ms": coord_ms,
           "bound": "instruction issue / ALU (per-pixel two-level DDA): ncu 76 % of issue slots, ALU pipe 73 %, DRAM traffic "
                    "= the 387 MB of images it writes (profiles/r2_raymarch_ncu_summary.json); the byte model charges a "
                    "full voxel-table read per frame that the traversal never needs"}
    del grid, d, s_img, i_img
    torch.cuda.empty_cache()
    return rec


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from infinicube_b200 import _lib
    from infinicube_b200.videogen.pipeline import (DenoiseLoop, FlowMatchScheduler, ParallelLayout, WanDiTEngine, model_timestep,
                                                   WanModelConfig, setup_kv_exchange, synthetic_context,
                                                   synthetic_state_dict)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.require_device()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    big = args.model == "14b"
    cfg = WanModelConfig.wan_14b() if big else WanModelConfig.wan_1_3b()
    C_, F_, H_, W_ = LAT
    layout = ParallelLayout.make(world, rank, None if args.cfg_parallel < 0 else bool(args.cfg_parallel))
    eng = WanDiTEngine(cfg, F_, H_, W_, guide_channels=32, world_size=layout.seq_world, rank=layout.seq_rank,
                       device=dev)
    eng.load_state_dict(synthetic_state_dict(cfg, 32, dev, seed=1234), strict=True)
    kv_exchange = setup_kv_exchange(layout, eng, dev) if world > 1 else "none"
    eng.set_context(0, synthetic_context("The video is about a driving scene captured at daytime. The weather is clear.", cfg, dev))
    eng.set_context(1, synthetic_context("negative prompt", cfg, dev))
    f0, fl = eng.frame0, eng.frames_local
    g = torch.Generator(device="cpu").manual_seed(0)
    noise = torch.randn(LAT, generator=g)
    guide = torch.randn((32, F_, H_, W_), generator=torch.Generator(device="cpu").manual_seed(5))
    eng.set_guidance(guide[:, f0:f0 + fl].to(dev))
    lat = noise[:, f0:f0 + fl].to(dev).contiguous()
    sch = FlowMatchScheduler().set_timesteps(NUM_INFERENCE_STEPS, shift=5.0)
    loop = DenoiseLoop(eng, cfg_scale=5.0, layout=layout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        k = i % NUM_INFERENCE_STEPS
        loop.step(lat, model_timestep(sch.timesteps[k]), sch.delta_sigma(k))

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing (value) -----------------------------------------------------------------
    eng.set_profiling(WanDiTEngine.PROF_FMHA_SELF)   # the roofline kernel only: two event records per launch are not free
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    prof = eng.profile_collect()
    # kernel shares of a step: two more (untimed) steps with every launch kind bracketed
    eng.set_profiling(WanDiTEngine.PROF_ALL)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for i in range(2):
        step(args.warmup + args.steps + i)
    ev3.record()
    torch.cuda.synchronize()
    shares_ms = ev2.elapsed_time(ev3)
    prof_all = eng.profile_collect()
    eng.set_profiling(0)
    barrier()
    launches_per_step = loop.forwards_per_step * eng.launch_count + 1
    flops_per_forward, n_loc, n_tot = eng.flops_per_forward, eng.tokens_local, eng.tokens_total

    # ---- parity record: 2 CFG steps from the seeded noise through THIS layout; at N > 1 rank 0 repeats them on a
    # world_size = 1 engine and reports the difference (north_star: "frames matching on identical noise/seed") -----
    parity = None
    if not args.skip_parity:
        lat2 = noise[:, f0:f0 + fl].to(dev).contiguous()
        for k in range(2):
            loop.step(lat2, model_timestep(sch.timesteps[k]), sch.delta_sigma(k))
        full_lat = lat2
        if world > 1:
            parts = [torch.empty_like(lat2) for _ in range(world)]
            dist.all_gather(parts, lat2)
            full_lat = torch.cat(parts[:layout.seq_world], dim=1)
        if rank == 0:
            d64 = full_lat.double()
            parity = {"steps": 2, "latent_sum": float(d64.sum()), "latent_abs_sum": float(d64.abs().sum()),
                      "finite": bool(torch.isfinite(full_lat).all()), "vs_single_gpu": None}
            if world > 1:
                one = WanDiTEngine(cfg, F_, H_, W_, guide_channels=32, world_size=1, rank=0, device=dev)
                one.load_state_dict(synthetic_state_dict(cfg, 32, dev, seed=1234), strict=True)
                one.set_context(0, synthetic_context("The video is about a driving scene captured at daytime. The weather is clear.", cfg, dev))
                one.set_context(1, synthetic_context("negative prompt", cfg, dev))
                one.set_guidance(guide.to(dev))
                ref = noise.to(dev).contiguous()
                l1 = DenoiseLoop(one, cfg_scale=5.0)
                for k in range(2):
                    l1.step(ref, model_timestep(sch.timesteps[k]), sch.delta_sigma(k))
                nz = noise.to(dev)
                parity["vs_single_gpu"] = {
                    "rel_l2_velocity": float((full_lat - ref).norm() / (ref - nz).norm()),
                    "max_abs_latent_diff": float((full_lat - ref).abs().max()),
                    "bit_identical": bool(torch.equal(full_lat, ref)),
                    "single_gpu_latent_abs_sum": float(ref.double().abs().sum())}
                del one, l1, ref
                torch.cuda.empty_cache()
        barrier()

    # ---- rasteriser sub-record (BASELINE configs[4] point S = 256: the stage that produces the guidance buffers the
    # loop consumes; HBM-bound by the §8d byte model, ALU-bound in practice - see DESIGN.md §4) --------------------
    raster = None
    if rank == 0 and world == 1 and not big and not args.skip_raster:
        raster = raster_record(dev)

    ms2 = torch.tensor([float('nan')], device=dev)
    if not args.skip_e2e:
        # ---- end-to-end through the public API: WanVideoGenerator.generate(host uint8 buffers) -> frames ----------
        # (VAE encode x2 + 50-step CFG loop + VAE decode, H2D of both buffers and D2H of the frames inside the region)
        import numpy as np
        from infinicube_b200.videogen import WanVideoGenerator
        del loop, eng
        torch.cuda.empty_cache()
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):  # the JSON line must be the only stdout output
            gen = WanVideoGenerator(checkpoint_path="synthetic.safetensors", device=f"cuda:{local_rank}", use_wan_1pt3b=not big,
                                    synthetic_weights=True, world_size=world, rank=rank,
                                    cfg_parallel=layout.cfg_parallel)
        rs = np.random.RandomState(0)
        sem_buf = (rs.randint(0, 10, size=(FRAMES, HEIGHT // 8, WIDTH // 8, 1)) * 25).astype(np.uint8)
        sem_buf = np.ascontiguousarray(np.broadcast_to(sem_buf.repeat(8, 1).repeat(8, 2), (FRAMES, HEIGHT, WIDTH, 3)))
        coord_buf = rs.randint(0, 256, size=(FRAMES, HEIGHT, WIDTH, 3), dtype=np.uint8)
        e2e_calls = 1
        with contextlib.redirect_stdout(io.StringIO()):
            if not args.skip_e2e_warmup:
                gen.pipe(prompt="warm", negative_prompt="up", semantic_buffer_video=sem_buf, coordinate_buffer_video=coord_buf,
                         height=HEIGHT, width=WIDTH, num_frames=FRAMES, seed=0, tiled=True, num_inference_steps=1)
            barrier()
            t_wall = time.perf_counter()
            ev0.record()
            video = gen.generate(sem_buf, coord_buf, seed=0, tiled=True)
            ev1.record()
            barrier()
            t_wall = time.perf_counter() - t_wall
        assert len(video) == FRAMES
        ms2 = torch.tensor([max(ev0.elapsed_time(ev1), t_wall * 1e3)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = load_peaks()
        ms_per_step = ms_total / args.steps
        value = FRAMES / (NUM_INFERENCE_STEPS * ms_per_step * 1e-3)
        e2e_call_ms = None if args.skip_e2e else float(ms2)
        e2e_value = None if args.skip_e2e else FRAMES / (e2e_call_ms * 1e-3)
        buf_bytes = FRAMES * HEIGHT * WIDTH * 3
        flops_step = 2.0 * flops_per_forward
        # dominant kernel: self-attention FMHA.  Algorithmic FLOPs per launch = 4 * N_local * N_total * D
        D = cfg.dim
        fmha_ms, fmha_n = prof["fmha_self"]
        fmha_flops = 4.0 * n_loc * n_tot * D
        achieved = fmha_flops / (fmha_ms / max(fmha_n, 1) * 1e-3) / 1e12 if fmha_n else None
        traffic = None
        tfile = ROOT / "profiles" / "fmha_traffic.json"
        if tfile.exists() and world == 1 and not big:  # the capture is of the 1-GPU 1.3B launch; other shapes: null
            traffic = json.loads(tfile.read_text()).get("dram_bytes_per_launch")
        gemm_ms, gemm_n = prof_all["gemm"]
        cross_ms, cross_n = prof_all["fmha_cross"]
        self_ms_all, _ = prof_all["fmha_self"]
        cores = os.cpu_count() or 1
        cpu_baseline = {"value": None, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "not run: the CPU leg is timed on rank 0 at N = 1 only"}
        if world == 1 and not big:
            times, fl_sample, fl_video = cpu_block_sample(cores, 6)
            cpu_sec = sum(times[1:]) / len(times[1:])
            cpu_baseline = {"value": FRAMES / (cpu_sec * fl_video / fl_sample), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "one fp32 oracle DiT block at N=2048 tokens x5, video extrapolated by "
                                      "algorithmic FLOPs (x%.0f)" % (fl_video / fl_sample)}
        line = {
            "metric": METRIC_14B if big else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("Wan2.1-14B full DiT (40 layers, dim 5120), one step = 2 CFG forwards + Euler update, "
                                    "93x480x832 -> 37440 tokens, synthetic weights / context / guidance latents "
                                    "(configs[3]); value = 93 / (50 x s_per_step)") if big else
                                   ("Wan2.1-1.3B full DiT (30 layers), one step = 2 CFG forwards + Euler update, "
                                    "93x480x832 -> 37440 tokens, synthetic weights / context / guidance latents "
                                    "(configs[1]); value = 93 / (50 x s_per_step)"),
                       "num_inference_steps": NUM_INFERENCE_STEPS, "cfg_scale": 5.0, "tokens": n_tot,
                       "parallelism": layout.describe(), "kv_exchange": kv_exchange,
                       "l2_policy": "inputs larger than L2 (weights %s GB, activations > 126 MB per pass)" % ("28" if big else "2.8")},
            "tensor_pipe_fraction": flops_step / (ms_per_step * 1e-3) / world / (peaks["bf16_sustained"] * 1e12),
            "tflops_per_gpu": flops_step / (ms_per_step * 1e-3) / world / 1e12,
            "roofline": {"bound": "tensor", "kernel": "fmha_fwd_kernel (self-attention)", "achieved": achieved,
                         "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_sustained"] if achieved else None, "traffic": traffic,
                         "peak_source": peaks["source"] + " sustained (kernel timed inside a long step)",
                         "avg_launch_ms": fmha_ms / max(fmha_n, 1), "launches_timed": fmha_n,
                         "share_of_step": fmha_ms / ms_total if ms_total else None},
            "kernel_shares": {"fmha_self": self_ms_all / shares_ms, "fmha_cross": cross_ms / shares_ms,
                              "gemm": gemm_ms / shares_ms, "gemm_launches_per_step": gemm_n // 2,
                              "what": "CUDA-event shares of 2 extra untimed steps with every launch kind bracketed "
                                      "(rank 0)"},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * buf_bytes // NUM_INFERENCE_STEPS,
                    "d2h_bytes_per_step": buf_bytes // NUM_INFERENCE_STEPS, "call_ms": e2e_call_ms,
                    "what": "one WanVideoGenerator.generate() call on host uint8 buffers (2 x %d B in, %d B of frames "
                            "out): tiled VAE encode x2 + 50 CFG steps + tiled VAE decode; bytes are per call / 50 steps"
                            % (buf_bytes, buf_bytes)},
            "parity": parity,
            "raster": raster,
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        emit(args.stdout_fd, line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="1.3b", choices=["1.3b", "14b"],
                    help="1.3b = BASELINE configs[1] (the headline metric); 14b = configs[3] (the reference's default model)")
    ap.add_argument("--cfg-parallel", type=int, default=-1, choices=[-1, 0, 1],
                    help="N>1: run the prompt / negative-prompt forwards on two rank groups (-1: library default)")
    ap.add_argument("--skip-e2e-warmup", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="developer runs only: e2e is reported as null")
    ap.add_argument("--skip-parity", action="store_true", help="developer runs only: no 2-step parity record")
    ap.add_argument("--skip-raster", action="store_true", help="developer runs only: no rasteriser sub-record")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    args.stdout_fd = claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
