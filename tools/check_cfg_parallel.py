"""N-GPU check (torchrun): both multi-GPU layouts reproduce the SINGLE-GPU denoising loop.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_cfg_parallel.py [--full] [--out profiles/r2_shard_parity_N.json]

Every rank runs `steps` CFG denoising steps twice - ParallelLayout(cfg_parallel=False) (plain temporal-token shard xN)
and cfg_parallel=True (prompt | negative groups x temporal shard N/2) - through the product's exchange path (NCCL
all-gather, or the peer-memory push with ICB_KV_P2P=1).  Rank 0 ALSO runs the same loop on a world_size = 1 engine and
compares the gathered latents of each layout with it:
  * default: Wan2.1-1.3B dims, 2 layers, 8 x 32 x 48 latent (384 tokens per frame: every shard is a multiple of the
    128-key tile);
  * --full:  4 layers on the 24 x 60 x 104 bench latent (37 440 / N tokens per rank: ragged key tiles, so agreement is
    to bf16-rounding level - P is rounded relative to a running maximum that depends on the tiling; rel-L2 of the
    accumulated velocity is reported and bounded by 1.5e-2, the same order as either run's distance to the fp32 oracle).
Exit status 1 if a bound is violated."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import torch
    import torch.distributed as dist
    from infinicube_b200.videogen.pipeline import (DenoiseLoop, FlowMatchScheduler, ParallelLayout, WanDiTEngine,
                                                   WanModelConfig, setup_kv_exchange, synthetic_context,
                                                   synthetic_state_dict)
    full = "--full" in sys.argv
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    cfg = WanModelConfig(num_layers=4 if full else 2)
    F_, H_, W_ = (24, 60, 104) if full else (8, 32, 48)
    steps = 2
    noise = torch.randn((16, F_, H_, W_), generator=torch.Generator().manual_seed(0))
    guide = torch.randn((32, F_, H_, W_), generator=torch.Generator().manual_seed(5))
    sd = synthetic_state_dict(cfg, 32, dev, seed=1234)
    sch = FlowMatchScheduler().set_timesteps(50, shift=5.0)

    def run(layout, with_comm=True):
        eng = WanDiTEngine(cfg, F_, H_, W_, 32, layout.seq_world, layout.seq_rank, dev)
        eng.load_state_dict(sd)
        kind = setup_kv_exchange(layout, eng, dev) if with_comm else "none"
        eng.set_context(0, synthetic_context("a street", cfg, dev))
        eng.set_context(1, synthetic_context("negative", cfg, dev))
        f0, fl = eng.frame0, eng.frames_local
        eng.set_guidance(guide[:, f0:f0 + fl].to(dev))
        lat = noise[:, f0:f0 + fl].to(dev).contiguous()
        DenoiseLoop(eng, 5.0, layout).run(lat, sch, steps=steps)
        torch.cuda.synchronize()
        return eng, lat, kind

    results, kinds = {}, {}
    for mode in (False, True):
        if mode and world % 2:
            continue
        layout = ParallelLayout.make(world, rank, mode)
        eng, lat, kind = run(layout)
        parts = [torch.empty_like(lat) for _ in range(world)]
        dist.all_gather(parts, lat)
        results[mode] = torch.cat(parts[:layout.seq_world], dim=1)
        kinds[mode] = kind
        if layout.cfg_parallel:  # both CFG groups must hold identical latents
            results["groups_equal"] = bool(torch.equal(results[mode], torch.cat(parts[layout.seq_world:], dim=1)))
        dist.barrier()
        torch.cuda.synchronize()
        del eng
        torch.cuda.synchronize()
        dist.barrier()
    ok = True
    if rank == 0:
        _, ref, _ = run(ParallelLayout(1, 0, False), with_comm=False)
        nz = noise.to(dev)
        ds = sch.delta_sigma(0) + sch.delta_sigma(1)
        out = {"world": world, "full": full, "latent": [F_, H_, W_], "layers": cfg.num_layers, "steps": steps,
               "tokens_per_rank_temporal": F_ * (H_ // 2) * (W_ // 2) // world, "finite": True, "layouts": {}}
        for mode, name in ((False, f"temporal-token shard x{world}"), (True, f"cfg x2 x temporal shard x{world // 2}")):
            if mode not in results:
                continue
            got = results[mode]
            rel = float(((got - ref) / ds).norm() / ((ref - nz) / ds).norm())
            rec = {"kv_exchange": kinds[mode], "bit_identical_to_single_gpu": bool(torch.equal(got, ref)),
                   "rel_l2_velocity_vs_single_gpu": rel, "max_abs_latent_diff": float((got - ref).abs().max())}
            out["finite"] = out["finite"] and bool(torch.isfinite(got).all())
            out["layouts"][name] = rec
            # Bit-identity with the single GPU is only guaranteed when both runs launch the same kernels on the same
            # tiling (tests/test_gpu_dit.py::test_token_shard_equals_single_gpu pins that case); here the shard's
            # GEMMs may pick another tile shape (M differs) and ragged shards tile the keys differently, so the bound
            # is the bf16-rounding level either run has against the fp32 oracle.  The flag is reported, not required.
            ok = ok and rel < 1.5e-2
        out["cfg_groups_equal"] = results.get("groups_equal")
        out["pass"] = bool(ok and out["finite"] and results.get("groups_equal", True))
        ok = out["pass"]
        print("SHARD_PARITY " + json.dumps(out))
        for p in filter(None, [out_path, f"gpurun_out/shard_parity_{world}gpu{'_full' if full else ''}.json"]):
            Path(p).parent.mkdir(parents=True, exist_ok=True)
            Path(p).write_text(json.dumps(out, indent=1))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
