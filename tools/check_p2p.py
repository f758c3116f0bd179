"""2+ GPU check (torchrun) of the peer-memory K / V^T exchange (ICB_KV_P2P path) against the NCCL
all-gather path: identical arithmetic up to the order in which a rank walks the key segments (bit-identical on
rank 0, rounding-level elsewhere); then both are timed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        tools/check_p2p.py [--full]

Default: a small Wan-shaped model (2 layers, 8 x 32 x 48 latent), 3 CFG steps.  --full: Wan2.1-1.3B dims, 4 layers, the
24 x 60 x 104 bench latent (segments of 37 440 / world tokens), 2 steps, with per-step timings of both paths.
NOT yet run on hardware (round-1 GPU budget was spent): run this first in round 2, under `timeout`."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import torch
    import torch.distributed as dist
    from infinicube_b200.videogen.pipeline import (DenoiseLoop, FlowMatchScheduler, ParallelLayout, WanDiTEngine, model_timestep,
                                                   WanModelConfig, exchange_nccl_unique_id, exchange_p2p_handles,
                                                   synthetic_context, synthetic_state_dict)
    full = "--full" in sys.argv
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    cfg = WanModelConfig(num_layers=4 if full else 2)
    F_, H_, W_ = (24, 60, 104) if full else (8, 32, 48)
    steps = 2 if full else 3
    noise = torch.randn((16, F_, H_, W_), generator=torch.Generator().manual_seed(0))
    guide = torch.randn((32, F_, H_, W_), generator=torch.Generator().manual_seed(5))
    sd = synthetic_state_dict(cfg, 32, dev, seed=1234)
    layout = ParallelLayout.make(world, rank, False)   # plain temporal shard: every rank is in one exchange group
    results, ms = {}, {}
    for mode in ("nccl", "p2p"):
        eng = WanDiTEngine(cfg, F_, H_, W_, 32, layout.seq_world, layout.seq_rank, dev)
        eng.load_state_dict(sd)
        if mode == "p2p":
            exchange_p2p_handles(layout, eng)
            assert eng.p2p_enabled
        else:
            eng.init_comm(exchange_nccl_unique_id(layout, dev))
        eng.set_context(0, synthetic_context("a street", cfg, dev))
        eng.set_context(1, synthetic_context("negative", cfg, dev))
        f0, fl = eng.frame0, eng.frames_local
        eng.set_guidance(guide[:, f0:f0 + fl].to(dev))
        lat = noise[:, f0:f0 + fl].to(dev).contiguous()
        sch = FlowMatchScheduler().set_timesteps(50, shift=5.0)
        loop = DenoiseLoop(eng, 5.0, layout)
        loop.run(lat, sch, steps=steps)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        probe = lat.clone()
        for i in range(3):
            loop.step(probe, model_timestep(sch.timesteps[i]), sch.delta_sigma(i))
        torch.cuda.synchronize()
        dist.barrier()
        ms[mode] = (time.perf_counter() - t0) / 3 * 1e3
        parts = [torch.empty_like(lat) for _ in range(world)]
        dist.all_gather(parts, lat)
        results[mode] = torch.cat(parts, dim=1)
        dist.barrier()          # nobody frees a buffer a peer may still be pushing into
        torch.cuda.synchronize()
        del loop, eng
        torch.cuda.synchronize()
        dist.barrier()
    if rank == 0:
        a, b = results["nccl"], results["p2p"]
        # rank 0 walks the segments in the same order on both paths (bit-identical there); ranks r > 0 start at their
        # own segment on the push path, so their online-softmax accumulation order differs at rounding level
        n0 = a.shape[1] // world
        out = {"world": world, "full": full, "bit_identical": bool(torch.equal(a, b)),
               "rank0_bit_identical": bool(torch.equal(a[:, :n0], b[:, :n0])),
               "rel_l2": float((a - b).norm() / a.norm()), "max_abs": float((a - b).abs().max()),
               "finite": bool(torch.isfinite(b).all()), "ms_per_step_nccl": ms["nccl"], "ms_per_step_p2p": ms["p2p"]}
        print("P2P_CHECK " + json.dumps(out))
        Path("gpurun_out").mkdir(exist_ok=True)
        Path("gpurun_out/check_p2p.json").write_text(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
