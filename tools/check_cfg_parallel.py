"""2+ GPU check (torchrun): the CFG-parallel layout reproduces the plain temporal-shard loop.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_cfg_parallel.py

Every rank runs `steps` denoising steps of a small Wan-shaped model twice - once with ParallelLayout(cfg_parallel=False)
and once with cfg_parallel=True - and rank 0 prints the difference of the gathered latents.  With two ranks the
CFG-parallel forwards are unsharded, so the comparison also pins the temporal all-gather path against the
single-GPU arithmetic."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import torch
    import torch.distributed as dist
    from infinicube_b200.videogen.pipeline import (DenoiseLoop, FlowMatchScheduler, ParallelLayout, WanDiTEngine,
                                                   WanModelConfig, exchange_nccl_unique_id, synthetic_context,
                                                   synthetic_state_dict)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    cfg = WanModelConfig(num_layers=2)
    F_, H_, W_ = 8, 32, 48
    steps = 3
    noise = torch.randn((16, F_, H_, W_), generator=torch.Generator().manual_seed(0))
    guide = torch.randn((32, F_, H_, W_), generator=torch.Generator().manual_seed(5))
    sd = synthetic_state_dict(cfg, 32, dev, seed=1234)
    results = {}
    for mode in (False, True):
        layout = ParallelLayout.make(world, rank, mode)
        eng = WanDiTEngine(cfg, F_, H_, W_, 32, layout.seq_world, layout.seq_rank, dev)
        eng.load_state_dict(sd)
        uid = exchange_nccl_unique_id(layout, dev)
        if uid is not None:
            eng.init_comm(uid)
        eng.set_context(0, synthetic_context("a street", cfg, dev))
        eng.set_context(1, synthetic_context("negative", cfg, dev))
        f0, fl = eng.frame0, eng.frames_local
        eng.set_guidance(guide[:, f0:f0 + fl].to(dev))
        lat = noise[:, f0:f0 + fl].to(dev).contiguous()
        sch = FlowMatchScheduler().set_timesteps(50, shift=5.0)
        DenoiseLoop(eng, 5.0, layout).run(lat, sch, steps=steps)
        parts = [torch.empty_like(lat) for _ in range(world)]
        pad = lat
        if layout.seq_world != world:   # shards are larger: gather per group through equal-sized pieces
            parts = [torch.empty_like(lat) for _ in range(world)]
        dist.all_gather(parts, pad)
        results[mode] = torch.cat(parts[:layout.seq_world], dim=1)
        if layout.cfg_parallel:  # both CFG groups must hold identical latents
            other = torch.cat(parts[layout.seq_world:], dim=1)
            results["groups_equal"] = bool(torch.equal(results[mode], other))
        del eng
        torch.cuda.synchronize()
    if rank == 0:
        a, b = results[False].float(), results[True].float()
        out = {"world": world, "rel_l2": float((a - b).norm() / b.norm()), "max_abs": float((a - b).abs().max()),
               "groups_equal": results.get("groups_equal"), "finite": bool(torch.isfinite(b).all())}
        print("CFG_PARALLEL_CHECK " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
