"""Single-process multi-rank entry (videogen/multiproc.py) on CPU with gloo: the calling process is rank 0, worker
ranks are spawned, every call runs on all ranks, big arrays travel by broadcast, failures on any rank raise in the
caller, the pool shuts its workers down.  (the stage-2 caller is one process with one cached generator.)"""
import numpy as np
import pytest


@pytest.mark.parametrize("world", [2, 3])
def test_pool_runs_collectives_on_all_ranks(world):
    from infinicube_b200.videogen.multiproc import EchoRank, RankPool
    pool = RankPool(world, EchoRank, dict(scale=0.5), backend="gloo", start_timeout_s=120)
    try:
        small = np.arange(6, dtype=np.float64).reshape(2, 3)                    # pipe path
        big = np.random.RandomState(0).rand(64, 300)                            # >= 64 KiB: broadcast path
        tri = world * (world + 1) / 2
        out = pool.call("weighted_sum", small, tag="a")
        assert np.allclose(out, small * tri * 0.5)
        out = pool.call("weighted_sum", arr=big)
        assert np.allclose(out, big * tri * 0.5)
        u8 = np.random.RandomState(1).randint(0, 255, size=(5, 64, 96, 3), dtype=np.uint8)   # guidance-buffer-like
        out = pool.call("weighted_sum", u8)
        assert np.allclose(out, u8.astype(np.float64) * tri * 0.5)
        assert pool.obj.calls == 3
        # a failure on a worker rank, and on rank 0, raises in the caller and leaves the pool usable
        with pytest.raises(RuntimeError, match="requested failure on rank 1"):
            pool.call("fail_on", 1)
        with pytest.raises(ValueError, match="rank 0"):
            pool.call("fail_on", 0)
        assert np.allclose(pool.call("weighted_sum", small), small * tri * 0.5)
        procs = list(pool._procs)
        assert len(procs) == world - 1 and all(p.is_alive() for p in procs)
    finally:
        pool.close()
    assert all(not p.is_alive() for p in procs)
    import torch.distributed as dist
    assert not dist.is_initialized()


def test_generator_single_process_entry_uses_the_pool(monkeypatch):
    """WanVideoGenerator(world_size=2) in a plain process builds a RankPool around make_rank_generator and routes
    generate() through it (the pool itself is replaced by a recorder here: no GPU in this test)."""
    import infinicube_b200.videogen.inference as inf
    import infinicube_b200.videogen.multiproc as mp_
    made = {}

    class FakePool:
        def __init__(self, world_size, factory, factory_kwargs, **kw):
            made.update(world=world_size, factory=factory, kwargs=factory_kwargs)
            self.obj = type("O", (), {"pipe": "rank0-pipe"})()
            self.calls = []

        def call(self, method, *a, **k):
            self.calls.append((method, a, k))
            return ["frame"] * a[0].shape[0]

        def close(self):
            made["closed"] = True

    monkeypatch.setattr(mp_, "RankPool", FakePool)
    monkeypatch.delenv("RANK", raising=False)
    gen = inf.WanVideoGenerator("ckpt.safetensors", device="cuda:0", use_wan_1pt3b=True, world_size=2,
                                synthetic_weights=True)
    assert made["world"] == 2 and made["factory"] is inf.make_rank_generator and gen.pipe == "rank0-pipe"
    assert made["kwargs"]["use_wan_1pt3b"] is True and made["kwargs"]["checkpoint_path"] == "ckpt.safetensors"
    buf = np.zeros((5, 16, 16, 3), np.uint8)
    out = gen.generate(buf, buf, seed=3, tiled=False)
    assert out == ["frame"] * 5
    method, a, k = gen._pool.calls[0]
    assert method == "generate" and a[0] is buf and k["seed"] == 3 and k["tiled"] is False
    with pytest.raises(ValueError):      # validation happens in the caller, before anything is shipped
        gen.generate(buf, buf[:4])
    assert len(gen._pool.calls) == 1
    gen.close()
    assert made.get("closed")
    # under a per-rank launcher (torchrun sets RANK) no pool is created
    monkeypatch.setenv("RANK", "0")
    assert mp_.launched_per_rank()
