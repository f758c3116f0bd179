"""Single-process entry to the multi-GPU path (SURVEY §5.8, §8b "Threading / lifetime").

The reference's stage-2 caller is ONE Python process that caches ONE generator on "cuda:0"
(`infinicube/inference/guidance_buffer_generation.py:755-768`) and calls `generate()` on it.  The sharded engine is
one process per GPU.  `RankPool` bridges the two without `torchrun`: the calling process becomes rank 0, spawns
ranks 1..W-1 (`multiprocessing` "spawn" context, one per GPU), joins them in a `torch.distributed` group over
127.0.0.1, and from then on every method call made through the pool runs on ALL ranks - the arguments are shipped to
the workers (small ones through pipes, arrays / tensors with `dist.broadcast`), each rank executes the same method
of its own object (whose internals do the token-shard collectives), and rank 0's return value is the caller's.

Workers are daemons, exit when the pool is closed or the parent dies, and report exceptions back: a failure on any
rank raises in the caller (the stage-2 script catches `Exception` and moves on, `guidance_buffer_generation.py:786-791`).
"""
from __future__ import annotations

import os
import socket
import traceback
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

_BIG = 1 << 16  # arrays at least this large travel by dist.broadcast instead of through the command pipe


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _split_args(args: tuple, kwargs: dict):
    """-> (args', kwargs', specs): big arrays / tensors replaced by placeholders, described in `specs`."""
    specs: List[Tuple[str, tuple, str]] = []
    big: List[torch.Tensor] = []

    def enc(v):
        if isinstance(v, np.ndarray) and v.nbytes >= _BIG:
            specs.append(("np", tuple(v.shape), str(v.dtype)))
            big.append(torch.from_numpy(np.ascontiguousarray(v)))
            return ("__big__", len(specs) - 1)
        if isinstance(v, torch.Tensor) and v.numel() * v.element_size() >= _BIG:
            specs.append(("torch", tuple(v.shape), str(v.dtype).replace("torch.", "")))
            big.append(v.contiguous())
            return ("__big__", len(specs) - 1)
        return v

    return tuple(enc(a) for a in args), {k: enc(v) for k, v in kwargs.items()}, specs, big


def _join_args(args: tuple, kwargs: dict, big: List[Any]):
    def dec(v):
        if isinstance(v, tuple) and len(v) == 2 and v[0] == "__big__":
            return big[v[1]]
        return v

    return tuple(dec(a) for a in args), {k: dec(v) for k, v in kwargs.items()}


def _exchange_big(specs, big: Optional[List[torch.Tensor]], device: torch.device, src: int = 0) -> List[Any]:
    """Rank `src` passes its tensors, the others pass None; everyone returns the materialised list."""
    import torch.distributed as dist
    out = []
    for i, (kind, shape, dtype) in enumerate(specs):
        tdt = getattr(torch, dtype) if kind == "torch" else torch.from_numpy(np.empty(0, dtype=dtype)).dtype
        t = big[i].to(device) if big is not None else torch.empty(shape, dtype=tdt, device=device)
        dist.broadcast(t, src=src)
        if kind == "np":
            out.append(t.cpu().numpy() if big is None else big[i].numpy())
        else:
            out.append(t if big is None else big[i])
    return out


def _worker(rank: int, world: int, port: int, backend: str, factory: Callable, factory_kwargs: Dict, conn) -> None:
    import torch.distributed as dist
    try:
        device = torch.device("cpu")
        if backend == "nccl":
            torch.cuda.set_device(rank)
            device = torch.device("cuda", rank)
        dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                **({"device_id": device} if backend == "nccl" else {}))
        obj = factory(rank=rank, world_size=world, **factory_kwargs)
        conn.send(("ready", None))
    except BaseException as e:  # noqa: BLE001 - reported to the parent, which raises
        conn.send(("error", f"rank {rank} failed to start: {e!r}\n{traceback.format_exc()}"))
        return
    while True:
        try:
            msg = conn.recv()
        except (EOFError, OSError):
            break
        if msg[0] == "close":
            break
        _, method, args, kwargs, specs = msg
        try:
            big = _exchange_big(specs, None, device)
            a, k = _join_args(args, kwargs, big)
            getattr(obj, method)(*a, **k)
            conn.send(("ok", None))
        except BaseException as e:  # noqa: BLE001
            conn.send(("error", f"rank {rank}: {e!r}\n{traceback.format_exc()}"))
    try:
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


class RankPool:
    """Rank 0 lives in the calling process; ranks 1..world_size-1 are spawned.  `factory(rank=, world_size=, **kw)`
    builds the per-rank object on every rank (it must be a picklable top-level callable)."""

    def __init__(self, world_size: int, factory: Callable, factory_kwargs: Optional[Dict] = None, backend: str = "nccl",
                 start_timeout_s: float = 600.0):
        import multiprocessing as mp

        import torch.distributed as dist
        if world_size < 2:
            raise ValueError("RankPool is for world_size >= 2")
        if dist.is_available() and dist.is_initialized():
            raise RuntimeError("torch.distributed is already initialised in this process: launch-per-rank mode "
                               "(torchrun) and the single-process RankPool are mutually exclusive")
        if backend == "nccl" and torch.cuda.device_count() < world_size:
            raise RuntimeError(f"world_size {world_size} needs {world_size} visible GPUs, found {torch.cuda.device_count()}")
        self.world_size, self.backend = world_size, backend
        self.device = torch.device("cuda", 0) if backend == "nccl" else torch.device("cpu")
        port = _free_port()
        ctx = mp.get_context("spawn")
        self._conns, self._procs = [], []
        kw = dict(factory_kwargs or {})
        for r in range(1, world_size):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(r, world_size, port, backend, factory, kw, child), daemon=True)
            p.start()
            child.close()
            self._conns.append(parent)
            self._procs.append(p)
        try:
            if backend == "nccl":
                torch.cuda.set_device(0)
            dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=world_size,
                                    **({"device_id": self.device} if backend == "nccl" else {}))
            self.obj = factory(rank=0, world_size=world_size, **kw)
            for r, c in enumerate(self._conns, start=1):
                if not c.poll(start_timeout_s):
                    raise RuntimeError(f"rank {r} did not come up within {start_timeout_s:.0f} s")
                tag, info = c.recv()
                if tag != "ready":
                    raise RuntimeError(info)
        except BaseException:
            self.close()
            raise

    def call(self, method: str, *args, **kwargs):
        """Runs `obj.method(*args, **kwargs)` on every rank; returns rank 0's result.  Raises if any rank failed."""
        a, k, specs, big = _split_args(args, kwargs)
        for c in self._conns:
            c.send(("call", method, a, k, specs))
        err = None
        result = None
        try:
            _exchange_big(specs, big, self.device)
            result = getattr(self.obj, method)(*args, **kwargs)
        except BaseException as e:  # noqa: BLE001 - still drain the workers' replies so the pool stays usable
            err = e
        failures = []
        for r, c in enumerate(self._conns, start=1):
            tag, info = c.recv()
            if tag != "ok":
                failures.append(info)
        if err is not None:
            raise err
        if failures:
            raise RuntimeError("worker rank(s) failed:\n" + "\n".join(failures))
        return result

    def close(self):
        import torch.distributed as dist
        for c in getattr(self, "_conns", []):
            try:
                c.send(("close",))
            except Exception:  # noqa: BLE001
                pass
        for p in getattr(self, "_procs", []):
            p.join(timeout=20)
            if p.is_alive():
                p.terminate()
        self._conns, self._procs = [], []
        if dist.is_available() and dist.is_initialized():
            try:
                dist.destroy_process_group()
            except Exception:  # noqa: BLE001
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def launched_per_rank() -> bool:
    """True when this process is one rank of an externally launched job (torchrun / already-initialised group)."""
    import torch.distributed as dist
    return (dist.is_available() and dist.is_initialized()) or "RANK" in os.environ


class EchoRank:
    """Minimal per-rank object used by the CPU (gloo) tests of the pool: every method is a collective."""

    def __init__(self, rank: int, world_size: int, scale: float = 1.0):
        self.rank, self.world_size, self.scale = rank, world_size, scale
        self.calls = 0

    def weighted_sum(self, arr: np.ndarray, tag: str = "") -> np.ndarray:
        """sum over ranks of arr * (rank + 1) * scale - proves that every rank received the same array."""
        import torch.distributed as dist
        t = torch.from_numpy(np.asarray(arr, dtype=np.float64) * (self.rank + 1) * self.scale)
        dist.all_reduce(t)
        self.calls += 1
        return t.numpy()

    def fail_on(self, rank: int) -> str:
        import torch.distributed as dist
        dist.barrier()
        if self.rank == rank:
            raise ValueError(f"requested failure on rank {rank}")
        return "ok"
