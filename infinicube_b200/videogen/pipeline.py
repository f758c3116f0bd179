"""Host side of the Wan2.1 video pipeline: the surface `infinicube/videogen/inference.py` consumes from
diffsynth (`WanVideoPipeline.from_pretrained / initialize_buffer_embedder / enable_vram_management /
__call__`, `ModelConfig`) re-implemented over the C-ABI engine.  Python here only
orchestrates; every FLOP of the denoising loop runs in libinfinicube_b200.so.

Defaults the reference does not override: cfg_scale 5.0, 50 steps, sigma_shift 5.0,
noise drawn on a CPU generator seeded with `seed`.
"""
from __future__ import annotations

import ctypes as C
import glob
import hashlib
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from .. import _lib
from .._lib import DitConfig, ICError, check, lib, require_device


# --------------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------------
@dataclass
class WanModelConfig:
    dim: int = 1536
    ffn_dim: int = 8960
    num_heads: int = 12
    num_layers: int = 30
    in_dim: int = 16
    out_dim: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    text_len: int = 512
    eps: float = 1e-6

    @staticmethod
    def wan_1_3b() -> "WanModelConfig":
        return WanModelConfig()

    @staticmethod
    def wan_14b() -> "WanModelConfig":
        return WanModelConfig(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40)


@dataclass
class ModelConfig:
    """Same fields the reference passes (videogen/inference.py:66-80)."""
    model_id: Optional[str] = None
    origin_file_pattern: Optional[str] = None
    skip_download: bool = False
    offload_device: Optional[str] = None
    path: Optional[str] = None
    local_model_path: str = "./models"

    def resolve(self) -> List[str]:
        if self.path:
            return [self.path]
        return sorted(glob.glob(os.path.join(self.local_model_path, self.model_id or "", self.origin_file_pattern or "")))


# --------------------------------------------------------------------------------------------------
# scheduler (diffsynth FlowMatchScheduler)
# --------------------------------------------------------------------------------------------------
class FlowMatchScheduler:
    def __init__(self, shift: float = 5.0, num_train_timesteps: int = 1000):
        self.shift = shift
        self.num_train_timesteps = num_train_timesteps
        self.set_timesteps(50, shift=shift)

    def set_timesteps(self, num_inference_steps: int = 50, denoising_strength: float = 1.0, shift: Optional[float] = None):
        if shift is not None:
            self.shift = shift
        s = torch.linspace(denoising_strength, 0.0, num_inference_steps + 1, dtype=torch.float32)[:-1]
        self.sigmas = self.shift * s / (1 + (self.shift - 1) * s)
        self.timesteps = self.sigmas * self.num_train_timesteps
        return self

    def delta_sigma(self, i: int) -> float:
        nxt = self.sigmas[i + 1] if i + 1 < len(self.sigmas) else torch.tensor(0.0)
        return float(nxt - self.sigmas[i])


# --------------------------------------------------------------------------------------------------
# synthetic weights (benchmarks / smoke tests: there are no Wan checkpoints offline)
# --------------------------------------------------------------------------------------------------
def synthetic_state_dict(cfg: WanModelConfig, guide_channels: int, device, seed: int = 1234,
                         zero_guidance: bool = False) -> Dict[str, torch.Tensor]:
    """Random-init weights of the Wan2.1 architecture under the official key names."""
    g = torch.Generator(device=device).manual_seed(seed)
    D, Fd = cfg.dim, cfg.ffn_dim
    sd: Dict[str, torch.Tensor] = {}

    def rnd(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, device=device) * std).to(torch.bfloat16)

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = rnd(out_f, in_f)
        sd[name + ".bias"] = rnd(out_f)

    sd["patch_embedding.weight"] = rnd(D, cfg.in_dim, 1, 2, 2)
    sd["patch_embedding.bias"] = rnd(D)
    lin("text_embedding.0", D, cfg.text_dim)
    lin("text_embedding.2", D, D)
    lin("time_embedding.0", D, cfg.freq_dim)
    lin("time_embedding.2", D, D)
    lin("time_projection.1", 6 * D, D)
    def norm_w():  # non-unit so that the norm-weight paths are exercised wherever these weights are used
        return (1.0 + 0.25 * torch.randn(D, generator=g, device=device)).to(torch.bfloat16)

    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        for a in ("self_attn", "cross_attn"):
            for n in ("q", "k", "v", "o"):
                lin(p + f"{a}.{n}", D, D)
            sd[p + f"{a}.norm_q.weight"] = norm_w()
            sd[p + f"{a}.norm_k.weight"] = norm_w()
        sd[p + "norm3.weight"] = norm_w()
        sd[p + "norm3.bias"] = rnd(D, std=0.1)
        lin(p + "ffn.0", Fd, D)
        lin(p + "ffn.2", D, Fd)
        sd[p + "modulation"] = rnd(1, 6, D, std=1.0 / math.sqrt(D))
    lin("head.head", 4 * cfg.out_dim, D)
    sd["head.modulation"] = rnd(1, 2, D, std=1.0 / math.sqrt(D))
    if guide_channels > 0:
        if zero_guidance:
            sd["buffer_embedder.weight"] = torch.zeros(D, guide_channels, 1, 2, 2, device=device, dtype=torch.bfloat16)
            sd["buffer_embedder.bias"] = torch.zeros(D, device=device, dtype=torch.bfloat16)
        else:
            sd["buffer_embedder.weight"] = rnd(D, guide_channels, 1, 2, 2)
            sd["buffer_embedder.bias"] = rnd(D)
    return sd


def synthetic_context(prompt: str, cfg: WanModelConfig, device) -> torch.Tensor:
    """Deterministic stand-in for the umT5-XXL prompt embedding (the text encoder runs once
    per call and is out of scope): [text_len, text_dim] bf16, rows past the 'prompt length' zeroed like the
    reference zeroes its padding rows."""
    h = int.from_bytes(hashlib.sha256(prompt.encode("utf-8")).digest()[:8], "little") % (2 ** 31)
    g = torch.Generator(device="cpu").manual_seed(h)
    ctx = torch.randn(cfg.text_len, cfg.text_dim, generator=g)
    n_tok = max(1, min(cfg.text_len, len(prompt.split()) + 1))
    ctx[n_tok:] = 0
    return ctx.to(device=device, dtype=torch.bfloat16)


# --------------------------------------------------------------------------------------------------
# engine wrapper
# --------------------------------------------------------------------------------------------------
def shard_frames(lat_f: int, world_size: int, rank: int):
    """Equal temporal-token shards: rank r owns latent frames [r*F/G, (r+1)*F/G)."""
    if lat_f % world_size:
        raise ValueError(f"{lat_f} latent frames do not shard evenly over {world_size} ranks")
    per = lat_f // world_size
    return rank * per, per


@dataclass(frozen=True)
class ParallelLayout:
    """How G ranks split one denoising step.  `seq_world` ranks share a DiT forward along the
    temporal-token axis (one all-gather of K / V^T per attention).  With `cfg_parallel` the prompt and the
    negative-prompt forward of a step run at the same time on two such groups - ranks [0, G/2) and [G/2, G) -
    and rank r swaps its head output (tokens_local x 64 fp32) with rank r +- G/2 before the CFG combine, so every
    forward is sharded half as finely (fuller attention waves, half the all-gathers per rank)."""
    world_size: int = 1
    rank: int = 0
    cfg_parallel: bool = False

    def __post_init__(self):
        if not 0 <= self.rank < self.world_size:
            raise ValueError(f"rank {self.rank} outside world of {self.world_size}")
        if self.cfg_parallel and self.world_size % 2:
            raise ValueError("cfg_parallel needs an even number of ranks")

    @staticmethod
    def make(world_size: int = 1, rank: int = 0, cfg_parallel: Optional[bool] = None) -> "ParallelLayout":
        """cfg_parallel=None: on for an even world unless ICB_CFG_PARALLEL=0."""
        if cfg_parallel is None:
            cfg_parallel = os.environ.get("ICB_CFG_PARALLEL", "1") != "0"
        return ParallelLayout(world_size, rank, bool(cfg_parallel) and world_size >= 2 and world_size % 2 == 0)

    @property
    def seq_world(self) -> int:
        return self.world_size // 2 if self.cfg_parallel else self.world_size

    @property
    def seq_rank(self) -> int:
        return self.rank % self.seq_world

    @property
    def cfg_rank(self) -> int:
        """0: this rank runs the prompt forward, 1: the negative-prompt forward (cfg_parallel only)."""
        return self.rank // self.seq_world

    @property
    def partner(self) -> int:
        return (self.rank + self.seq_world) % self.world_size

    @property
    def group_leader(self) -> int:
        return self.cfg_rank * self.seq_world

    def describe(self) -> str:
        if self.world_size == 1:
            return "single GPU"
        if self.cfg_parallel:
            return f"cfg x2 (prompt | negative) x temporal-token shard x{self.seq_world}"
        return f"temporal-token shard x{self.world_size}"


def exchange_nccl_unique_id(layout: ParallelLayout, device) -> Optional[bytes]:
    """Collective over torch.distributed: the leader of every temporal-shard group draws an ncclUniqueId and each
    rank returns its own group's (None when the group is a single rank)."""
    if layout.seq_world == 1:
        return None
    import torch.distributed as dist
    mine = torch.zeros(128, dtype=torch.uint8)
    if layout.rank == layout.group_leader:
        buf = C.create_string_buffer(128)
        check(lib().ic_nccl_unique_id(buf), "ic_nccl_unique_id")
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    mine = mine.to(device)
    ids = [torch.empty_like(mine) for _ in range(layout.world_size)]
    dist.all_gather(ids, mine)
    return bytes(ids[layout.group_leader].cpu().tolist())


def group_blobs(blobs: List[bytes], layout: ParallelLayout) -> bytes:
    """From one fixed-size blob per world rank (rank order) keep the blobs of this rank's temporal-shard group, in
    group-rank order, concatenated - the layout ic_dit_p2p_attach expects."""
    if len(blobs) != layout.world_size:
        raise ValueError(f"{len(blobs)} blobs for a world of {layout.world_size}")
    lead = layout.group_leader
    return b"".join(blobs[lead:lead + layout.seq_world])


def p2p_requested() -> bool:
    """The peer-memory K / V^T exchange is the default for a temporal-shard group; ICB_KV_P2P=0 selects
    the NCCL all-gather instead."""
    return os.environ.get("ICB_KV_P2P", "1") != "0"


def exchange_p2p_handles(layout: ParallelLayout, engine: "WanDiTEngine") -> bool:
    """Collective over torch.distributed: every rank exports the IPC handles of its gather buffer and flags, the 128-
    byte blobs (plus an "export worked" byte) are all-gathered, and - only if EVERY rank could export (stream memory
    operations and IPC available) - each rank attaches to the peers of its own temporal-shard group; a barrier
    guarantees nobody pushes before everybody is attached.  Returns whether the push path is active (same answer on
    every rank)."""
    if layout.seq_world == 1:
        return False
    import torch.distributed as dist
    try:
        blob, ok = engine.p2p_export(), 1
    except ICError:
        blob, ok = bytes(128), 0
    mine = torch.frombuffer(bytearray(blob + bytes([ok])), dtype=torch.uint8).clone().to(engine.device)
    parts = [torch.empty_like(mine) for _ in range(layout.world_size)]
    dist.all_gather(parts, mine)
    raw = [bytes(t.cpu().tolist()) for t in parts]
    if not all(r[128] for r in raw):
        return False
    engine.p2p_attach(group_blobs([r[:128] for r in raw], layout))
    dist.barrier()
    return True


def setup_kv_exchange(layout: ParallelLayout, engine: "WanDiTEngine", device) -> str:
    """Collective over torch.distributed: wires the per-layer (K || V^T) exchange of a temporal-shard group - the
    peer-memory push when requested and available, else the NCCL all-gather - and says which one is active."""
    if layout.seq_world == 1:
        return "none"
    if p2p_requested() and exchange_p2p_handles(layout, engine):
        return "peer-memory push"
    engine.init_comm(exchange_nccl_unique_id(layout, device))
    return "nccl all-gather"


class WanDiTEngine:
    """Owns an `ic_dit` handle (weights + workspaces resident in HBM)."""

    def __init__(self, cfg: WanModelConfig, lat_f: int, lat_h: int, lat_w: int, guide_channels: int = 32,
                 world_size: int = 1, rank: int = 0, device: Optional[torch.device] = None):
        require_device()
        self.cfg = cfg
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.lat = (lat_f, lat_h, lat_w)
        self.world_size, self.rank = world_size, rank
        self.frame0, self.frames_local = shard_frames(lat_f, world_size, rank)
        self.guide_channels = guide_channels
        c = DitConfig(cfg.dim, cfg.ffn_dim, cfg.num_heads, cfg.num_layers, cfg.in_dim, cfg.out_dim, cfg.text_dim,
                      cfg.freq_dim, cfg.text_len, guide_channels, cfg.eps, lat_f, lat_h, lat_w, self.frame0,
                      self.frames_local, world_size, rank)
        h = C.c_void_p()
        check(lib().ic_dit_create(C.byref(c), C.byref(h)), "ic_dit_create")
        self._h = h
        self.tokens_local = self.frames_local * (lat_h // 2) * (lat_w // 2)
        self.tokens_total = lat_f * (lat_h // 2) * (lat_w // 2)
        self.loaded = set()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().ic_dit_destroy(h)
            except Exception:  # noqa: BLE001
                pass
            self._h = None

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        unexpected = []
        for k, v in sd.items():
            t = v.detach().to(self.device)
            if t.dtype == torch.bfloat16:
                dt = _lib.IC_DTYPE_BF16
            else:
                t = t.to(torch.float32)
                dt = _lib.IC_DTYPE_F32
            t = t.contiguous()
            r = lib().ic_dit_load_tensor(self._h, k.encode(), C.c_void_p(t.data_ptr()), dt, t.numel(), self._stream())
            if r != 0:
                unexpected.append(k)
            else:
                self.loaded.add(k)
        torch.cuda.current_stream().synchronize()
        if strict and unexpected:
            raise KeyError(f"unexpected or mis-shaped keys: {unexpected[:8]}{'...' if len(unexpected) > 8 else ''}")
        return unexpected

    def init_comm(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        check(lib().ic_dit_init_comm(self._h, buf), "ic_dit_init_comm")

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().ic_dit_p2p_export(self._h, buf), "ic_dit_p2p_export")
        return buf.raw

    def p2p_attach(self, group_handles: bytes):
        if len(group_handles) != 128 * self.world_size:
            raise ValueError(f"expected {128 * self.world_size} bytes of handles, got {len(group_handles)}")
        buf = C.create_string_buffer(group_handles, len(group_handles))
        check(lib().ic_dit_p2p_attach(self._h, buf), "ic_dit_p2p_attach")

    @property
    def p2p_enabled(self) -> bool:
        return bool(lib().ic_dit_p2p_enabled(self._h))

    def set_context(self, slot: int, ctx: torch.Tensor):
        t = ctx.to(self.device).contiguous()
        dt = _lib.IC_DTYPE_BF16 if t.dtype == torch.bfloat16 else _lib.IC_DTYPE_F32
        if dt == _lib.IC_DTYPE_F32:
            t = t.to(torch.float32)
        check(lib().ic_dit_set_context(self._h, slot, C.c_void_p(t.data_ptr()), dt, self._stream()), "ic_dit_set_context")
        torch.cuda.current_stream().synchronize()  # t may be a temporary

    def set_guidance(self, guide_latents: Optional[torch.Tensor]):
        if guide_latents is None:
            check(lib().ic_dit_set_guidance(self._h, None, self._stream()), "ic_dit_set_guidance")
            return
        t = guide_latents.to(self.device, torch.float32).contiguous()
        check(lib().ic_dit_set_guidance(self._h, C.c_void_p(t.data_ptr()), self._stream()), "ic_dit_set_guidance")
        torch.cuda.current_stream().synchronize()

    def forward(self, latents: torch.Tensor, timestep: float, slot: int, head_out: torch.Tensor):
        """latents fp32 [C, frames_local, H, W] -> head_out fp32 [tokens_local, 4*out_dim]."""
        check(lib().ic_dit_forward(self._h, C.c_void_p(latents.data_ptr()), float(timestep), slot,
                                   C.c_void_p(head_out.data_ptr()), self._stream()), "ic_dit_forward")

    # parity hooks
    def embed(self, latents: torch.Tensor, timestep: float):
        check(lib().ic_dit_embed(self._h, C.c_void_p(latents.data_ptr()), float(timestep), self._stream()), "ic_dit_embed")

    def run_block(self, layer: int, slot: int):
        check(lib().ic_dit_run_block(self._h, layer, slot, self._stream()), "ic_dit_run_block")

    def run_block_phase(self, layer: int, slot: int, phase: int):
        """Parity hook: phase 0 produces q and this rank's (K || V^T) segment, phase 1 runs the attention over all
        segments (no collective) and the rest of the block."""
        check(lib().ic_dit_run_block_phase(self._h, layer, slot, phase, self._stream()), "ic_dit_run_block_phase")

    def kv_segment(self, rank: int) -> torch.Tensor:
        """Zero-copy uint8 view of rank `rank`'s (K || V^T) segment inside this engine's gather buffer."""
        ptr, nbytes = C.c_void_p(), C.c_longlong()
        check(lib().ic_dit_kv_segment(self._h, rank, C.byref(ptr), C.byref(nbytes)), "ic_dit_kv_segment")
        return _wrap_device(ptr.value, nbytes.value, "|u1", self.device)

    def missing_tensors(self) -> List[str]:
        """Registered weights nobody loaded yet (the engine's storage is plain cudaMalloc: running on it would be
        running on garbage)."""
        cap = 1 << 20
        buf = C.create_string_buffer(cap)
        n = lib().ic_dit_missing_tensors(self._h, buf, cap)
        if n < 0:
            check(n, "ic_dit_missing_tensors")
        return [x for x in buf.value.decode().split("\n") if x]

    def head(self, head_out: torch.Tensor):
        check(lib().ic_dit_head(self._h, C.c_void_p(head_out.data_ptr()), self._stream()), "ic_dit_head")

    def tokens(self) -> torch.Tensor:
        """View of the engine's fp32 residual stream [tokens_local, dim] (copy)."""
        ptr = lib().ic_dit_tokens(self._h)
        n = self.tokens_local * self.cfg.dim
        out = torch.empty(self.tokens_local, self.cfg.dim, dtype=torch.float32, device=self.device)
        # device-to-device copy through a zero-copy torch view of the engine buffer
        src = _wrap_device_f32(ptr, n, self.device)
        out.view(-1).copy_(src)
        return out

    def set_tokens(self, x: torch.Tensor):
        ptr = lib().ic_dit_tokens(self._h)
        n = self.tokens_local * self.cfg.dim
        _wrap_device_f32(ptr, n, self.device).copy_(x.to(self.device, torch.float32).contiguous().view(-1))

    PROF_FMHA_SELF, PROF_FMHA_CROSS, PROF_GEMM, PROF_ALL = 1, 2, 4, 7

    def set_profiling(self, kinds: int):
        """Bit mask of launch kinds to bracket with CUDA events (0 = off; True counts as the self-attention only)."""
        check(lib().ic_dit_set_profiling(self._h, int(kinds)), "ic_dit_set_profiling")

    def profile_collect(self):
        """-> {kind: (total_ms, launches)} for kinds fmha_self / fmha_cross / gemm (synchronises)."""
        ms = (C.c_float * 3)()
        cnt = (C.c_int * 3)()
        check(lib().ic_dit_profile_collect(self._h, ms, cnt), "ic_dit_profile_collect")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(("fmha_self", "fmha_cross", "gemm"))}

    @property
    def flops_per_forward(self) -> int:
        return int(lib().ic_dit_flops_per_forward(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().ic_dit_launch_count(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(lib().ic_dit_workspace_bytes(self._h))


def _wrap_device(ptr: int, numel: int, typestr: str, device) -> torch.Tensor:
    """Zero-copy torch view of engine-owned device memory (via __cuda_array_interface__)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=device)


def _wrap_device_f32(ptr: int, numel: int, device) -> torch.Tensor:
    return _wrap_device(ptr, numel, "<f4", device)


# --------------------------------------------------------------------------------------------------
# denoising loop
# --------------------------------------------------------------------------------------------------
def model_timestep(t, dtype: torch.dtype = torch.bfloat16) -> float:
    """The timestep value the DiT sees: the scheduler's fp32 timestep rounded through the pipeline dtype."""
    return float(torch.as_tensor(float(t), dtype=torch.float32).to(dtype))


class DenoiseLoop:
    """The hot loop of WanVideoPipeline.__call__: per step two DiT forwards (prompt / negative prompt), CFG
    combine and the flow-match Euler update, all on the current CUDA stream."""

    def __init__(self, engine: WanDiTEngine, cfg_scale: float = 5.0, layout: Optional[ParallelLayout] = None):
        self.e = engine
        self.cfg_scale = cfg_scale
        self.layout = layout or ParallelLayout(engine.world_size, engine.rank, False)
        if self.layout.seq_world != engine.world_size or self.layout.seq_rank != engine.rank:
            raise ValueError("engine shard does not match the parallel layout")
        n = engine.tokens_local * 4 * engine.cfg.out_dim
        self.head_pos = torch.empty(n, dtype=torch.float32, device=engine.device)
        self.head_neg = torch.empty(n, dtype=torch.float32, device=engine.device)

    @property
    def forwards_per_step(self) -> int:
        return 1 if self.layout.cfg_parallel else 2

    def step(self, latents: torch.Tensor, timestep: float, dsigma: float):
        e = self.e
        if self.layout.cfg_parallel:
            import torch.distributed as dist
            heads = (self.head_pos, self.head_neg)
            own, other = heads[self.layout.cfg_rank], heads[1 - self.layout.cfg_rank]
            e.forward(latents, timestep, self.layout.cfg_rank, own)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, own, self.layout.partner),
                                           dist.P2POp(dist.irecv, other, self.layout.partner)])
            for r in reqs:
                r.wait()
        else:
            e.forward(latents, timestep, 0, self.head_pos)
            e.forward(latents, timestep, 1, self.head_neg)
        Cc = e.cfg.out_dim
        _, H, W = e.lat
        check(lib().ic_unpatchify_cfg_step(C.c_void_p(latents.data_ptr()), C.c_void_p(self.head_pos.data_ptr()),
                                           C.c_void_p(self.head_neg.data_ptr()), Cc, e.frames_local, H, W,
                                           float(self.cfg_scale), float(dsigma), None, e._stream()),
              "ic_unpatchify_cfg_step")

    def run(self, latents: torch.Tensor, scheduler: FlowMatchScheduler, steps: Optional[int] = None,
            timestep_dtype: torch.dtype = torch.bfloat16):
        """`timestep_dtype`: diffsynth's pipeline hands the model `timestep.to(dtype=pipe.torch_dtype)` (bfloat16 in
        the reference, videogen/inference.py:44), so 995.9 reaches the sinusoidal embedding as 996 - the step sizes
        (sigmas) stay fp32.  Mirrored by default (un-vendored dependency: restated from the published pipeline, see
        DESIGN.md §2); pass torch.float32 for the exact schedule."""
        n = len(scheduler.timesteps) if steps is None else steps
        for i in range(n):
            self.step(latents, model_timestep(scheduler.timesteps[i], timestep_dtype), scheduler.delta_sigma(i))
        return latents


# --------------------------------------------------------------------------------------------------
# pipeline object with the surface videogen/inference.py consumes
# --------------------------------------------------------------------------------------------------
class _StateDictSink:
    """`pipe.dit` / `pipe.buffer_embedder`: accepts load_state_dict like an nn.Module and forwards the tensors
    to the engine (immediately if it exists, else when it is built)."""

    def __init__(self, pipe: "WanVideoPipeline", prefix: str):
        self._pipe, self._prefix = pipe, prefix

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        if self._prefix == "buffer_embedder.":
            state_dict = map_buffer_embedder(state_dict, self._pipe.model_cfg.dim, self._pipe.buffer_channels)
        self._pipe._stage_weights({self._prefix + k: v for k, v in state_dict.items()}, strict)


def map_buffer_embedder(sd: Dict[str, torch.Tensor], dim: int, buffer_channels: int) -> Dict[str, torch.Tensor]:
    """Checkpoint `buffer_embedder.*` tensors -> the engine's one patch-embedding conv over the channel-concatenated
    (semantic, coordinate) latents: {weight [dim, 2C, 1, 2, 2], bias [dim]}.  The embedder's layout lives in the
    reference's un-vendored diffsynth fork (`pipe.initialize_buffer_embedder`, videogen/inference.py:86-88; SURVEY
    A.10), so the shapes are INSPECTED here instead of assumed.  Accepted: (a) exactly that single conv; (b) two convs
    over C channels each whose names tell semantic from coordinate (their sum over separate inputs is the single
    conv over the concatenated input, biases added).  Anything else raises with the shapes found - never a silent
    mis-load."""
    C2 = 2 * buffer_channels
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    w = [k for k, v in sd.items() if v.ndim == 5]
    b = [k for k, v in sd.items() if v.ndim == 1]
    if len(w) == 1 and shapes[w[0]] == (dim, C2, 1, 2, 2) and len(b) <= 1 and len(sd) == len(w) + len(b):
        out = {"weight": sd[w[0]]}
        out["bias"] = sd[b[0]] if b else torch.zeros(dim, dtype=sd[w[0]].dtype, device=sd[w[0]].device)
        if out["bias"].shape != (dim,):
            raise KeyError(f"buffer_embedder bias has shape {tuple(out['bias'].shape)}, expected ({dim},)")
        return out
    sem_w = [k for k in w if "sem" in k.lower()]
    crd_w = [k for k in w if "coord" in k.lower()]
    if len(w) == 2 and len(sem_w) == 1 and len(crd_w) == 1 and all(
            shapes[k] == (dim, buffer_channels, 1, 2, 2) for k in w) and len(sd) == len(w) + len(b):
        weight = torch.cat([sd[sem_w[0]], sd[crd_w[0]]], dim=1)
        bias = torch.zeros(dim, dtype=torch.float32, device=weight.device)
        for k in b:
            if shapes[k] != (dim,):
                raise KeyError(f"buffer_embedder tensor {k} has shape {shapes[k]}, expected ({dim},)")
            bias = bias + sd[k].to(bias)
        return {"weight": weight, "bias": bias.to(weight.dtype)}
    raise KeyError(
        f"buffer_embedder layout not understood: expected one Conv3d weight ({dim}, {C2}, 1, 2, 2) [+ bias ({dim},)] or "
        f"a semantic / coordinate pair of ({dim}, {buffer_channels}, 1, 2, 2) convs; the checkpoint holds {shapes}")


class WanVideoPipeline:
    def __init__(self, device="cuda:0", torch_dtype=torch.bfloat16, model_cfg: Optional[WanModelConfig] = None,
                 world_size: int = 1, rank: int = 0, cfg_parallel: Optional[bool] = None):
        self.device = torch.device(device)
        self.torch_dtype = torch_dtype
        self.model_cfg = model_cfg or WanModelConfig.wan_14b()
        self.world_size, self.rank = world_size, rank
        self.layout = ParallelLayout.make(world_size, rank, cfg_parallel)
        self.scheduler = FlowMatchScheduler(shift=5.0)
        self.buffer_channels = 0
        self.buffer_embedder: Optional[_StateDictSink] = None
        self.dit = _StateDictSink(self, "")
        self.vae = None           # set by from_pretrained when a VAE checkpoint / synthetic VAE is available
        self.text_encoder = None  # WanTextEncoder (umT5-XXL, row A11) once its checkpoint is attached
        self.prompter = None      # WanPrompter (tokenizer + text_encoder); None -> deterministic synthetic contexts
        self._ctx_cache: Dict[str, torch.Tensor] = {}
        self._weights: Dict[str, torch.Tensor] = {}
        self._engine: Optional[WanDiTEngine] = None
        self._engine_key = None
        self._nccl_id: Optional[bytes] = None
        self.synthetic = False

    # ---- construction ---------------------------------------------------------------------------
    @staticmethod
    def from_pretrained(torch_dtype=torch.bfloat16, device="cuda:0", model_configs: Sequence[ModelConfig] = (),
                        synthetic_weights: Optional[bool] = None, world_size: int = 1, rank: int = 0,
                        cfg_parallel: Optional[bool] = None, **_ignored) -> "WanVideoPipeline":
        require_device()
        ids = " ".join(str(getattr(m, "model_id", "")) for m in model_configs)
        cfg = WanModelConfig.wan_1_3b() if "1.3B" in ids else WanModelConfig.wan_14b()
        pipe = WanVideoPipeline(device, torch_dtype, cfg, world_size, rank, cfg_parallel)
        if synthetic_weights is None:
            synthetic_weights = os.environ.get("INFINICUBE_B200_SYNTHETIC", "0") == "1"
        files = [f for m in model_configs for f in (m.resolve() if isinstance(m, ModelConfig) else [])
                 if f.endswith(".safetensors")]
        if files:
            from safetensors.torch import load_file
            for f in files:
                pipe._stage_weights(load_file(f, device=str(pipe.device)), strict=False)
            vae_files = [f for m in model_configs for f in (m.resolve() if isinstance(m, ModelConfig) else [])
                         if f.endswith("VAE.pth")]
            if vae_files:
                from .vae import WanVideoVAE
                vsd = torch.load(vae_files[0], map_location="cpu", weights_only=True)
                vsd = {k.replace("model.", "", 1) if k.startswith("model.") else k: v for k, v in vsd.items()}
                pipe.vae = WanVideoVAE(vsd, device=pipe.device, world_size=world_size, rank=rank)
            t5_files = [f for m in model_configs for f in (m.resolve() if isinstance(m, ModelConfig) else [])
                        if os.path.basename(f).startswith("models_t5_umt5-xxl")]
            if t5_files:
                pipe.attach_text_encoder(t5_files[0], os.path.join(os.path.dirname(t5_files[0]), "google", "umt5-xxl"))
        elif synthetic_weights:
            pipe.synthetic = True
            pipe._stage_weights(synthetic_state_dict(cfg, 0, pipe.device), strict=False)
            from .vae import WanVideoVAE, synthetic_vae_state_dict
            pipe.vae = WanVideoVAE(synthetic_vae_state_dict(), device=pipe.device, world_size=world_size, rank=rank)
        else:
            raise FileNotFoundError(
                "Wan2.1 DiT weights not found under ./models (expected "
                f"{[getattr(m, 'origin_file_pattern', None) for m in model_configs]}); set "
                "INFINICUBE_B200_SYNTHETIC=1 or synthetic_weights=True to run with random-init weights")
        return pipe

    def initialize_buffer_embedder(self, buffer_channels: int = 16, zero_init: bool = True):
        """Conv3d(2*buffer_channels -> dim, kernel = stride = (1,2,2)) on the channel-concatenated (semantic,
        coordinate) VAE latents, zero-initialised so that an untrained embedder reproduces plain Wan2.1
        (videogen/inference.py:86-88; layout is this repo's choice)."""
        self.buffer_channels = buffer_channels
        D = self.model_cfg.dim
        w = torch.zeros(D, 2 * buffer_channels, 1, 2, 2, device=self.device, dtype=torch.bfloat16)
        b = torch.zeros(D, device=self.device, dtype=torch.bfloat16)
        if not zero_init:
            w.normal_(0, 0.02)
        self._weights["buffer_embedder.weight"] = w
        self._weights["buffer_embedder.bias"] = b
        self.buffer_embedder = _StateDictSink(self, "buffer_embedder.")
        self._engine = None

    def enable_vram_management(self):
        """No-op: all weights stay resident (1.3B: 2.8 GB, 14B: 28 GB of 180 GB HBM)."""

    def set_nccl_unique_id(self, uid: bytes):
        self._nccl_id = uid

    def _stage_weights(self, sd: Dict[str, torch.Tensor], strict: bool):
        for k, v in sd.items():
            self._weights[k] = v
        if self._engine is not None:
            self._engine.load_state_dict(sd, strict=False)

    # ---- engine -----------------------------------------------------------------------------------
    def engine_for(self, lat_f: int, lat_h: int, lat_w: int) -> WanDiTEngine:
        key = (lat_f, lat_h, lat_w, 2 * self.buffer_channels)
        if self._engine is None or self._engine_key != key:
            eng = WanDiTEngine(self.model_cfg, lat_f, lat_h, lat_w, 2 * self.buffer_channels, self.layout.seq_world,
                               self.layout.seq_rank, self.device)
            bad = eng.load_state_dict(self._weights, strict=False)
            bad = [k for k in bad if not k.startswith(("vae.", "text_encoder."))]
            if bad:
                raise KeyError(f"state-dict keys the DiT engine does not know: {bad[:8]}")
            missing = eng.missing_tensors()
            if not self.buffer_channels:
                missing = [k for k in missing if not k.startswith("buffer_embedder.")]
            if missing:  # a missing shard / renamed key must not run on uninitialised HBM
                raise KeyError(f"{len(missing)} DiT tensors were never loaded: {missing[:8]}"
                               f"{'...' if len(missing) > 8 else ''}")
            if self.layout.seq_world > 1:
                if self._nccl_id is not None:
                    eng.init_comm(self._nccl_id)
                else:
                    import torch.distributed as dist
                    if not (dist.is_available() and dist.is_initialized()):
                        raise ICError("multi-GPU pipeline needs torch.distributed initialised (or "
                                      "set_nccl_unique_id()) before the first call")
                    setup_kv_exchange(self.layout, eng, self.device)
            self._engine, self._engine_key = eng, key
        return self._engine

    def attach_text_encoder(self, checkpoint, tokenizer_path: Optional[str] = None, t5_cfg=None):
        """Loads the umT5 prompt encoder (`models_t5_umt5-xxl-enc-bf16.pth`, or an already loaded state dict, Wan2.1
        or transformers key names) onto this pipeline's device and the tokenizer from `tokenizer_path` (the
        `google/umt5-xxl` directory that ships with the Wan2.1 checkpoints).  Row A11 of SURVEY §8(a)."""
        from .text_encoder import T5Config, WanPrompter, WanTextEncoder
        sd = torch.load(checkpoint, map_location="cpu", weights_only=True) if isinstance(checkpoint, str) else checkpoint
        cfg = t5_cfg or T5Config(text_len=self.model_cfg.text_len)
        if cfg.dim != self.model_cfg.text_dim:
            raise ValueError(f"text encoder width {cfg.dim} does not match the DiT's text_dim {self.model_cfg.text_dim}")
        enc = WanTextEncoder(cfg, self.device)
        enc.load_state_dict(sd, strict=True)
        self.text_encoder = enc
        self.prompter = WanPrompter(text_len=cfg.text_len)
        self.prompter.fetch_models(enc)
        if tokenizer_path is not None and os.path.isdir(tokenizer_path):
            self.prompter.fetch_tokenizer(tokenizer_path)
        self._ctx_cache.clear()
        return enc

    def encode_prompt(self, prompt: str) -> torch.Tensor:
        """Prompt -> context [text_len, text_dim] bf16 (cached per prompt: the reference re-encodes both prompts on
        every call, the result only depends on the string)."""
        if self.prompter is not None and self.prompter.tokenizer is not None:
            if prompt not in self._ctx_cache:
                if len(self._ctx_cache) >= 16:
                    self._ctx_cache.clear()
                self._ctx_cache[prompt] = self.prompter.encode_prompt(prompt)
            return self._ctx_cache[prompt]
        if not (self.synthetic or os.environ.get("INFINICUBE_B200_SYNTHETIC_CONTEXT", "0") == "1"):
            # real Wan weights conditioned on noise would produce plausible but wrong video; the reference fails to
            # load without its text encoder (videogen/inference.py:63-81)
            raise FileNotFoundError(
                "no umT5 text encoder / tokenizer attached (models_t5_umt5-xxl-enc-bf16.pth and the google/umt5-xxl "
                "tokenizer directory next to it): call attach_text_encoder(), or set "
                "INFINICUBE_B200_SYNTHETIC_CONTEXT=1 to condition on deterministic synthetic contexts")
        return synthetic_context(prompt, self.model_cfg, self.device)

    def encode_prompt_ids(self, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """Tokenised prompt -> context (the entry for callers that tokenise themselves)."""
        if self.prompter is None:
            raise ICError("no text encoder attached: call attach_text_encoder() first")
        return self.prompter.encode_ids(ids, mask)

    @torch.no_grad()
    def denoise(self, noise: torch.Tensor, ctx_pos: torch.Tensor, ctx_neg: torch.Tensor,
                guide_latents: Optional[torch.Tensor], num_inference_steps: int = 50, sigma_shift: float = 5.0,
                cfg_scale: float = 5.0) -> torch.Tensor:
        """noise fp32 [16, F, H, W] (global) -> denoised latents for this rank's frames."""
        _, F, H, W = noise.shape
        eng = self.engine_for(F, H, W)
        f0, fl = eng.frame0, eng.frames_local
        eng.set_context(0, ctx_pos)
        eng.set_context(1, ctx_neg)
        eng.set_guidance(None if guide_latents is None else guide_latents[:, f0:f0 + fl])
        lat = noise[:, f0:f0 + fl].to(self.device, torch.float32).contiguous()
        self.scheduler.set_timesteps(num_inference_steps, shift=sigma_shift)
        DenoiseLoop(eng, cfg_scale, self.layout).run(lat, self.scheduler)
        return lat

    @torch.no_grad()
    def __call__(self, prompt: str = "", negative_prompt: str = "", semantic_buffer_video=None,
                 coordinate_buffer_video=None, height: int = 480, width: int = 832, num_frames: int = 81,
                 seed: Optional[int] = None, tiled: bool = True, cfg_scale: float = 5.0,
                 num_inference_steps: int = 50, sigma_shift: float = 5.0, rand_device: str = "cpu",
                 output_type: str = "pil"):
        if height % 16 or width % 16:
            raise ValueError(f"height and width must be multiples of 16, got {height}x{width}")
        if num_frames % 4 != 1:
            raise ValueError(f"num_frames must satisfy T % 4 == 1, got {num_frames}")
        if self.vae is None:
            raise ICError("this pipeline has no VAE attached: use denoise() with latent-space guidance, or attach "
                          "pipe.vae (WanVideoVAE) to run from uint8 buffers to frames")
        lat_shape = (16, (num_frames - 1) // 4 + 1, height // 8, width // 8)
        g = torch.Generator(device=rand_device)
        if seed is not None:
            g.manual_seed(seed)
        noise = torch.randn(lat_shape, generator=g, device=rand_device, dtype=torch.float32)
        guide = None
        if semantic_buffer_video is not None and coordinate_buffer_video is not None and self.buffer_channels:
            z_s, z_c = self.vae.encode_frames_many([semantic_buffer_video, coordinate_buffer_video], tiled=tiled)
            guide = torch.cat([z_s, z_c], dim=0)
        lat = self.denoise(noise, self.encode_prompt(prompt), self.encode_prompt(negative_prompt), guide,
                           num_inference_steps, sigma_shift, cfg_scale)
        if self.world_size > 1:
            import torch.distributed as dist
            parts = [torch.empty_like(lat) for _ in range(self.world_size)]
            dist.all_gather(parts, lat)
            lat = torch.cat(parts[:self.layout.seq_world], dim=1)  # both CFG groups hold the same latents
        return self.vae.decode_to_frames(lat, tiled=tiled, output_type=output_type)
