"""CPU fp32 restatement of the Wan2.1 DiT denoising path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (infinicube_b200/) never does.

PARITY UNPINNED for the DiT arithmetic: the reference executes it inside the un-vendored, un-pinned
dependency `diffsynth @ git+https://github.com/yifanlu0227/DiffSynth-Studio-InfiniCube`
(/root/reference/pyproject.toml:71; imported at infinicube/videogen/inference.py:25-26), which is
absent from /root/reference and from this image, and the reference holds no golden vectors for it
(SURVEY.md §4, §8c).  This file restates the published Wan2.1 / DiffSynth algorithm (SURVEY.md
Appendix A.1-A.8) and anchors on the reference's own call site
(infinicube/videogen/inference.py:216-226) and the scheduler known-answers of Appendix A.7.
What *is* pinned: the flow-match schedule (tests/test_oracle_dit.py checks the A.7 values) and the
zero-init invariant of the buffer embedder (inference.py:86-88).

Everything is plain PyTorch fp32 on CPU (fp64 where the reference uses fp64: sinusoid, RoPE).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class WanConfig:
    """Appendix A.1 hyper-parameters (Wan2.1-T2V)."""
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
    guide_channels: int = 32  # two 16-channel VAE latents (semantic + coordinate buffers)

    @staticmethod
    def wan_1_3b(**kw) -> "WanConfig":
        return WanConfig(**kw)

    @staticmethod
    def wan_14b(**kw) -> "WanConfig":
        return WanConfig(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40, **kw)

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads


# --------------------------------------------------------------------------------------------------
# synthetic weights (official Wan key names, Appendix A.6); values are bf16-representable because the
# reference runs the model with torch_dtype=bfloat16 (inference.py:45,63).
# --------------------------------------------------------------------------------------------------
def _bf16r(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def make_weights(cfg: WanConfig, seed: int = 1234, layers: Optional[int] = None,
                 zero_guidance: bool = False, random_norms: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic random-init state dict (Linear ~ N(0, 0.02^2), modulation ~ N(0,1)/sqrt(D));
    created in the key order of Appendix A.6.  §8d says "norm weights 1"; with `random_norms` (default) the RMSNorm /
    LayerNorm affine parameters are drawn as 1 + 0.25 N(0,1) (bias 0.1 N(0,1)) from a SEPARATE generator, so that a
    wrong index or stride in a norm-weight path cannot hide behind all-ones weights while every other tensor keeps
    the values it had with unit norms."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    gn = torch.Generator(device="cpu").manual_seed(seed + 7919)

    def norm_w():
        return _bf16r(1.0 + 0.25 * torch.randn(D, generator=gn)) if random_norms else torch.ones(D)

    def norm_b():
        return _bf16r(0.1 * torch.randn(D, generator=gn)) if random_norms else torch.zeros(D)

    D, Fd = cfg.dim, cfg.ffn_dim
    L = cfg.num_layers if layers is None else layers
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, std=0.02):
        sd[name + ".weight"] = _bf16r(torch.randn(out_f, in_f, generator=g) * std)
        sd[name + ".bias"] = _bf16r(torch.randn(out_f, generator=g) * std)

    sd["patch_embedding.weight"] = _bf16r(torch.randn(D, cfg.in_dim, 1, 2, 2, generator=g) * 0.02)
    sd["patch_embedding.bias"] = _bf16r(torch.randn(D, generator=g) * 0.02)
    lin("text_embedding.0", D, cfg.text_dim)
    lin("text_embedding.2", D, D)
    lin("time_embedding.0", D, cfg.freq_dim)
    lin("time_embedding.2", D, D)
    lin("time_projection.1", 6 * D, D)
    for i in range(L):
        p = f"blocks.{i}."
        for a in ("self_attn", "cross_attn"):
            for n in ("q", "k", "v", "o"):
                lin(p + f"{a}.{n}", D, D)
            sd[p + f"{a}.norm_q.weight"] = norm_w()
            sd[p + f"{a}.norm_k.weight"] = norm_w()
        sd[p + "norm3.weight"] = norm_w()
        sd[p + "norm3.bias"] = norm_b()
        lin(p + "ffn.0", Fd, D)
        lin(p + "ffn.2", D, Fd)
        sd[p + "modulation"] = _bf16r(torch.randn(1, 6, D, generator=g) / math.sqrt(D))
    lin("head.head", 4 * cfg.out_dim, D)
    sd["head.modulation"] = _bf16r(torch.randn(1, 2, D, generator=g) / math.sqrt(D))
    if cfg.guide_channels > 0:
        if zero_guidance:  # initialize_buffer_embedder(zero_init=True), inference.py:86-88
            sd["buffer_embedder.weight"] = torch.zeros(D, cfg.guide_channels, 1, 2, 2)
            sd["buffer_embedder.bias"] = torch.zeros(D)
        else:
            sd["buffer_embedder.weight"] = _bf16r(torch.randn(D, cfg.guide_channels, 1, 2, 2, generator=g) * 0.02)
            sd["buffer_embedder.bias"] = _bf16r(torch.randn(D, generator=g) * 0.02)
    return sd


# --------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """Appendix A.5: [cos(t w_i) || sin(t w_i)], w_i = 10000^(-i/(dim/2)), computed in fp64."""
    half = dim // 2
    pos = position.to(torch.float64)
    w = torch.pow(torch.tensor(10000.0, dtype=torch.float64), -torch.arange(half, dtype=torch.float64) / half)
    ang = torch.outer(pos, w)
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1).to(torch.float32)


def rope_angles(f: int, h: int, w: int, head_dim: int = 128, frame0: int = 0, frames_total: Optional[int] = None
                ) -> torch.Tensor:
    """Appendix A.4: 64 complex pairs split 22|21|21 over (frame, height, width);
    theta_j = 10000^(-2j/axis_dim).  Returns angles [f*h*w, head_dim/2] in fp64."""
    d = head_dim
    f_dim = d - 2 * (d // 3)
    h_dim = w_dim = d // 3

    def ax(n, axis_dim, start=0):
        th = torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                       -torch.arange(0, axis_dim, 2, dtype=torch.float64)[: axis_dim // 2] / axis_dim)
        return torch.outer(torch.arange(start, start + n, dtype=torch.float64), th)

    af, ah, aw = ax(f, f_dim, frame0), ax(h, h_dim), ax(w, w_dim)
    ang = torch.cat([
        af[:, None, None, :].expand(f, h, w, -1),
        ah[None, :, None, :].expand(f, h, w, -1),
        aw[None, None, :, :].expand(f, h, w, -1),
    ], dim=-1)
    return ang.reshape(f * h * w, d // 2)


def rope_apply(x: torch.Tensor, ang: torch.Tensor, num_heads: int) -> torch.Tensor:
    """Rotate adjacent pairs (x_{2j}, x_{2j+1}) per head; fp64 like the reference's complex128."""
    S, D = x.shape
    xr = x.to(torch.float64).view(S, num_heads, D // num_heads // 2, 2)
    c, s = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
    a, b = xr[..., 0], xr[..., 1]
    out = torch.stack([a * c - b * s, a * s + b * c], dim=-1)
    return out.reshape(S, D).to(torch.float32)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """Appendix A.3: RMSNorm over the whole model width."""
    return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps) * w


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int) -> torch.Tensor:
    Sq, D = q.shape
    Sk = k.shape[0]
    hd = D // num_heads
    qh = q.view(Sq, num_heads, hd).transpose(0, 1)
    kh = k.view(Sk, num_heads, hd).transpose(0, 1)
    vh = v.view(Sk, num_heads, hd).transpose(0, 1)
    o = F.scaled_dot_product_attention(qh[None], kh[None], vh[None])[0]
    return o.transpose(0, 1).reshape(Sq, D)


def dit_block(x: torch.Tensor, ctx: torch.Tensor, t_mod: torch.Tensor, sd: Dict[str, torch.Tensor], i: int,
              cfg: WanConfig, ang: torch.Tensor) -> torch.Tensor:
    """Appendix A.2 / A.3.  x [S, D], ctx [L, D] (post text_embedding), t_mod [6, D]."""
    p = f"blocks.{i}."
    D, H, eps = cfg.dim, cfg.num_heads, cfg.eps
    e = sd[p + "modulation"][0] + t_mod  # [6, D]

    def lin(name, t):
        return F.linear(t, sd[p + name + ".weight"], sd[p + name + ".bias"])

    # self attention
    xn = F.layer_norm(x, (D,), eps=eps) * (1 + e[1]) + e[0]
    q = rms_norm(lin("self_attn.q", xn), sd[p + "self_attn.norm_q.weight"], eps)
    k = rms_norm(lin("self_attn.k", xn), sd[p + "self_attn.norm_k.weight"], eps)
    v = lin("self_attn.v", xn)
    q, k = rope_apply(q, ang, H), rope_apply(k, ang, H)
    x = x + e[2] * lin("self_attn.o", attention(q, k, v, H))
    # cross attention (T2V: no mask over the 512 context rows)
    xn = F.layer_norm(x, (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], eps=eps)
    q = rms_norm(lin("cross_attn.q", xn), sd[p + "cross_attn.norm_q.weight"], eps)
    k = rms_norm(lin("cross_attn.k", ctx), sd[p + "cross_attn.norm_k.weight"], eps)
    v = lin("cross_attn.v", ctx)
    x = x + lin("cross_attn.o", attention(q, k, v, H))
    # feed forward
    xn = F.layer_norm(x, (D,), eps=eps) * (1 + e[4]) + e[3]
    x = x + e[5] * lin("ffn.2", F.gelu(lin("ffn.0", xn), approximate="tanh"))
    return x


def patchify(lat: torch.Tensor) -> torch.Tensor:
    """[C, F, H, W] -> [F*(H/2)*(W/2), C*4], column = c*4 + py*2 + px (im2col of Conv3d k=s=(1,2,2))."""
    C, Fr, H, W = lat.shape
    return lat.view(C, Fr, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5).reshape(Fr * (H // 2) * (W // 2), C * 4)


def unpatchify(tok: torch.Tensor, C: int, Fr: int, H: int, W: int) -> torch.Tensor:
    """Appendix A.5: b (f h w) (x y z c) -> b c (f x) (h y) (w z), (x,y,z) = (1,2,2)."""
    return tok.view(Fr, H // 2, W // 2, 2, 2, C).permute(5, 0, 1, 3, 2, 4).reshape(C, Fr, H, W)


def time_embed(t: float, sd: Dict[str, torch.Tensor], cfg: WanConfig) -> Tuple[torch.Tensor, torch.Tensor]:
    s = sinusoidal_embedding_1d(cfg.freq_dim, torch.tensor([t], dtype=torch.float64))[0]
    h = F.silu(F.linear(s, sd["time_embedding.0.weight"], sd["time_embedding.0.bias"]))
    t_emb = F.linear(h, sd["time_embedding.2.weight"], sd["time_embedding.2.bias"])
    t_mod = F.linear(F.silu(t_emb), sd["time_projection.1.weight"], sd["time_projection.1.bias"]).view(6, cfg.dim)
    return t_emb, t_mod


def text_embed(ctx_raw: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    h = F.gelu(F.linear(ctx_raw, sd["text_embedding.0.weight"], sd["text_embedding.0.bias"]), approximate="tanh")
    return F.linear(h, sd["text_embedding.2.weight"], sd["text_embedding.2.bias"])


def guidance_tokens(guide_lat: Optional[torch.Tensor], sd: Dict[str, torch.Tensor], cfg: WanConfig
                    ) -> Optional[torch.Tensor]:
    """Buffer-token injection (README.md:65; Appendix A.10): patch-embed the channel-concatenated
    (semantic, coordinate) latents with `buffer_embedder` -> additive tokens [S, D]."""
    if guide_lat is None:
        return None
    w = sd["buffer_embedder.weight"].reshape(cfg.dim, -1)
    return F.linear(patchify(guide_lat), w, sd["buffer_embedder.bias"])


def embed_tokens(lat: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: WanConfig,
                 guide: Optional[torch.Tensor]) -> torch.Tensor:
    x = F.linear(patchify(lat), sd["patch_embedding.weight"].reshape(cfg.dim, -1), sd["patch_embedding.bias"])
    if guide is not None:
        x = x + guide
    return x


def head(x: torch.Tensor, t_emb: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: WanConfig) -> torch.Tensor:
    e = sd["head.modulation"][0] + t_emb[None, :]  # [2, D]
    xn = F.layer_norm(x, (cfg.dim,), eps=cfg.eps) * (1 + e[1]) + e[0]
    return F.linear(xn, sd["head.head.weight"], sd["head.head.bias"])


def dit_forward(lat: torch.Tensor, t: float, ctx_raw: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: WanConfig,
                guide: Optional[torch.Tensor] = None, layers: Optional[int] = None, frame0: int = 0,
                return_tokens: bool = False):
    """WanModel.forward: lat [C, F, H, W] fp32 -> head output tokens [S, 4*out_dim]."""
    C, Fr, H, W = lat.shape
    t_emb, t_mod = time_embed(t, sd, cfg)
    ctx = text_embed(ctx_raw, sd)
    x = embed_tokens(lat, sd, cfg, guide)
    ang = rope_angles(Fr, H // 2, W // 2, cfg.head_dim, frame0)
    L = cfg.num_layers if layers is None else layers
    for i in range(L):
        x = dit_block(x, ctx, t_mod, sd, i, cfg, ang)
    out = head(x, t_emb, sd, cfg)
    return (out, x) if return_tokens else out


# --------------------------------------------------------------------------------------------------
# scheduler + CFG loop (Appendix A.7 / A.8)
# --------------------------------------------------------------------------------------------------
def flow_match_sigmas(num_steps: int = 50, shift: float = 5.0) -> torch.Tensor:
    """FlowMatchScheduler.set_timesteps(num_steps, denoising_strength=1, shift): fp32 like the reference.
    Returns num_steps+1 sigmas (last = 0)."""
    s = torch.linspace(1.0, 0.0, num_steps + 1, dtype=torch.float32)[:-1]
    sig = shift * s / (1 + (shift - 1) * s)
    return torch.cat([sig, torch.zeros(1)])


def denoise(lat: torch.Tensor, ctx_pos: torch.Tensor, ctx_neg: torch.Tensor, sd: Dict[str, torch.Tensor],
            cfg: WanConfig, guide: Optional[torch.Tensor], num_steps: int = 50, shift: float = 5.0,
            cfg_scale: float = 5.0, layers: Optional[int] = None, steps_to_run: Optional[int] = None,
            timestep_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """The hot loop of WanVideoPipeline.__call__.  The model receives the timestep cast to the pipeline
    dtype (`timestep.to(dtype=pipe.torch_dtype)` in diffsynth's pipeline; bfloat16 for the reference), the Euler update
    uses the fp32 sigmas."""
    sig = flow_match_sigmas(num_steps, shift)
    C, Fr, H, W = lat.shape
    n = num_steps if steps_to_run is None else steps_to_run
    for i in range(n):
        t = float((sig[i] * 1000.0).to(timestep_dtype))
        vp = dit_forward(lat, t, ctx_pos, sd, cfg, guide, layers)
        vn = dit_forward(lat, t, ctx_neg, sd, cfg, guide, layers)
        v = unpatchify(vn + cfg_scale * (vp - vn), cfg.out_dim, Fr, H, W)
        lat = lat + v * float(sig[i + 1] - sig[i])
    return lat


def dit_flops_per_forward(cfg: WanConfig, n_tokens: int) -> float:
    """SURVEY §8(d) algorithmic FLOP formula."""
    N, D, Fd, L = float(n_tokens), float(cfg.dim), float(cfg.ffn_dim), float(cfg.text_len)
    per_layer = 8 * N * D * D + 4 * N * N * D + 4 * N * D * D + 4 * L * D * D + 4 * N * L * D + 4 * N * D * Fd
    return cfg.num_layers * per_layer + 2 * N * (cfg.in_dim * 4) * D + 2 * N * D * (4 * cfg.out_dim)
