"""CPU fp32 restatement of the umT5-XXL prompt encoder of the Wan2.1 pipeline — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module; the product path
(infinicube_b200/) never does.

Row A11 of SURVEY.md §8(a): the reference encodes the prompt and the negative prompt once per call
inside the un-vendored `diffsynth` dependency (`prompters/wan_prompter.py`,
`models/wan_video_text_encoder.py`; call site infinicube/videogen/inference.py:216-226, models named at
:63-81 — `models_t5_umt5-xxl-enc-bf16.pth`).  diffsynth is absent from /root/reference and this image, so
the arithmetic below restates the published Wan2.1 `T5Encoder` (shared_pos=False, i.e. umT5: one
relative-position table per block, gated tanh-GELU feed-forward, RMS "T5LayerNorm", attention without
1/sqrt(d) scaling, right-padded prompts with padded keys masked, rows past the prompt length zeroed by
the prompter).

PINNED: tests/test_oracle_umt5.py checks this file against `transformers.UMT5EncoderModel` — an
independent, published implementation of the same umT5 architecture that ships in this image — on
randomly initialised small configurations (bit-for-bit same weights, max-abs <= 2e-5), against the
committed fixture tests/golden/umt5_small.npz that the same model produced
(tests/golden/gen_umt5_golden.py), and the bucket function against hand-derived known answers.
What stays unpinned is only the checkpoint key naming (taken from the published Wan2.1 t5.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch


@dataclass
class T5Config:
    """umT5-XXL encoder as instantiated by Wan2.1 (`umt5_xxl(encoder_only=True)`)."""
    vocab_size: int = 256384
    dim: int = 4096
    dim_attn: int = 4096
    dim_ffn: int = 10240
    num_heads: int = 64
    num_layers: int = 24
    num_buckets: int = 32
    max_dist: int = 128
    eps: float = 1e-6
    text_len: int = 512

    @property
    def head_dim(self) -> int:
        return self.dim_attn // self.num_heads


def relative_position_bucket(rel_pos: torch.Tensor, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """Bidirectional T5 bucket of rel_pos = key_index - query_index (int64 in, int64 out)."""
    nb = num_buckets // 2
    out = (rel_pos > 0).long() * nb
    rp = rel_pos.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(rp < max_exact, rp, large)


def position_bias(table: torch.Tensor, lq: int, lk: int, num_buckets: int, max_dist: int) -> torch.Tensor:
    """table [num_buckets, heads] -> bias [heads, lq, lk]."""
    rel = torch.arange(lk)[None, :] - torch.arange(lq)[:, None]
    return table[relative_position_bucket(rel, num_buckets, max_dist)].permute(2, 0, 1)


def t5_layer_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    return w * (x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps))


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def make_weights(cfg: T5Config, seed: int = 4321, std: float = 0.05) -> Dict[str, torch.Tensor]:
    """Random weights under the published Wan2.1 T5Encoder key names, bf16-representable (the reference loads a
    bf16 checkpoint)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).bfloat16().float()

    sd = {"token_embedding.weight": rnd(cfg.vocab_size, cfg.dim, s=1.0)}
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = (1.0 + rnd(cfg.dim, s=0.1))
        # T5 initialisation: q ~ (dim * head_dim)^-0.5 so that the un-scaled logits q.k are O(1); k, v ~ dim^-0.5
        sd[p + "attn.q.weight"] = rnd(cfg.dim_attn, cfg.dim, s=(cfg.dim * cfg.head_dim) ** -0.5)
        for n in "kv":
            sd[p + f"attn.{n}.weight"] = rnd(cfg.dim_attn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "attn.o.weight"] = rnd(cfg.dim, cfg.dim_attn, s=cfg.dim_attn ** -0.5)
        sd[p + "norm2.weight"] = (1.0 + rnd(cfg.dim, s=0.1))
        sd[p + "ffn.gate.0.weight"] = rnd(cfg.dim_ffn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "ffn.fc1.weight"] = rnd(cfg.dim_ffn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "ffn.fc2.weight"] = rnd(cfg.dim, cfg.dim_ffn, s=cfg.dim_ffn ** -0.5)
        sd[p + "pos_embedding.embedding.weight"] = rnd(cfg.num_buckets, cfg.num_heads, s=0.5)
    sd["norm.weight"] = (1.0 + rnd(cfg.dim, s=0.1))
    return sd


def encode(ids: torch.Tensor, mask: Optional[torch.Tensor], sd: Dict[str, torch.Tensor], cfg: T5Config) -> torch.Tensor:
    """ids int64 [L], mask [L] (1 = token, 0 = padding) -> hidden states fp32 [L, dim] (T5Encoder.forward)."""
    L = ids.shape[0]
    H, dk = cfg.num_heads, cfg.head_dim
    x = sd["token_embedding.weight"][ids].float()
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        e = position_bias(sd[p + "pos_embedding.embedding.weight"].float(), L, L, cfg.num_buckets, cfg.max_dist)
        h = t5_layer_norm(x, sd[p + "norm1.weight"].float(), cfg.eps)
        q = (h @ sd[p + "attn.q.weight"].float().T).view(L, H, dk)
        k = (h @ sd[p + "attn.k.weight"].float().T).view(L, H, dk)
        v = (h @ sd[p + "attn.v.weight"].float().T).view(L, H, dk)
        s = torch.einsum("inc,jnc->nij", q, k)          # no 1/sqrt(d): T5 folds it into the initialisation
        bias = e.clone()
        if mask is not None:
            bias.masked_fill_(mask.view(1, 1, L) == 0, torch.finfo(torch.float32).min)
        a = torch.softmax(s + bias, dim=-1)
        o = torch.einsum("nij,jnc->inc", a, v).reshape(L, H * dk)
        x = x + o @ sd[p + "attn.o.weight"].float().T
        h = t5_layer_norm(x, sd[p + "norm2.weight"].float(), cfg.eps)
        u = (h @ sd[p + "ffn.fc1.weight"].float().T) * gelu_tanh(h @ sd[p + "ffn.gate.0.weight"].float().T)
        x = x + u @ sd[p + "ffn.fc2.weight"].float().T
    return t5_layer_norm(x, sd["norm.weight"].float(), cfg.eps)


def encode_prompt_ids(ids: torch.Tensor, mask: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: T5Config) -> torch.Tensor:
    """WanPrompter.encode_prompt after tokenisation: encoder, then rows at and past the prompt length set to 0."""
    out = encode(ids, mask, sd, cfg)
    n = int((mask > 0).sum())
    out[n:] = 0
    return out


# ---- key-name bridge to the HF implementation used for pinning (and for loading HF-format checkpoints) -----------
def hf_to_wan_keys(hf_sd: Dict[str, torch.Tensor], num_layers: int) -> Dict[str, torch.Tensor]:
    out = {"token_embedding.weight": hf_sd["shared.weight"] if "shared.weight" in hf_sd else hf_sd["encoder.embed_tokens.weight"],
           "norm.weight": hf_sd["encoder.final_layer_norm.weight"]}
    for i in range(num_layers):
        s, d = f"encoder.block.{i}.layer.", f"blocks.{i}."
        out[d + "norm1.weight"] = hf_sd[s + "0.layer_norm.weight"]
        for n in "qkvo":
            out[d + f"attn.{n}.weight"] = hf_sd[s + f"0.SelfAttention.{n}.weight"]
        out[d + "pos_embedding.embedding.weight"] = hf_sd[s + "0.SelfAttention.relative_attention_bias.weight"]
        out[d + "norm2.weight"] = hf_sd[s + "1.layer_norm.weight"]
        out[d + "ffn.gate.0.weight"] = hf_sd[s + "1.DenseReluDense.wi_0.weight"]
        out[d + "ffn.fc1.weight"] = hf_sd[s + "1.DenseReluDense.wi_1.weight"]
        out[d + "ffn.fc2.weight"] = hf_sd[s + "1.DenseReluDense.wo.weight"]
    return out
