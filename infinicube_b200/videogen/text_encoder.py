"""umT5-XXL prompt encoder of the Wan2.1 pipeline on the sm_100a kernels.

Mirrors what the reference reaches through diffsynth when `WanVideoGenerator.generate` forwards `prompt` /
`negative_prompt` to the pipeline (infinicube/videogen/inference.py:216-226; checkpoint
`models_t5_umt5-xxl-enc-bf16.pth` named at :63-81): `WanTextEncoder` (the published Wan2.1 `T5Encoder`, umT5 layout:
per-block relative-position table, gated tanh-GELU FFN, RMS layer norm, un-scaled attention) and `WanPrompter`
(clean -> tokenise to 512 -> encode -> zero the rows past the prompt length).

Every nn.Linear runs on the tcgen05 GEMM (`ic_gemm_bf16`, residual adds fused into its epilogue); the embedding
gather, the norms, the biased/masked attention and the gate product are the `ic_t5_*` kernels.  Weights stay
resident in HBM (XXL: 11.4 GB of 180 GB).  There is no CPU path.
"""
from __future__ import annotations

import html
import math
import re
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from .. import ops
from .._lib import ICError, require_device


@dataclass
class T5Config:
    """`umt5_xxl(encoder_only=True)` of Wan2.1."""
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


def relative_position_bucket(rel_pos: torch.Tensor, num_buckets: int, max_dist: int) -> torch.Tensor:
    """Bidirectional T5 bucket of (key index - query index); CPU int64, fp32 log exactly as T5RelativeEmbedding."""
    nb = num_buckets // 2
    out = (rel_pos > 0).long() * nb
    rp = rel_pos.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(rp < max_exact, rp, large)


def bias_by_offset(table: torch.Tensor, L: int, num_buckets: int, max_dist: int) -> torch.Tensor:
    """Relative-position table [num_buckets, heads] -> [heads, 2L-1], entry (h, j - i + L - 1): the bias only
    depends on the offset, so the [heads, L, L] tensor of the reference is never materialised."""
    off = torch.arange(-(L - 1), L, dtype=torch.long)
    return table.float().cpu()[relative_position_bucket(off, num_buckets, max_dist)].t().contiguous()


_HF_PREFIX = "encoder.block."


def _from_hf_keys(sd: Dict[str, torch.Tensor], num_layers: int) -> Dict[str, torch.Tensor]:
    """transformers' UMT5EncoderModel key names -> the Wan2.1 T5Encoder names."""
    out = {"token_embedding.weight": sd["shared.weight"] if "shared.weight" in sd else sd["encoder.embed_tokens.weight"],
           "norm.weight": sd["encoder.final_layer_norm.weight"]}
    for i in range(num_layers):
        s, d = f"{_HF_PREFIX}{i}.layer.", f"blocks.{i}."
        out[d + "norm1.weight"] = sd[s + "0.layer_norm.weight"]
        for n in "qkvo":
            out[d + f"attn.{n}.weight"] = sd[s + f"0.SelfAttention.{n}.weight"]
        out[d + "pos_embedding.embedding.weight"] = sd[s + "0.SelfAttention.relative_attention_bias.weight"]
        out[d + "norm2.weight"] = sd[s + "1.layer_norm.weight"]
        out[d + "ffn.gate.0.weight"] = sd[s + "1.DenseReluDense.wi_0.weight"]
        out[d + "ffn.fc1.weight"] = sd[s + "1.DenseReluDense.wi_1.weight"]
        out[d + "ffn.fc2.weight"] = sd[s + "1.DenseReluDense.wo.weight"]
    return out


class WanTextEncoder:
    """Device-resident umT5 encoder.  `forward(ids, mask)` -> bf16 hidden states (same rank as `ids`)."""

    def __init__(self, cfg: Optional[T5Config] = None, device="cuda:0"):
        require_device()
        self.cfg = cfg or T5Config()
        if self.cfg.head_dim != 64:
            raise ICError(f"ic_t5_attention is built for head_dim 64 (umT5), got {self.cfg.head_dim}")
        for n in (self.cfg.dim, self.cfg.dim_attn, self.cfg.dim_ffn):
            if n % 8:
                raise ICError("umT5 widths must be multiples of 8")
        self.device = torch.device(device)
        self.layers = []
        self.token_embedding: Optional[torch.Tensor] = None
        self.norm: Optional[torch.Tensor] = None
        self._pos_tables = []          # per block [num_buckets, heads] fp32 on the host
        self._bias_cache: Dict[int, list] = {}
        self._ws: Dict[int, dict] = {}
        self.launch_count = 0

    # ---- weights ------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        c = self.cfg
        if any(k.startswith(_HF_PREFIX) for k in sd):
            sd = _from_hf_keys(sd, c.num_layers)
        want = {"token_embedding.weight", "norm.weight"}
        for i in range(c.num_layers):
            want |= {f"blocks.{i}.{n}" for n in ("norm1.weight", "attn.q.weight", "attn.k.weight", "attn.v.weight",
                                                  "attn.o.weight", "norm2.weight", "ffn.gate.0.weight", "ffn.fc1.weight",
                                                  "ffn.fc2.weight", "pos_embedding.embedding.weight")}
        missing = sorted(want - sd.keys())
        unexpected = sorted(sd.keys() - want)
        if missing or (strict and unexpected):
            raise KeyError(f"umT5 state dict: missing {missing[:6]}, unexpected {unexpected[:6]}")

        def w16(name, shape):
            t = sd[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected {tuple(shape)}, got {tuple(t.shape)}")
            return t.to(self.device, torch.bfloat16).contiguous()

        def w32(name, shape):
            t = sd[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected {tuple(shape)}, got {tuple(t.shape)}")
            return t.to(self.device, torch.float32).contiguous()

        self.token_embedding = w16("token_embedding.weight", (c.vocab_size, c.dim))
        self.norm = w32("norm.weight", (c.dim,))
        self.layers, self._pos_tables = [], []
        for i in range(c.num_layers):
            p = f"blocks.{i}."
            qkv = torch.cat([w16(p + f"attn.{n}.weight", (c.dim_attn, c.dim)) for n in "qkv"], dim=0).contiguous()
            self.layers.append(dict(
                norm1=w32(p + "norm1.weight", (c.dim,)), wqkv=qkv, wo=w16(p + "attn.o.weight", (c.dim, c.dim_attn)),
                norm2=w32(p + "norm2.weight", (c.dim,)), w_gate=w16(p + "ffn.gate.0.weight", (c.dim_ffn, c.dim)),
                w_fc1=w16(p + "ffn.fc1.weight", (c.dim_ffn, c.dim)), w_fc2=w16(p + "ffn.fc2.weight", (c.dim, c.dim_ffn))))
            t = sd[p + "pos_embedding.embedding.weight"]
            if tuple(t.shape) != (c.num_buckets, c.num_heads):
                raise ValueError(f"{p}pos_embedding.embedding.weight: expected {(c.num_buckets, c.num_heads)}")
            self._pos_tables.append(t.detach().float().cpu())
        self._bias_cache.clear()
        return [] if strict else unexpected

    def _bias(self, L: int):
        if L not in self._bias_cache:
            c = self.cfg
            self._bias_cache[L] = [bias_by_offset(t, L, c.num_buckets, c.max_dist).to(self.device) for t in self._pos_tables]
        return self._bias_cache[L]

    def _workspace(self, L: int) -> dict:
        if L not in self._ws:
            c, dev = self.cfg, self.device
            bf = dict(device=dev, dtype=torch.bfloat16)
            self._ws = {L: dict(x=torch.empty(L, c.dim, device=dev, dtype=torch.float32), h=torch.empty(L, c.dim, **bf),
                                qkv=torch.empty(L, 3 * c.dim_attn, **bf), att=torch.empty(L, c.dim_attn, **bf),
                                g=torch.empty(L, c.dim_ffn, **bf), u=torch.empty(L, c.dim_ffn, **bf))}
        return self._ws[L]

    # ---- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, ids: torch.Tensor, mask: Optional[torch.Tensor] = None, zero_from_row: int = -1) -> torch.Tensor:
        """ids [L] or [1, L] integer tokens, mask same shape (1 = token, 0 = right padding).  Returns bf16
        [L, dim] (or [1, L, dim]); rows >= zero_from_row (if >= 0) are zeros."""
        if self.token_embedding is None:
            raise ICError("WanTextEncoder has no weights: call load_state_dict first")
        batched = ids.dim() == 2
        if batched:
            if ids.shape[0] != 1:
                raise ValueError("one prompt per call (the reference encodes prompt and negative prompt separately)")
            ids = ids[0]
            mask = None if mask is None else mask[0]
        if ids.dim() != 1 or ids.numel() == 0:
            raise ValueError(f"ids must be a non-empty [L] or [1, L] tensor, got {tuple(ids.shape)}")
        if ids.dtype not in (torch.int32, torch.int64):
            raise TypeError(f"ids must be int32 / int64, got {ids.dtype}")
        c, L = self.cfg, ids.numel()
        lo, hi = int(ids.min()), int(ids.max())
        if lo < 0 or hi >= c.vocab_size:
            raise ValueError(f"token ids outside [0, {c.vocab_size}): min {lo}, max {hi}")
        if mask is not None and mask.shape != ids.shape:
            raise ValueError("mask must have the shape of ids")
        with torch.cuda.device(self.device):
            ids_d = ids.to(self.device, torch.int32).contiguous()
            mask_d = None if mask is None else (mask.to(self.device) != 0).to(torch.uint8).contiguous()
            ws, bias = self._workspace(L), self._bias(L)
            x, h, qkv, att, g, u = ws["x"], ws["h"], ws["qkv"], ws["att"], ws["g"], ws["u"]
            A = c.dim_attn
            ops.t5_embed(ids_d, self.token_embedding, x)
            n = 1
            for lw, b in zip(self.layers, bias):
                ops.t5_rmsnorm(x, lw["norm1"], h, c.eps)
                ops.gemm(h, lw["wqkv"], out_bf16=qkv)
                ops.t5_attention(qkv[:, :A], qkv[:, A:2 * A], qkv[:, 2 * A:], b, mask_d, att, c.num_heads)
                ops.gemm(att, lw["wo"], resid=x)                        # x += attn(norm1(x))
                ops.t5_rmsnorm(x, lw["norm2"], h, c.eps)
                ops.gemm(h, lw["w_gate"], act=1, out_bf16=g)            # gelu_tanh(gate(x)) in the GEMM epilogue
                ops.gemm(h, lw["w_fc1"], out_bf16=u)
                ops.mul_bf16(u, g, g)
                ops.gemm(g, lw["w_fc2"], resid=x)                       # x += fc2(fc1(x) * gelu(gate(x)))
                n += 9
            out = torch.empty(L, c.dim, device=self.device, dtype=torch.bfloat16)
            ops.t5_rmsnorm(x, self.norm, out, c.eps, zero_from_row)
            self.launch_count = n + 1
        return out[None] if batched else out

    __call__ = forward

    @property
    def flops_per_prompt(self) -> int:
        c, L = self.cfg, self.cfg.text_len
        per = 2 * L * c.dim * 4 * c.dim_attn + 4 * L * L * c.dim_attn + 6 * L * c.dim * c.dim_ffn
        return c.num_layers * per


# ------------------------------------------------------------------------------------------------------
# prompter
# ------------------------------------------------------------------------------------------------------
def whitespace_clean(text: str) -> str:
    return re.sub(r"\s+", " ", text).strip()


def basic_clean(text: str) -> str:
    """Wan's basic_clean: ftfy.fix_text (when ftfy is installed) and a double html.unescape."""
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    return html.unescape(html.unescape(text)).strip()


class WanPrompter:
    """diffsynth `WanPrompter`: tokenise (umT5 sentencepiece, padded / truncated to `text_len`, EOS appended) ->
    encoder -> rows past the prompt length set to 0 -> context [text_len, dim] for the DiT's text_embedding."""

    def __init__(self, tokenizer_path: Optional[str] = None, text_len: int = 512):
        self.text_len = text_len
        self.text_encoder: Optional[WanTextEncoder] = None
        self.tokenizer = None
        if tokenizer_path is not None:
            self.fetch_tokenizer(tokenizer_path)

    def fetch_models(self, text_encoder: Optional[WanTextEncoder] = None):
        self.text_encoder = text_encoder

    def fetch_tokenizer(self, tokenizer_path: str):
        """`tokenizer_path`: the `google/umt5-xxl` directory shipped with the Wan2.1 checkpoints."""
        from transformers import AutoTokenizer
        self.tokenizer = AutoTokenizer.from_pretrained(tokenizer_path)

    def process_prompt(self, prompt: str) -> str:
        return whitespace_clean(basic_clean(prompt))

    def tokenize(self, prompt: str) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.tokenizer is None:
            raise ICError("WanPrompter has no tokenizer: call fetch_tokenizer(<google/umt5-xxl directory>)")
        enc = self.tokenizer([self.process_prompt(prompt)], return_tensors="pt", padding="max_length", truncation=True,
                             max_length=self.text_len, add_special_tokens=True)
        return enc.input_ids[0], enc.attention_mask[0]

    @torch.no_grad()
    def encode_ids(self, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """Already tokenised prompt -> context bf16 [L, dim] with the padding rows zeroed."""
        if self.text_encoder is None:
            raise ICError("WanPrompter has no text encoder: call fetch_models(WanTextEncoder)")
        n_valid = int((mask.reshape(-1) > 0).sum())
        return self.text_encoder(ids.reshape(-1), mask.reshape(-1), zero_from_row=n_valid)

    def encode_prompt(self, prompt: str, positive: bool = True, device=None) -> torch.Tensor:
        ids, mask = self.tokenize(prompt)
        return self.encode_ids(ids, mask)

    __call__ = encode_prompt


def synthetic_t5_state_dict(cfg: T5Config, seed: int = 4321, std: float = 0.05) -> Dict[str, torch.Tensor]:
    """Random-init encoder weights under the Wan2.1 key names (benchmarks / tests without the 11 GB checkpoint).
    Generated on the CPU generator so every rank and the oracle draw the same values."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).bfloat16()

    sd = {"token_embedding.weight": rnd(cfg.vocab_size, cfg.dim, s=1.0)}
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + rnd(cfg.dim, s=0.1)
        sd[p + "attn.q.weight"] = rnd(cfg.dim_attn, cfg.dim, s=(cfg.dim * cfg.head_dim) ** -0.5)  # T5 init: logits O(1)
        for n in "kv":
            sd[p + f"attn.{n}.weight"] = rnd(cfg.dim_attn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "attn.o.weight"] = rnd(cfg.dim, cfg.dim_attn, s=cfg.dim_attn ** -0.5)
        sd[p + "norm2.weight"] = 1.0 + rnd(cfg.dim, s=0.1)
        sd[p + "ffn.gate.0.weight"] = rnd(cfg.dim_ffn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "ffn.fc1.weight"] = rnd(cfg.dim_ffn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "ffn.fc2.weight"] = rnd(cfg.dim, cfg.dim_ffn, s=cfg.dim_ffn ** -0.5)
        sd[p + "pos_embedding.embedding.weight"] = rnd(cfg.num_buckets, cfg.num_heads, s=0.5)
    sd["norm.weight"] = 1.0 + rnd(cfg.dim, s=0.1)
    return sd
