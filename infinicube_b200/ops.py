"""Thin torch-tensor front end over the C ABI for the granular ops.

torch is plumbing here (device memory + streams); every op below runs a hand-written sm_100a kernel
from libinfinicube_b200.so and raises if the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmEpilogue, check, lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.ICError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name} must be contiguous in its last dimension")


def gemm_block_n(n: int) -> int:
    return lib().ic_gemm_block_n(n)


def gemm(a: torch.Tensor, b: torch.Tensor, *, bias=None, bias_per_row=False, act=0, out_bf16=None, rowss=None,
         out_f32=None, addend=None, resid=None, gate=None) -> None:
    """C = a[M,K] @ b[N,K]^T with the fused epilogue (see include/infinicube_b200.h: ic_gemm_bf16)."""
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    M, K = a.shape
    N, K2 = b.shape
    if K != K2:
        raise ValueError("inner dimensions differ")
    ep = GemmEpilogue()
    ep.bias = _ptr(bias)
    ep.bias_per_row = int(bias_per_row)
    ep.act = int(act)
    if out_bf16 is not None:
        _req(out_bf16, torch.bfloat16, "out_bf16")
        ep.out_bf16, ep.ld_out = _ptr(out_bf16), out_bf16.stride(0)
    if rowss is not None:
        _req(rowss, torch.float32, "rowss")
        ep.rowss, ep.rowss_ld = _ptr(rowss), rowss.stride(0)
    if out_f32 is not None:
        _req(out_f32, torch.float32, "out_f32")
        ep.out_f32, ep.ld_f32 = _ptr(out_f32), out_f32.stride(0)
    if addend is not None:
        _req(addend, torch.float32, "addend")
        ep.addend, ep.ld_add = _ptr(addend), addend.stride(0)
    if resid is not None:
        _req(resid, torch.float32, "resid")
        ep.resid, ep.ld_res = _ptr(resid), resid.stride(0)
    ep.gate = _ptr(gate)
    check(lib().ic_gemm_bf16(_ptr(a), a.stride(0), _ptr(b), b.stride(0), M, N, K, C.byref(ep), _stream()), "ic_gemm_bf16")


def fmha(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, out: torch.Tensor, n_heads: int, softmax_scale: float,
         seg_len: Optional[int] = None, n_seg: int = 1, k_seg_stride: int = 0, vt_seg_stride: int = 0) -> None:
    """out[Sq, H*128] = attention(q[Sq, H*128], k[S, H*128], vt[H*128, S]) per head (ic_fmha_fwd)."""
    for t, n in ((q, "q"), (k, "k"), (vt, "vt"), (out, "out")):
        _req(t, torch.bfloat16, n)
    Sq = q.shape[0]
    if seg_len is None:
        seg_len = k.shape[0]
    ldk = k.stride(-2)
    ldvt = vt.stride(-2)
    check(lib().ic_fmha_fwd(_ptr(q), q.stride(0), _ptr(k), ldk, k_seg_stride, _ptr(vt), ldvt, vt_seg_stride, _ptr(out),
                            out.stride(0), Sq, seg_len, n_seg, n_heads, float(softmax_scale), _stream()), "ic_fmha_fwd")


def ln_modulate(x: torch.Tensor, mul: torch.Tensor, add: torch.Tensor, out: torch.Tensor, mul_plus_one: bool,
                eps: float = 1e-6) -> None:
    _req(x, torch.float32, "x")
    _req(out, torch.bfloat16, "out")
    rows, D = x.shape
    check(lib().ic_ln_modulate(_ptr(x), x.stride(0), _ptr(mul), _ptr(add), int(mul_plus_one), _ptr(out), out.stride(0),
                               rows, D, eps, _stream()), "ic_ln_modulate")


def rmsnorm_rope(src: torch.Tensor, rowss: torch.Tensor, ss_off: int, ss_cnt: int, weight: torch.Tensor,
                 dst: torch.Tensor, eps: float = 1e-6, rope=None, frame0: int = 0) -> None:
    """rope = (tab_f, tab_h, tab_w) fp32 (cos,sin) tables or None."""
    _req(src, torch.bfloat16, "src")
    _req(dst, torch.bfloat16, "dst")
    rows, D = src.shape
    tf = th = tw = None
    nf = nh = nw = 0
    if rope is not None:
        tf, th, tw = rope
        nf, nh, nw = tf.shape[0], th.shape[0], tw.shape[0]
    check(lib().ic_rmsnorm_rope(_ptr(src), src.stride(0), _ptr(rowss), rowss.stride(0), ss_off, ss_cnt, _ptr(weight),
                                _ptr(dst), dst.stride(0), rows, D, eps, _ptr(tf), _ptr(th), _ptr(tw), nf, nh, nw,
                                frame0, _stream()), "ic_rmsnorm_rope")


def patchify(latents: torch.Tensor, out: torch.Tensor, col_off: int = 0) -> None:
    _req(latents, torch.float32, "latents")
    _req(out, torch.bfloat16, "out")
    Cc, F, H, W = latents.shape
    check(lib().ic_patchify(_ptr(latents), _ptr(out), Cc, F, H, W, out.stride(0), col_off, _stream()), "ic_patchify")


def unpatchify_cfg_step(latents: Optional[torch.Tensor], head_pos: torch.Tensor, head_neg: Optional[torch.Tensor],
                        shape, cfg_scale: float, dsigma: float, v_out: Optional[torch.Tensor] = None) -> None:
    Cc, F, H, W = shape
    check(lib().ic_unpatchify_cfg_step(_ptr(latents), _ptr(head_pos), _ptr(head_neg), Cc, F, H, W, float(cfg_scale),
                                       float(dsigma), _ptr(v_out), _stream()), "ic_unpatchify_cfg_step")


# ---- umT5 prompt encoder pieces ---------------------------------------------------
def t5_embed(ids: torch.Tensor, table: torch.Tensor, x: torch.Tensor) -> None:
    """x[r] = float(table[ids[r]]) (ic_t5_embed); ids int32 [L], table bf16 [V, D], x fp32 [L, D]."""
    _req(ids, torch.int32, "ids")
    _req(table, torch.bfloat16, "table")
    _req(x, torch.float32, "x")
    check(lib().ic_t5_embed(_ptr(ids), ids.shape[0], _ptr(table), table.shape[0], table.shape[1], _ptr(x), x.stride(0),
                            _stream()), "ic_t5_embed")


def t5_rmsnorm(x: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, eps: float, zero_from_row: int = -1) -> None:
    _req(x, torch.float32, "x")
    _req(weight, torch.float32, "weight")
    _req(out, torch.bfloat16, "out")
    check(lib().ic_t5_rmsnorm(_ptr(x), x.stride(0), _ptr(weight), _ptr(out), out.stride(0), x.shape[0], x.shape[1],
                              float(eps), int(zero_from_row), _stream()), "ic_t5_rmsnorm")


def t5_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, bias_by_offset: torch.Tensor,
                 key_mask: Optional[torch.Tensor], out: torch.Tensor, n_heads: int) -> None:
    """q / k / v: bf16 [L, *] views sharing one row pitch, head h = columns [64h, 64h+64) (ic_t5_attention)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.bfloat16, n)
    _req(bias_by_offset, torch.float32, "bias_by_offset")
    L = q.shape[0]
    if not (q.stride(0) == k.stride(0) == v.stride(0)):
        raise ValueError("q, k, v must share one row pitch")
    if tuple(bias_by_offset.shape) != (n_heads, 2 * L - 1) or not bias_by_offset.is_contiguous():
        raise ValueError(f"bias_by_offset must be a contiguous [{n_heads}, {2 * L - 1}] tensor")
    if key_mask is not None:
        _req(key_mask, torch.uint8, "key_mask")
        if key_mask.numel() != L:
            raise ValueError("key_mask must hold one byte per key")
    check(lib().ic_t5_attention(_ptr(q), _ptr(k), _ptr(v), q.stride(0), _ptr(bias_by_offset), _ptr(key_mask), _ptr(out),
                                out.stride(0), L, n_heads, _stream()), "ic_t5_attention")


def mul_bf16(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> None:
    for t, n in ((a, "a"), (b, "b"), (out, "out")):
        _req(t, torch.bfloat16, n)
        if not t.is_contiguous():
            raise ValueError(f"{n} must be contiguous")
    check(lib().ic_mul_bf16(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), "ic_mul_bf16")
