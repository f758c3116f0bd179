"""Wan2.1 3-D causal VAE on the B200 kernels (host orchestration).

Stands in for diffsynth's `WanVideoVAE` as used by the reference pipeline call
(infinicube/videogen/inference.py:216-226: two tiled buffer encodes + one tiled decode; architecture per
SURVEY.md Appendix A.9).  Activations are channels-last bf16 `[T, H, W, C]`; every convolution is the tcgen05
implicit-GEMM kernel (`ic_conv_cl`), everything else a kernel from csrc/vae_ops.cu or the tcgen05 GEMM.
The whole frame sequence is processed at once: causal padding comes from TMA zero fill, so no per-chunk feature
cache exists (tests prove the equivalence against the chunked reference formulation).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import ops
from .._lib import ICError, check, lib, require_device

LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
               -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
              1.1253, 2.8251, 1.9160]


def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def encoder_layers(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_down=(False, True, True)):
    dims = [dim * u for u in (1,) + tuple(dim_mult)]
    out = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            out.append(("res", cin, cout))
            cin = cout
        if i != len(dim_mult) - 1:
            out.append(("down3d" if temporal_down[i] else "down2d", cout))
    return out, dims[-1]


def decoder_layers(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_up=(True, True, False)):
    dims = [dim * u for u in (dim_mult[-1],) + tuple(dim_mult[::-1])]
    out = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(num_res_blocks + 1):
            out.append(("res", cin, cout))
            cin = cout
        if i != len(dim_mult) - 1:
            out.append(("up3d" if temporal_up[i] else "up2d", cout))
    return out, dims[0]


def tile_tasks(H: int, W: int, size: Tuple[int, int], stride: Tuple[int, int]):
    """DiffSynth's tile enumeration (WanVideoVAE.tiled_decode / tiled_encode)."""
    tasks = []
    for h in range(0, H, stride[0]):
        if h - stride[0] >= 0 and h - stride[0] + size[0] >= H:
            continue
        for w in range(0, W, stride[1]):
            if w - stride[1] >= 0 and w - stride[1] + size[1] >= W:
                continue
            tasks.append((h, min(h + size[0], H), w, min(w + size[1], W), h + size[0] >= H, w + size[1] >= W))
    return tasks


class _Conv:
    """One convolution: weight matrix [Cout_pad, ntaps*Cin_pad] bf16 + fp32 bias + tap list."""

    def __init__(self, w2: torch.Tensor, bias: torch.Tensor, taps: Sequence[Tuple[int, int, int]], cin: int, cout: int):
        self.w = w2.contiguous()
        self.b = bias.contiguous()
        self.taps = (C.c_int * (3 * len(taps)))(*[v for t in taps for v in t])
        self.ntaps = len(taps)
        self.cin, self.cout = cin, cout


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class WanVideoVAE:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0", dim: int = 96, z_dim: int = 16,
                 world_size: int = 1, rank: int = 0):
        require_device()
        self.device = torch.device(device)
        # multi-GPU: the spatial tiles of the tiled encode / decode are independent units -> round-robin over ranks,
        # blended accumulators summed with one all-reduce
        self.world_size, self.rank = world_size, rank
        self.dim, self.z_dim = dim, z_dim
        self.sd = state_dict
        self.mean = torch.tensor(LATENT_MEAN, device=self.device, dtype=torch.float32)
        self.std = torch.tensor(LATENT_STD, device=self.device, dtype=torch.float32)
        self.convs: Dict[str, _Conv] = {}
        self.gammas: Dict[str, torch.Tensor] = {}
        self.attn_wv: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._prepare()

    # ---- weight preparation (load time) --------------------------------------------------------------------
    def _mat(self, w: torch.Tensor, cin_pad: int, cout_pad: int) -> torch.Tensor:
        """[Cout, Cin, *k] -> [Cout_pad, ntaps*Cin_pad] with K index = tap*Cin_pad + c (taps in kt,kh,kw order)."""
        co, ci = w.shape[:2]
        k = w.reshape(co, ci, -1).permute(0, 2, 1)  # [co, ntaps, ci]
        out = torch.zeros(cout_pad, k.shape[1], cin_pad, dtype=torch.float32)
        out[:co, :, :ci] = k.float()
        return out.reshape(cout_pad, -1).to(self.device, torch.bfloat16)

    def _bias(self, b: torch.Tensor, cout_pad: int) -> torch.Tensor:
        out = torch.zeros(cout_pad, dtype=torch.float32)
        out[: b.shape[0]] = b.float()
        return out.to(self.device)

    def _add_conv(self, name: str, mode: str = "auto"):
        w, b = self.sd[name + ".weight"].cpu(), self.sd[name + ".bias"].cpu()
        co, ci = w.shape[:2]
        cin_pad, cout_pad = _pad_to(ci, 32), _pad_to(co, 8)
        if cout_pad < 32 and name != "decoder.head.2":
            cout_pad = 32  # keeps the next layer's Cin a multiple of 32
        if mode == "down2d":  # stride-2 3x3 with pad (0,1,0,1) as a 2x2-cell conv over space-to-depth input
            w4 = torch.zeros(co, 4, 4, ci)  # [co, cell(cy*2+cx), phase(a*2+b), c]
            for cy in range(2):
                for cx in range(2):
                    for a in range(2):
                        for bb in range(2):
                            dy, dx = 2 * cy + a, 2 * cx + bb
                            if dy <= 2 and dx <= 2:
                                w4[:, cy * 2 + cx, a * 2 + bb, :] = w[:, :, dy, dx].float()
            w2 = w4.reshape(co, -1).to(self.device, torch.bfloat16)
            taps = [(0, cy, cx) for cy in range(2) for cx in range(2)]
            self.convs[name] = _Conv(w2, self._bias(b, co), taps, 4 * ci, co)
            return
        if mode == "time3_gather":  # encoder time_conv (3,1,1) stride 2 on channel-concatenated frame triples
            w2 = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(co, 3 * ci).to(self.device, torch.bfloat16)
            self.convs[name] = _Conv(w2, self._bias(b, co), [(0, 0, 0)], 3 * ci, co)
            return
        if w.dim() == 5:
            kt, kh, kw = w.shape[2:]
            taps = [(it - (kt - 1), ih - kh // 2, iw - kw // 2) for it in range(kt) for ih in range(kh) for iw in range(kw)]
        else:
            kh, kw = w.shape[2:]
            taps = [(0, ih - kh // 2, iw - kw // 2) for ih in range(kh) for iw in range(kw)]
        self.convs[name] = _Conv(self._mat(w, cin_pad, cout_pad), self._bias(b, cout_pad), taps, cin_pad, cout_pad)

    def _add_gamma(self, name: str):
        self.gammas[name] = self.sd[name].reshape(-1).to(self.device, torch.float32).contiguous()

    def _add_res(self, p: str):
        self._add_gamma(p + ".residual.0.gamma")
        self._add_conv(p + ".residual.2")
        self._add_gamma(p + ".residual.3.gamma")
        self._add_conv(p + ".residual.6")
        if p + ".shortcut.weight" in self.sd:
            self._add_conv(p + ".shortcut")

    def _add_attn(self, p: str):
        self._add_gamma(p + ".norm.gamma")
        w, b = self.sd[p + ".to_qkv.weight"].cpu(), self.sd[p + ".to_qkv.bias"].cpu()
        c = w.shape[1]
        wqk = w[: 2 * c].reshape(2 * c, c).to(self.device, torch.bfloat16).contiguous()
        self.convs[p + ".qk"] = _Conv(wqk, b[: 2 * c].float().to(self.device), [(0, 0, 0)], c, 2 * c)
        self.attn_wv[p] = (w[2 * c:].reshape(c, c).to(self.device, torch.bfloat16).contiguous(),
                           b[2 * c:].float().to(self.device).contiguous())
        self._add_conv(p + ".proj")

    def _prepare(self):
        enc, _ = encoder_layers(self.dim)
        self._add_conv("encoder.conv1")
        for i, l in enumerate(enc):
            p = f"encoder.downsamples.{i}"
            if l[0] == "res":
                self._add_res(p)
            else:
                self._add_conv(p + ".resample.1", "down2d")
                if l[0] == "down3d":
                    self._add_conv(p + ".time_conv", "time3_gather")
        self._add_res("encoder.middle.0")
        self._add_attn("encoder.middle.1")
        self._add_res("encoder.middle.2")
        self._add_gamma("encoder.head.0.gamma")
        self._add_conv("encoder.head.2")
        self._add_conv("conv1")
        self._add_conv("conv2")
        dec, _ = decoder_layers(self.dim)
        self._add_conv("decoder.conv1")
        self._add_res("decoder.middle.0")
        self._add_attn("decoder.middle.1")
        self._add_res("decoder.middle.2")
        for i, l in enumerate(dec):
            p = f"decoder.upsamples.{i}"
            if l[0] == "res":
                self._add_res(p)
            else:
                self._add_conv(p + ".resample.1")
                if l[0] == "up3d":
                    self._add_conv(p + ".time_conv")
        self._add_gamma("decoder.head.0.gamma")
        self._add_conv("decoder.head.2")
        self.sd = None  # originals no longer needed

    # ---- op wrappers ------------------------------------------------------------------------------------------
    def _conv(self, x: torch.Tensor, name: str, resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              t_out: Optional[int] = None) -> torch.Tensor:
        cv = self.convs[name]
        Tin, H, W, Cin = x.shape
        if Cin != cv.cin:
            raise ICError(f"{name}: input has {Cin} channels, weights expect {cv.cin}")
        T = Tin if t_out is None else t_out
        if out is None:
            out = torch.empty((T, H, W, cv.cout), dtype=torch.bfloat16, device=x.device)
        check(lib().ic_conv_cl(_p(x), Tin, H, W, Cin, _p(cv.w), _p(cv.b), cv.taps, cv.ntaps, _p(out), T, H, W, cv.cout,
                               out.stride(2), _p(resid), 0 if resid is None else resid.stride(2), _stream()),
              f"ic_conv_cl({name})")
        return out

    def _norm(self, x: torch.Tensor, gname: str, silu: bool = True) -> torch.Tensor:
        out = torch.empty_like(x)
        npix = x.numel() // x.shape[-1]
        check(lib().ic_rmsnorm_cl(_p(x), _p(self.gammas[gname]), _p(out), npix, x.shape[-1], int(silu), _stream()),
              "ic_rmsnorm_cl")
        return out

    def _res(self, x: torch.Tensor, p: str) -> torch.Tensor:
        y = self._conv(self._norm(x, p + ".residual.0.gamma"), p + ".residual.2")
        h = self._conv(x, p + ".shortcut") if (p + ".shortcut") in self.convs else x
        return self._conv(self._norm(y, p + ".residual.3.gamma"), p + ".residual.6", resid=h)

    def _attn(self, x: torch.Tensor, p: str) -> torch.Tensor:
        T, H, W, Cc = x.shape
        n = H * W
        if n % 8:
            raise ICError("attention needs h*w to be a multiple of 8")
        xn = self._norm(x, p + ".norm.gamma", silu=False)
        qk = self._conv(xn, p + ".qk").view(T, n, 2 * Cc)
        wv, bv = self.attn_wv[p]
        xn2 = xn.view(T, n, Cc)
        att = torch.empty((T, H, W, Cc), dtype=torch.bfloat16, device=x.device)
        s = torch.empty((n, n), dtype=torch.float32, device=x.device)
        pm = torch.empty((n, n), dtype=torch.bfloat16, device=x.device)
        vt = torch.empty((Cc, n), dtype=torch.bfloat16, device=x.device)
        for t in range(T):
            ops.gemm(qk[t, :, :Cc], qk[t, :, Cc:], out_f32=s)                      # S = Q K^T
            check(lib().ic_softmax_rows(_p(s), n, _p(pm), n, n, n, 1.0 / math.sqrt(Cc), _stream()), "ic_softmax_rows")
            ops.gemm(wv, xn2[t], bias=bv, bias_per_row=True, out_bf16=vt)          # V^T = W_v xn^T + b
            ops.gemm(pm, vt, out_bf16=att[t].view(n, Cc))                          # O = P V
        return self._conv(att, p + ".proj", resid=x)

    def _upsample(self, x: torch.Tensor, p: str, temporal: bool) -> torch.Tensor:
        T, H, W, Cc = x.shape
        if temporal and T > 1:
            y = self._conv(x[1:], p + ".time_conv")  # causal over frames >= 1 only; frame 0 passes through
            out = torch.empty((2 * T - 1, H, W, Cc), dtype=torch.bfloat16, device=x.device)
            check(lib().ic_time_interleave_cl(_p(x), _p(y), _p(out), T - 1, H * W, Cc, _stream()), "ic_time_interleave_cl")
            x, T = out, 2 * T - 1
        up = torch.empty((T, 2 * H, 2 * W, Cc), dtype=torch.bfloat16, device=x.device)
        check(lib().ic_upsample2x_cl(_p(x), _p(up), T, H, W, Cc, _stream()), "ic_upsample2x_cl")
        return self._conv(up, p + ".resample.1")

    def _downsample(self, x: torch.Tensor, p: str, temporal: bool) -> torch.Tensor:
        T, H, W, Cc = x.shape
        if H % 2 or W % 2:
            raise ICError("downsample needs even H and W")
        s2d = torch.empty((T, H // 2, W // 2, 4 * Cc), dtype=torch.bfloat16, device=x.device)
        check(lib().ic_space_to_depth_cl(_p(x), _p(s2d), T, H, W, Cc, _stream()), "ic_space_to_depth_cl")
        x = self._conv(s2d, p + ".resample.1")
        if temporal and T > 1:
            K = (T - 1) // 2
            g = torch.empty((K, H // 2, W // 2, 3 * Cc), dtype=torch.bfloat16, device=x.device)
            check(lib().ic_time_gather3_cl(_p(x), _p(g), K, (H // 2) * (W // 2), Cc, _stream()), "ic_time_gather3_cl")
            out = torch.empty((1 + K, H // 2, W // 2, Cc), dtype=torch.bfloat16, device=x.device)
            out[0].copy_(x[0])
            self._conv(g, p + ".time_conv", out=out[1:])
            x = out
        return x

    # ---- decode / encode one (tile of a) sequence ------------------------------------------------------------
    def decode_cl(self, z: torch.Tensor) -> torch.Tensor:
        """normalised latents fp32 [16, T, h, w] -> bf16 channels-last video [4T-3, 8h, 8w, 8] (3 used, unclamped)."""
        Cz, T, h, w = z.shape
        z = z.to(self.device, torch.float32).contiguous()
        x = torch.empty((T, h, w, 32), dtype=torch.bfloat16, device=self.device)
        check(lib().ic_latent_to_cl(_p(z), _p(self.mean), _p(self.std), _p(x), Cz, T * h * w, 32, _stream()), "ic_latent_to_cl")
        x = self._conv(x, "conv2")
        x = self._conv(x, "decoder.conv1")
        x = self._res(x, "decoder.middle.0")
        x = self._attn(x, "decoder.middle.1")
        x = self._res(x, "decoder.middle.2")
        layers, _ = decoder_layers(self.dim)
        for i, l in enumerate(layers):
            p = f"decoder.upsamples.{i}"
            x = self._res(x, p) if l[0] == "res" else self._upsample(x, p, l[0] == "up3d")
        return self._conv(self._norm(x, "decoder.head.0.gamma"), "decoder.head.2")

    def encode_cl(self, frames_cl: torch.Tensor) -> torch.Tensor:
        """bf16 channels-last [-1,1] video [T, H, W, 32] -> bf16 [T', H/8, W/8, 32] (first 16 = un-normalised mu)."""
        x = self._conv(frames_cl, "encoder.conv1")
        layers, _ = encoder_layers(self.dim)
        for i, l in enumerate(layers):
            p = f"encoder.downsamples.{i}"
            x = self._res(x, p) if l[0] == "res" else self._downsample(x, p, l[0] == "down3d")
        x = self._res(x, "encoder.middle.0")
        x = self._attn(x, "encoder.middle.1")
        x = self._res(x, "encoder.middle.2")
        x = self._conv(self._norm(x, "encoder.head.0.gamma"), "encoder.head.2")
        return self._conv(x, "conv1")

    # ---- public surface -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, z: torch.Tensor, tiled: bool = True, tile_size=(30, 52), tile_stride=(15, 26),
               want_frames: bool = True, want_f32: bool = False):
        """latents [16, T, h, w] -> (uint8 frames [4T-3, 8h, 8w, 3] on device, fp32 video [3, 4T-3, 8h, 8w] or None),
        both clamped to [-1, 1] like the reference."""
        _, T, h, w = z.shape
        To, H, W = 4 * T - 3, 8 * h, 8 * w
        values = torch.zeros((3, To, H, W), dtype=torch.float32, device=self.device)
        weight = torch.zeros((H, W), dtype=torch.float32, device=self.device)
        if tiled:
            tasks = tile_tasks(h, w, tile_size, tile_stride)
            border = ((tile_size[0] - tile_stride[0]) * 8, (tile_size[1] - tile_stride[1]) * 8)
        else:
            tasks = [(0, h, 0, w, True, True)]
            border = (1, 1)
        shard = self.world_size > 1 and len(tasks) > 1
        for ti, (h0, h1, w0, w1, bot, right) in enumerate(tasks):
            if shard and ti % self.world_size != self.rank:
                continue
            y = self.decode_cl(z[:, :, h0:h1, w0:w1])
            bm = (1 if h0 == 0 else 0) | (2 if bot else 0) | (4 if w0 == 0 else 0) | (8 if right else 0)
            check(lib().ic_blend_accumulate(_p(y), y.shape[-1], 3, To, y.shape[1], y.shape[2], _p(values), _p(weight), H, W,
                                            h0 * 8, w0 * 8, bm, border[0], border[1], _stream()), "ic_blend_accumulate")
        if shard:
            self._all_reduce(values, weight)
        frames = torch.empty((To, H, W, 3), dtype=torch.uint8, device=self.device) if want_frames else None
        vid = torch.empty((3, To, H, W), dtype=torch.float32, device=self.device) if want_f32 else None
        check(lib().ic_blend_finalize(_p(values), _p(weight), 3, To, H, W, 1, _p(vid), _p(frames), _stream()),
              "ic_blend_finalize")
        return frames, vid

    @torch.no_grad()
    def encode(self, frames_u8: torch.Tensor, tiled: bool = True, tile_size=(30, 52), tile_stride=(15, 26)) -> torch.Tensor:
        """uint8 frames [T, H, W, 3] (device) -> normalised latents fp32 [16, (T-1)/4+1, H/8, W/8]."""
        return self.encode_many([frames_u8], tiled, tile_size, tile_stride)[0]

    @torch.no_grad()
    def encode_many(self, videos, tiled: bool = True, tile_size=(30, 52), tile_stride=(15, 26)):
        """Several uint8 videos of one shape -> their latents.  With more than one rank the tiles of ALL videos are
        dealt round-robin together (the two guidance buffers of a call are 18 tiles: 3 rounds on 8 ranks instead of
        2 + 2) and the blended accumulators of all videos cross the ranks in one all-reduce."""
        T, H, W, _ = videos[0].shape
        for v in videos:
            if tuple(v.shape) != (T, H, W, 3):
                raise ValueError(f"encode_many needs equally shaped [T, H, W, 3] videos, got {tuple(v.shape)}")
        if T % 4 != 1 or H % 8 or W % 8:
            raise ValueError(f"encode needs T % 4 == 1 and H, W multiples of 8, got {tuple(videos[0].shape)}")
        nv = len(videos)
        Tl, h, w = (T - 1) // 4 + 1, H // 8, W // 8
        values = torch.zeros((nv, 16, Tl, h, w), dtype=torch.float32, device=self.device)
        weight = torch.zeros((nv, h, w), dtype=torch.float32, device=self.device)
        if tiled:
            size = (tile_size[0] * 8, tile_size[1] * 8)
            stride = (tile_stride[0] * 8, tile_stride[1] * 8)
            tasks = tile_tasks(H, W, size, stride)
            border = ((size[0] - stride[0]) // 8, (size[1] - stride[1]) // 8)
        else:
            tasks = [(0, H, 0, W, True, True)]
            border = (1, 1)
        jobs = [(vi, task) for vi in range(nv) for task in tasks]
        shard = self.world_size > 1 and len(jobs) > 1
        x, x_of = None, -1
        for ji, (vi, (h0, h1, w0, w1, bot, right)) in enumerate(jobs):
            if shard and ji % self.world_size != self.rank:
                continue
            if x_of != vi:   # channels-last bf16 copy of this video, made only where one of its tiles runs
                fr = videos[vi].to(self.device).contiguous()
                x = torch.empty((T, H, W, 32), dtype=torch.bfloat16, device=self.device)
                check(lib().ic_frames_to_cl(_p(fr), _p(x), T * H * W, 32, _stream()), "ic_frames_to_cl")
                x_of = vi
            tile = x[:, h0:h1, w0:w1].contiguous() if (h1 - h0, w1 - w0) != (H, W) else x
            y = self.encode_cl(tile)
            bm = (1 if h0 == 0 else 0) | (2 if bot else 0) | (4 if w0 == 0 else 0) | (8 if right else 0)
            check(lib().ic_blend_accumulate(_p(y), y.shape[-1], 16, Tl, y.shape[1], y.shape[2], _p(values[vi]), _p(weight[vi]),
                                            h, w, h0 // 8, w0 // 8, bm, border[0], border[1], _stream()), "ic_blend_accumulate")
        if shard:
            self._all_reduce(values, weight)
        out = []
        for vi in range(nv):
            mu = torch.empty((16, Tl, h, w), dtype=torch.float32, device=self.device)
            check(lib().ic_blend_finalize(_p(values[vi]), _p(weight[vi]), 16, Tl, h, w, 0, _p(mu), None, _stream()),
                  "ic_blend_finalize")
            out.append((mu - self.mean.view(-1, 1, 1, 1)) / self.std.view(-1, 1, 1, 1))
        return out

    @staticmethod
    def _all_reduce(values: torch.Tensor, weight: torch.Tensor):
        import torch.distributed as dist
        dist.all_reduce(values)
        dist.all_reduce(weight)

    # names used by WanVideoPipeline
    def encode_frames(self, video, tiled: bool = True) -> torch.Tensor:
        if isinstance(video, np.ndarray):
            video = torch.from_numpy(video)
        elif isinstance(video, (list, tuple)):  # list of PIL images, like the reference passes
            video = torch.from_numpy(np.stack([np.asarray(f) for f in video]))
        return self.encode(video, tiled=tiled)

    def encode_frames_many(self, videos, tiled: bool = True):
        return self.encode_many([self._as_u8_tensor(v) for v in videos], tiled=tiled)

    @staticmethod
    def _as_u8_tensor(video) -> torch.Tensor:
        if isinstance(video, np.ndarray):
            return torch.from_numpy(video)
        if isinstance(video, (list, tuple)):
            return torch.from_numpy(np.stack([np.asarray(f) for f in video]))
        return video

    def decode_to_frames(self, latents: torch.Tensor, tiled: bool = True, output_type: str = "pil"):
        frames, _ = self.decode(latents, tiled=tiled)
        if output_type == "tensor":
            return frames
        arr = frames.cpu().numpy()
        if output_type == "np":
            return arr
        from PIL import Image
        return [Image.fromarray(arr[i], mode="RGB") for i in range(arr.shape[0])]


def synthetic_vae_state_dict(seed: int = 4321, dim: int = 96, z_dim: int = 16) -> Dict[str, torch.Tensor]:
    """Random-init weights of the Wan2.1 VAE architecture under the official module names (there is no
    Wan2.1_VAE.pth offline); scaled so activations stay O(1) through the stack."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k, gain=1.0):
        fan = cin * math.prod(k)
        sd[name + ".weight"] = (torch.randn(cout, cin, *k, generator=g) * (gain / math.sqrt(fan))).to(torch.bfloat16)
        sd[name + ".bias"] = (torch.randn(cout, generator=g) * 0.02).to(torch.bfloat16)

    def res(p, cin, cout):
        sd[p + ".residual.0.gamma"] = (1.0 + 0.1 * torch.randn(cin, 1, 1, 1, generator=g)).to(torch.bfloat16)
        conv(p + ".residual.2", cout, cin, (3, 3, 3), 1.4)
        sd[p + ".residual.3.gamma"] = (1.0 + 0.1 * torch.randn(cout, 1, 1, 1, generator=g)).to(torch.bfloat16)
        conv(p + ".residual.6", cout, cout, (3, 3, 3), 0.7)
        if cin != cout:
            conv(p + ".shortcut", cout, cin, (1, 1, 1))

    def attn(p, c):
        sd[p + ".norm.gamma"] = (1.0 + 0.1 * torch.randn(c, 1, 1, generator=g)).to(torch.bfloat16)
        conv(p + ".to_qkv", 3 * c, c, (1, 1))
        conv(p + ".proj", c, c, (1, 1), 0.5)

    enc, top = encoder_layers(dim)
    conv("encoder.conv1", dim, 3, (3, 3, 3))
    for i, l in enumerate(enc):
        p = f"encoder.downsamples.{i}"
        if l[0] == "res":
            res(p, l[1], l[2])
        else:
            conv(p + ".resample.1", l[1], l[1], (3, 3))
            if l[0] == "down3d":
                conv(p + ".time_conv", l[1], l[1], (3, 1, 1))
    res("encoder.middle.0", top, top)
    attn("encoder.middle.1", top)
    res("encoder.middle.2", top, top)
    sd["encoder.head.0.gamma"] = (1.0 + 0.1 * torch.randn(top, 1, 1, 1, generator=g)).to(torch.bfloat16)
    conv("encoder.head.2", 2 * z_dim, top, (3, 3, 3))
    conv("conv1", 2 * z_dim, 2 * z_dim, (1, 1, 1))
    conv("conv2", z_dim, z_dim, (1, 1, 1))
    dec, dtop = decoder_layers(dim)
    conv("decoder.conv1", dtop, z_dim, (3, 3, 3))
    res("decoder.middle.0", dtop, dtop)
    attn("decoder.middle.1", dtop)
    res("decoder.middle.2", dtop, dtop)
    for i, l in enumerate(dec):
        p = f"decoder.upsamples.{i}"
        if l[0] == "res":
            res(p, l[1], l[2])
        else:
            conv(p + ".resample.1", l[1] // 2, l[1], (3, 3))
            if l[0] == "up3d":
                conv(p + ".time_conv", 2 * l[1], l[1], (3, 1, 1))
    sd["decoder.head.0.gamma"] = (1.0 + 0.1 * torch.randn(dim, 1, 1, 1, generator=g)).to(torch.bfloat16)
    conv("decoder.head.2", 3, dim, (3, 3, 3))
    return sd
