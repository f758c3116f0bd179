"""CPU fp32 restatement of the Wan2.1 3-D causal VAE (encode / decode, plain and tiled) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

PARITY UNPINNED: the reference runs `WanVideoVAE` inside the un-vendored `diffsynth` dependency
(/root/reference/pyproject.toml:71; constructed at infinicube/videogen/inference.py:66-80 through the
"Wan2.1_VAE.pth" ModelConfig, used by `self.pipe(...)` at :216-226 for the buffer encodes and the final decode).
This file restates the published Wan2.1 VAE (z_dim 16, base 96, mult [1,2,4,4],
2 / 3 res-blocks, temporal down [F,T,T], causal 3-D convs with a 2-frame feature cache, first frame handled
without temporal resampling) and DiffSynth's tiled encode/decode blending (tile (30,52), stride (15,26), linear
ramps).  Two formulations are given and tested equal: `*_chunked` follows the reference's frame-chunk loop with an
explicit feature cache; `*_full` processes the whole sequence at once (what the CUDA path implements).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

CACHE_T = 2
LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
               -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
              1.1253, 2.8251, 1.9160]


# --------------------------------------------------------------------------------------------------
# architecture tables
# --------------------------------------------------------------------------------------------------
def encoder_layers(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_down=(False, True, True)):
    """Flat list of ('res', cin, cout) / ('down2d', c) / ('down3d', c) in module order (encoder.downsamples.{i})."""
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
    """Flat list for decoder.upsamples.{i}; res blocks after an upsample take dim//2 inputs."""
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


def make_weights(seed: int = 4321, dim: int = 96, z_dim: int = 16) -> Dict[str, torch.Tensor]:
    """Deterministic random-init VAE weights under the official module names; bf16-representable values
    (the reference runs the VAE in torch_dtype=bfloat16).  Scales keep activations O(1) through the stack."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def bf(t):
        return t.to(torch.bfloat16).to(torch.float32)

    def conv(name, cout, cin, k, gain=1.0):
        fan = cin * math.prod(k)
        sd[name + ".weight"] = bf(torch.randn(cout, cin, *k, generator=g) * (gain / math.sqrt(fan)))
        sd[name + ".bias"] = bf(torch.randn(cout, generator=g) * 0.02)

    def res(prefix, cin, cout):
        sd[prefix + ".residual.0.gamma"] = bf(1.0 + 0.1 * torch.randn(cin, 1, 1, 1, generator=g))
        conv(prefix + ".residual.2", cout, cin, (3, 3, 3), 1.4)
        sd[prefix + ".residual.3.gamma"] = bf(1.0 + 0.1 * torch.randn(cout, 1, 1, 1, generator=g))
        conv(prefix + ".residual.6", cout, cout, (3, 3, 3), 0.7)
        if cin != cout:
            conv(prefix + ".shortcut", cout, cin, (1, 1, 1))

    def attn(prefix, c):
        sd[prefix + ".norm.gamma"] = bf(1.0 + 0.1 * torch.randn(c, 1, 1, generator=g))
        conv(prefix + ".to_qkv", 3 * c, c, (1, 1))
        conv(prefix + ".proj", c, c, (1, 1), 0.5)

    enc, enc_top = encoder_layers(dim)
    conv("encoder.conv1", dim, 3, (3, 3, 3))
    for i, l in enumerate(enc):
        p = f"encoder.downsamples.{i}"
        if l[0] == "res":
            res(p, l[1], l[2])
        else:
            conv(p + ".resample.1", l[1], l[1], (3, 3))
            if l[0] == "down3d":
                conv(p + ".time_conv", l[1], l[1], (3, 1, 1))
    res("encoder.middle.0", enc_top, enc_top)
    attn("encoder.middle.1", enc_top)
    res("encoder.middle.2", enc_top, enc_top)
    sd["encoder.head.0.gamma"] = bf(1.0 + 0.1 * torch.randn(enc_top, 1, 1, 1, generator=g))
    conv("encoder.head.2", 2 * z_dim, enc_top, (3, 3, 3))
    conv("conv1", 2 * z_dim, 2 * z_dim, (1, 1, 1))
    conv("conv2", z_dim, z_dim, (1, 1, 1))
    dec, dec_top = decoder_layers(dim)
    conv("decoder.conv1", dec_top, z_dim, (3, 3, 3))
    res("decoder.middle.0", dec_top, dec_top)
    attn("decoder.middle.1", dec_top)
    res("decoder.middle.2", dec_top, dec_top)
    for i, l in enumerate(dec):
        p = f"decoder.upsamples.{i}"
        if l[0] == "res":
            res(p, l[1], l[2])
        else:
            conv(p + ".resample.1", l[1] // 2, l[1], (3, 3))
            if l[0] == "up3d":
                conv(p + ".time_conv", 2 * l[1], l[1], (3, 1, 1))
    sd["decoder.head.0.gamma"] = bf(1.0 + 0.1 * torch.randn(dim, 1, 1, 1, generator=g))
    conv("decoder.head.2", 3, dim, (3, 3, 3))
    return sd


# --------------------------------------------------------------------------------------------------
# primitives (x: [C, T, H, W])
# --------------------------------------------------------------------------------------------------
def causal_conv3d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, cache: Optional[torch.Tensor] = None,
                  stride=(1, 1, 1)) -> torch.Tensor:
    """CausalConv3d: temporal padding 2*pt in front only (minus the cached frames), spatial padding symmetric."""
    kt, kh, kw = w.shape[2:]
    pt, ph, pw = 2 * (kt // 2), kh // 2, kw // 2
    if stride[0] == 2:  # the strided time_conv uses padding (0,0,0)
        pt = 0
    if cache is not None and pt > 0:
        x = torch.cat([cache, x], dim=1)
        pt -= cache.shape[1]
    x = F.pad(x, (pw, pw, ph, ph, max(pt, 0), 0))
    return F.conv3d(x[None], w, b, stride=stride)[0]


def rms_norm(x: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """RMS_norm(channel_first): F.normalize(x, dim=channels) * sqrt(C) * gamma."""
    c = x.shape[0]
    return F.normalize(x, dim=0) * math.sqrt(c) * gamma.reshape(c, *([1] * (x.dim() - 1)))


def _res_full(x, sd, p):
    h = x
    y = causal_conv3d(F.silu(rms_norm(x, sd[p + ".residual.0.gamma"])), sd[p + ".residual.2.weight"], sd[p + ".residual.2.bias"])
    y = causal_conv3d(F.silu(rms_norm(y, sd[p + ".residual.3.gamma"])), sd[p + ".residual.6.weight"], sd[p + ".residual.6.bias"])
    if p + ".shortcut.weight" in sd:
        h = causal_conv3d(x, sd[p + ".shortcut.weight"], sd[p + ".shortcut.bias"])
    return y + h


def _attn(x, sd, p):
    """AttentionBlock: per frame, single head over h*w tokens."""
    c, t, h, w = x.shape
    xn = rms_norm(x, sd[p + ".norm.gamma"].reshape(c, 1, 1, 1))
    out = []
    for i in range(t):
        f = xn[:, i]
        qkv = F.conv2d(f[None], sd[p + ".to_qkv.weight"], sd[p + ".to_qkv.bias"])[0].reshape(3, c, h * w)
        q, k, v = qkv[0].t(), qkv[1].t(), qkv[2].t()
        a = F.scaled_dot_product_attention(q[None, None], k[None, None], v[None, None])[0, 0]
        a = a.t().reshape(c, h, w)
        out.append(F.conv2d(a[None], sd[p + ".proj.weight"], sd[p + ".proj.bias"])[0])
    return torch.stack(out, dim=1) + x


def _conv2d_frames(x, w, b, stride=1, pad=None):
    c, t, h, ww = x.shape
    f = x.permute(1, 0, 2, 3)
    if pad is not None:
        f = F.pad(f, pad)
        y = F.conv2d(f, w, b, stride=stride)
    else:
        y = F.conv2d(f, w, b, stride=stride, padding=1)
    return y.permute(1, 0, 2, 3)


def _upsample2x(x):
    c, t, h, w = x.shape
    return F.interpolate(x.permute(1, 0, 2, 3), scale_factor=(2.0, 2.0), mode="nearest-exact").permute(1, 0, 2, 3)


# --------------------------------------------------------------------------------------------------
# full-sequence formulation
# --------------------------------------------------------------------------------------------------
def decode_full(z: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """z [16, T, h, w] (normalised latents) -> video [3, 4T-3, 8h, 8w] (unclamped)."""
    mean = torch.tensor(LATENT_MEAN).reshape(-1, 1, 1, 1)
    std = torch.tensor(LATENT_STD).reshape(-1, 1, 1, 1)
    x = z * std + mean
    x = causal_conv3d(x, sd["conv2.weight"], sd["conv2.bias"])
    x = causal_conv3d(x, sd["decoder.conv1.weight"], sd["decoder.conv1.bias"])
    x = _res_full(x, sd, "decoder.middle.0")
    x = _attn(x, sd, "decoder.middle.1")
    x = _res_full(x, sd, "decoder.middle.2")
    layers, _ = decoder_layers()
    for i, l in enumerate(layers):
        p = f"decoder.upsamples.{i}"
        if l[0] == "res":
            x = _res_full(x, sd, p)
            continue
        if l[0] == "up3d" and x.shape[1] > 1:
            # frame 0 is not temporally upsampled; the causal time_conv sees only frames >= 1 (zero history)
            c = x.shape[0]
            y = causal_conv3d(x[:, 1:], sd[p + ".time_conv.weight"], sd[p + ".time_conv.bias"])  # [2C, T-1, h, w]
            y = y.reshape(2, c, *y.shape[1:])
            y = torch.stack((y[0], y[1]), dim=2).reshape(c, -1, *y.shape[3:])                     # interleave in time
            x = torch.cat([x[:, :1], y], dim=1)
        x = _conv2d_frames(_upsample2x(x), sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"])
    x = F.silu(rms_norm(x, sd["decoder.head.0.gamma"]))
    return causal_conv3d(x, sd["decoder.head.2.weight"], sd["decoder.head.2.bias"])


def encode_full(video: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """video [3, T, H, W] in [-1,1], T % 4 == 1 -> normalised latent mean [16, (T-1)/4+1, H/8, W/8]."""
    x = causal_conv3d(video, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"])
    layers, _ = encoder_layers()
    for i, l in enumerate(layers):
        p = f"encoder.downsamples.{i}"
        if l[0] == "res":
            x = _res_full(x, sd, p)
            continue
        x = _conv2d_frames(x, sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"], stride=2, pad=(0, 1, 0, 1))
        if l[0] == "down3d" and x.shape[1] > 1:
            # frame 0 passes through; frame k >= 1 = time_conv(x[2k-2], x[2k-1], x[2k])
            y = F.conv3d(x[None], sd[p + ".time_conv.weight"], sd[p + ".time_conv.bias"], stride=(2, 1, 1))[0]
            x = torch.cat([x[:, :1], y], dim=1)
    x = _res_full(x, sd, "encoder.middle.0")
    x = _attn(x, sd, "encoder.middle.1")
    x = _res_full(x, sd, "encoder.middle.2")
    x = F.silu(rms_norm(x, sd["encoder.head.0.gamma"]))
    x = causal_conv3d(x, sd["encoder.head.2.weight"], sd["encoder.head.2.bias"])
    x = causal_conv3d(x, sd["conv1.weight"], sd["conv1.bias"])
    mu = x[:16]
    mean = torch.tensor(LATENT_MEAN).reshape(-1, 1, 1, 1)
    std = torch.tensor(LATENT_STD).reshape(-1, 1, 1, 1)
    return (mu - mean) / std


# --------------------------------------------------------------------------------------------------
# chunked formulation with the reference's feature cache (validates the full-sequence form)
# --------------------------------------------------------------------------------------------------
class _Cache:
    def __init__(self):
        self.store: List = []
        self.idx = 0

    def begin(self):
        self.idx = 0

    def conv(self, x, w, b):
        """cached CausalConv3d call as in ResidualBlock / Encoder3d / Decoder3d forward."""
        i = self.idx
        if i >= len(self.store):
            self.store.append(None)
        cache_x = x[:, -CACHE_T:].clone()
        if cache_x.shape[1] < 2 and self.store[i] is not None:
            cache_x = torch.cat([self.store[i][:, -1:], cache_x], dim=1)
        y = causal_conv3d(x, w, b, self.store[i])
        self.store[i] = cache_x
        self.idx += 1
        return y


def _res_chunk(x, sd, p, cache: _Cache):
    h = x
    if p + ".shortcut.weight" in sd:
        h = causal_conv3d(x, sd[p + ".shortcut.weight"], sd[p + ".shortcut.bias"])
    y = cache.conv(F.silu(rms_norm(x, sd[p + ".residual.0.gamma"])), sd[p + ".residual.2.weight"], sd[p + ".residual.2.bias"])
    y = cache.conv(F.silu(rms_norm(y, sd[p + ".residual.3.gamma"])), sd[p + ".residual.6.weight"], sd[p + ".residual.6.bias"])
    return y + h


def decode_chunked(z: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    mean = torch.tensor(LATENT_MEAN).reshape(-1, 1, 1, 1)
    std = torch.tensor(LATENT_STD).reshape(-1, 1, 1, 1)
    x_all = causal_conv3d(z * std + mean, sd["conv2.weight"], sd["conv2.bias"])
    cache = _Cache()
    layers, _ = decoder_layers()
    outs = []
    for ti in range(x_all.shape[1]):
        cache.begin()
        x = x_all[:, ti:ti + 1]
        x = cache.conv(x, sd["decoder.conv1.weight"], sd["decoder.conv1.bias"])
        x = _res_chunk(x, sd, "decoder.middle.0", cache)
        x = _attn(x, sd, "decoder.middle.1")
        x = _res_chunk(x, sd, "decoder.middle.2", cache)
        for i, l in enumerate(layers):
            p = f"decoder.upsamples.{i}"
            if l[0] == "res":
                x = _res_chunk(x, sd, p, cache)
                continue
            if l[0] == "up3d":
                k = cache.idx
                if k >= len(cache.store):
                    cache.store.append(None)
                if cache.store[k] is None:
                    cache.store[k] = "Rep"
                    cache.idx += 1
                else:
                    cache_x = x[:, -CACHE_T:].clone()
                    if cache_x.shape[1] < 2 and not isinstance(cache.store[k], str):
                        cache_x = torch.cat([cache.store[k][:, -1:], cache_x], dim=1)
                    if cache_x.shape[1] < 2 and isinstance(cache.store[k], str):
                        cache_x = torch.cat([torch.zeros_like(cache_x), cache_x], dim=1)
                    w, b = sd[p + ".time_conv.weight"], sd[p + ".time_conv.bias"]
                    y = causal_conv3d(x, w, b, None if isinstance(cache.store[k], str) else cache.store[k])
                    cache.store[k] = cache_x
                    cache.idx += 1
                    c = x.shape[0]
                    y = y.reshape(2, c, *y.shape[1:])
                    x = torch.stack((y[0], y[1]), dim=2).reshape(c, -1, *y.shape[3:])
            x = _conv2d_frames(_upsample2x(x), sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"])
        x = F.silu(rms_norm(x, sd["decoder.head.0.gamma"]))
        x = cache.conv(x, sd["decoder.head.2.weight"], sd["decoder.head.2.bias"])
        outs.append(x)
    return torch.cat(outs, dim=1)


def encode_chunked(video: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    t = video.shape[1]
    n_iter = 1 + (t - 1) // 4
    cache = _Cache()
    layers, _ = encoder_layers()
    outs = []
    for it in range(n_iter):
        cache.begin()
        x = video[:, :1] if it == 0 else video[:, 1 + 4 * (it - 1):1 + 4 * it]
        x = cache.conv(x, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"])
        for i, l in enumerate(layers):
            p = f"encoder.downsamples.{i}"
            if l[0] == "res":
                x = _res_chunk(x, sd, p, cache)
                continue
            x = _conv2d_frames(x, sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"], stride=2, pad=(0, 1, 0, 1))
            if l[0] == "down3d":
                k = cache.idx
                if k >= len(cache.store):
                    cache.store.append(None)
                if cache.store[k] is None:
                    cache.store[k] = x.clone()
                    cache.idx += 1
                else:
                    cache_x = x[:, -1:].clone()
                    x = F.conv3d(torch.cat([cache.store[k][:, -1:], x], dim=1)[None], sd[p + ".time_conv.weight"],
                                 sd[p + ".time_conv.bias"], stride=(2, 1, 1))[0]
                    cache.store[k] = cache_x
                    cache.idx += 1
        x = _res_chunk(x, sd, "encoder.middle.0", cache)
        x = _attn(x, sd, "encoder.middle.1")
        x = _res_chunk(x, sd, "encoder.middle.2", cache)
        x = F.silu(rms_norm(x, sd["encoder.head.0.gamma"]))
        x = cache.conv(x, sd["encoder.head.2.weight"], sd["encoder.head.2.bias"])
        outs.append(x)
    x = torch.cat(outs, dim=1)
    x = causal_conv3d(x, sd["conv1.weight"], sd["conv1.bias"])
    mean = torch.tensor(LATENT_MEAN).reshape(-1, 1, 1, 1)
    std = torch.tensor(LATENT_STD).reshape(-1, 1, 1, 1)
    return (x[:16] - mean) / std


# --------------------------------------------------------------------------------------------------
# tiling (DiffSynth WanVideoVAE.tiled_decode / tiled_encode)
# --------------------------------------------------------------------------------------------------
def tile_tasks(H: int, W: int, size: Tuple[int, int], stride: Tuple[int, int]):
    tasks = []
    for h in range(0, H, stride[0]):
        if h - stride[0] >= 0 and h - stride[0] + size[0] >= H:
            continue
        for w in range(0, W, stride[1]):
            if w - stride[1] >= 0 and w - stride[1] + size[1] >= W:
                continue
            tasks.append((h, h + size[0], w, w + size[1]))
    return tasks


def _mask_1d(length: int, left_bound: bool, right_bound: bool, border: int) -> torch.Tensor:
    x = torch.ones(length)
    if not left_bound:
        x[:border] = (torch.arange(border) + 1) / border
    if not right_bound:
        x[-border:] = torch.flip((torch.arange(border) + 1) / border, dims=(0,))
    return x


def build_mask(H: int, W: int, is_bound, border) -> torch.Tensor:
    h = _mask_1d(H, is_bound[0], is_bound[1], border[0])[:, None].expand(H, W)
    w = _mask_1d(W, is_bound[2], is_bound[3], border[1])[None, :].expand(H, W)
    return torch.minimum(h, w)


def tiled_decode(z: torch.Tensor, sd, tile_size=(30, 52), tile_stride=(15, 26), decode_fn=decode_full) -> torch.Tensor:
    _, T, H, W = z.shape
    out_t = T * 4 - 3
    weight = torch.zeros(1, out_t, H * 8, W * 8)
    values = torch.zeros(3, out_t, H * 8, W * 8)
    for h, h_, w, w_ in tile_tasks(H, W, tile_size, tile_stride):
        y = decode_fn(z[:, :, h:h_, w:w_], sd)
        m = build_mask(y.shape[2], y.shape[3], (h == 0, h_ >= H, w == 0, w_ >= W),
                       ((tile_size[0] - tile_stride[0]) * 8, (tile_size[1] - tile_stride[1]) * 8))
        values[:, :, h * 8:h * 8 + y.shape[2], w * 8:w * 8 + y.shape[3]] += y * m
        weight[:, :, h * 8:h * 8 + y.shape[2], w * 8:w * 8 + y.shape[3]] += m
    return (values / weight).clamp(-1, 1)


def tiled_encode(video: torch.Tensor, sd, tile_size=(30, 52), tile_stride=(15, 26), encode_fn=encode_full) -> torch.Tensor:
    size = (tile_size[0] * 8, tile_size[1] * 8)
    stride = (tile_stride[0] * 8, tile_stride[1] * 8)
    _, T, H, W = video.shape
    out_t = (T + 3) // 4
    weight = torch.zeros(1, out_t, H // 8, W // 8)
    values = torch.zeros(16, out_t, H // 8, W // 8)
    for h, h_, w, w_ in tile_tasks(H, W, size, stride):
        y = encode_fn(video[:, :, h:h_, w:w_], sd)
        m = build_mask(y.shape[2], y.shape[3], (h == 0, h_ >= H, w == 0, w_ >= W),
                       ((size[0] - stride[0]) // 8, (size[1] - stride[1]) // 8))
        values[:, :, h // 8:h // 8 + y.shape[2], w // 8:w // 8 + y.shape[3]] += y * m
        weight[:, :, h // 8:h // 8 + y.shape[2], w // 8:w // 8 + y.shape[3]] += m
    return values / weight


def frames_to_video(frames_u8: torch.Tensor) -> torch.Tensor:
    """uint8 [T, H, W, 3] -> [-1,1] float [3, T, H, W]  (pixel * (2/255) - 1, Appendix A.8)."""
    return frames_u8.permute(3, 0, 1, 2).to(torch.float32) * (2.0 / 255.0) - 1.0


def video_to_frames(video: torch.Tensor) -> torch.Tensor:
    """[3, T, H, W] -> uint8 [T, H, W, 3]: ((x+1)*127.5).clip(0,255).uint8."""
    return ((video.clamp(-1, 1) + 1.0) * 127.5).clip(0, 255).to(torch.uint8).permute(1, 2, 3, 0)
