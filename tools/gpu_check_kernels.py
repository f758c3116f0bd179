"""Developer harness (NOT the parity gate): runs each kernel against a torch-on-GPU fp32 reference,
each case in its own subprocess with a timeout, so a trapped kernel cannot hide the other results.

    python tools/gpu_check_kernels.py            # all cases
    python tools/gpu_check_kernels.py gemm_small # one case (runs in-process)
Writes gpurun_out/kernel_check.json.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _err(a, b):
    import torch
    a = a.float()
    b = b.float()
    d = (a - b).abs()
    return {"max_abs": d.max().item(), "rel_l2": (d.norm() / (b.norm() + 1e-30)).item(),
            "ref_absmax": b.abs().max().item(), "nan": bool(torch.isnan(a).any().item())}


def _time(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def case_gemm(M, N, K, mode):
    import torch
    from infinicube_b200 import ops
    torch.manual_seed(0)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    ref = a.float() @ b.float().t()
    res = {}
    if mode == "f32":
        out = torch.zeros(M, N, device="cuda")
        ops.gemm(a, b, out_f32=out)
        torch.cuda.synchronize()
        res = _err(out, ref)
        if res["rel_l2"] > 1e-2:
            bad = ((out - ref).abs() > 0.05 * ref.abs().max()).nonzero()
            res["n_bad"] = int(bad.shape[0])
            res["first_bad"] = bad[:8].tolist()
            res["rows_bad_hist"] = torch.bincount((bad[:, 0] % 128), minlength=128)[:16].tolist()
            res["cols_bad_hist"] = torch.bincount((bad[:, 1] % 64), minlength=64)[:16].tolist()
            res["out_sample"] = out[:2, :8].tolist()
            res["ref_sample"] = ref[:2, :8].tolist()
    elif mode == "bf16_bias_ss":
        bias = torch.randn(N, device="cuda")
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        nt = (N + ops.gemm_block_n(N) - 1) // ops.gemm_block_n(N)
        ss = torch.zeros(M, nt, device="cuda")
        ops.gemm(a, b, bias=bias, out_bf16=out, rowss=ss)
        torch.cuda.synchronize()
        r = ref + bias
        res = _err(out, r)
        res["ss"] = _err(ss.sum(1), (r.bfloat16().float() ** 2).sum(1))
    elif mode == "gelu":
        bias = torch.randn(N, device="cuda")
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, b, bias=bias, act=1, out_bf16=out)
        torch.cuda.synchronize()
        res = _err(out, torch.nn.functional.gelu(ref + bias, approximate="tanh"))
    elif mode == "resid_gate":
        bias = torch.randn(N, device="cuda")
        gate = torch.randn(N, device="cuda")
        x = torch.randn(M, N, device="cuda")
        x0 = x.clone()
        ops.gemm(a, b, bias=bias, resid=x, gate=gate)
        torch.cuda.synchronize()
        want = x0 + gate * (ref + bias)
        res = _err(x, want)
        if res["rel_l2"] > 1e-3:
            bad_rows = ((x - want).abs().amax(1) > 1e-2 * want.abs().max()).nonzero().flatten()
            res["bad_rows"] = [int(bad_rows.numel()), bad_rows[:4].tolist(), bad_rows[-4:].tolist()]
    elif mode == "rowbias_add":
        bias = torch.randn(M, device="cuda")
        add = torch.randn(M, N, device="cuda")
        out = torch.zeros(M, N, device="cuda")
        ops.gemm(a, b, bias=bias, bias_per_row=True, out_f32=out, addend=add)
        torch.cuda.synchronize()
        res = _err(out, ref + bias[:, None] + add)
    return res


def case_gemm_perf(M, N, K):
    import torch
    from infinicube_b200 import ops
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ms = _time(lambda: ops.gemm(a, b, out_bf16=out))
    ms_t = _time(lambda: torch.matmul(a, b.t()))
    fl = 2.0 * M * N * K
    return {"ms": ms, "tflops": fl / ms / 1e9, "torch_ms": ms_t, "torch_tflops": fl / ms_t / 1e9}


def case_gemm_resid_perf(M, N, K):
    import torch
    from infinicube_b200 import ops
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    bias = torch.randn(N, device="cuda")
    gate = torch.randn(N, device="cuda") * 0.1
    x = torch.randn(M, N, device="cuda")
    ms = _time(lambda: ops.gemm(a, b, bias=bias, resid=x, gate=gate))
    fl = 2.0 * M * N * K
    return {"ms": ms, "tflops": fl / ms / 1e9, "rmw_GBps": (M * N * 8 + M * K * 2) / ms / 1e6}


def case_rmsnorm_rope_perf(rows=37440, D=1536):
    import torch
    from infinicube_b200 import ops
    src = torch.randn(rows, 2 * D, device="cuda").bfloat16()
    ss = torch.rand(rows, 12, device="cuda") * D
    w = torch.randn(D, device="cuda")
    dst = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    ms = _time(lambda: ops.rmsnorm_rope(src[:, :D], ss, 0, 6, w, dst, 1e-6, None))
    return {"ms": ms, "GBps": rows * D * 4 / ms / 1e6}


def case_rmsnorm_rope_perf_rope(f=24, hh=30, ww=52, D=1536):
    """The q / k RMSNorm + RoPE pass at the bench size.  A/B: run once plain and once with ICB_RMSROPE_V2=1."""
    import os
    import torch
    from infinicube_b200 import ops
    rows = f * hh * ww
    src = torch.randn(rows, 2 * D, device="cuda").bfloat16()
    ss = torch.rand(rows, 12, device="cuda") * D
    w = torch.randn(D, device="cuda")
    dst = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    tabs = tuple(torch.randn(n, p, 2, device="cuda").contiguous() for n, p in ((f, 22), (hh, 21), (ww, 21)))
    ms = _time(lambda: ops.rmsnorm_rope(src[:, :D], ss, 0, 6, w, dst, 1e-6, tabs))
    return {"ms": ms, "GBps": rows * D * 4 / ms / 1e6, "variant": os.environ.get("ICB_RMSROPE_V2", "0")}


def _attn_ref(q, k, v, H, scale):
    import torch
    Sq = q.shape[0]
    S = k.shape[0]
    qh = q.float().view(Sq, H, 128).transpose(0, 1)
    kh = k.float().view(S, H, 128).transpose(0, 1)
    vh = v.float().view(S, H, 128).transpose(0, 1)
    p = torch.softmax(qh @ kh.transpose(1, 2) * scale, dim=-1)
    return (p @ vh).transpose(0, 1).reshape(Sq, H * 128)


def case_fmha(Sq, S, H, n_seg=1):
    import torch
    from infinicube_b200 import ops
    torch.manual_seed(1)
    D = H * 128
    q = torch.randn(Sq, D, device="cuda").bfloat16()
    k = torch.randn(S * n_seg, D, device="cuda").bfloat16()
    v = torch.randn(S * n_seg, D, device="cuda").bfloat16()
    scale = 1.0 / 128 ** 0.5
    # large logits on a few keys exercise the lazy rescale path
    k[S // 2] *= 4.0
    if n_seg == 1:
        vt = v.t().contiguous()
        kk = k
        kst = vst = 0
    else:
        # segment layout of the all-gather buffer: [seg][K (S x D) || V^T (D x S)]
        buf = torch.zeros(n_seg, 2 * S * D, device="cuda", dtype=torch.bfloat16)
        for s in range(n_seg):
            buf[s, : S * D] = k[s * S:(s + 1) * S].reshape(-1)
            buf[s, S * D:] = v[s * S:(s + 1) * S].t().contiguous().reshape(-1)
        kk = buf.view(-1)[: S * D].view(S, D)
        vt = buf.view(-1)[S * D: 2 * S * D].view(D, S)
        kst = vst = 2 * S * D
    out = torch.zeros(Sq, D, device="cuda", dtype=torch.bfloat16)
    ops.fmha(q, kk, vt, out, H, scale, seg_len=S, n_seg=n_seg, k_seg_stride=kst, vt_seg_stride=vst)
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, H, scale)
    res = _err(out, ref)
    if res["rel_l2"] > 2e-2:
        res["out_sample"] = out[:2, :6].float().tolist()
        res["ref_sample"] = ref[:2, :6].tolist()
        d = (out.float() - ref).abs()
        res["row_err_first16"] = d.max(1).values[:16].tolist()
        res["row_err_128_144"] = d.max(1).values[128:144].tolist()
    return res


def case_fmha_perf(S, H):
    import torch
    from infinicube_b200 import ops
    D = H * 128
    q = torch.randn(S, D, device="cuda").bfloat16()
    k = torch.randn(S, D, device="cuda").bfloat16()
    vt = torch.randn(D, S, device="cuda").bfloat16()
    out = torch.zeros(S, D, device="cuda", dtype=torch.bfloat16)
    scale = 1.0 / 128 ** 0.5
    ms = _time(lambda: ops.fmha(q, k, vt, out, H, scale), iters=3, warm=1)
    fl = 4.0 * S * S * D
    res = {"ms": ms, "tflops": fl / ms / 1e9}
    try:
        qh = q.view(S, H, 128).transpose(0, 1)[None]
        kh = k.view(S, H, 128).transpose(0, 1)[None]
        vh = vt.t().contiguous().view(S, H, 128).transpose(0, 1)[None]
        ms_t = _time(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh), iters=3, warm=1)
        res.update({"sdpa_ms": ms_t, "sdpa_tflops": fl / ms_t / 1e9})
    except Exception as ex:  # noqa: BLE001
        res["sdpa_error"] = repr(ex)[:200]
    return res


def case_elementwise():
    import torch
    from infinicube_b200 import ops
    torch.manual_seed(2)
    res = {}
    S, D = 1000, 1536
    x = torch.randn(S, D, device="cuda") * 3 + 1
    mul = torch.randn(D, device="cuda") * 0.1
    add = torch.randn(D, device="cuda") * 0.1
    out = torch.zeros(S, D, device="cuda", dtype=torch.bfloat16)
    ops.ln_modulate(x, mul, add, out, True)
    ref = torch.nn.functional.layer_norm(x, (D,), eps=1e-6) * (1 + mul) + add
    res["ln_modulate"] = _err(out, ref)
    # rmsnorm + rope
    f, hh, ww = 4, 10, 25
    S = f * hh * ww
    src = torch.randn(S, D, device="cuda").bfloat16()
    w = torch.rand(D, device="cuda") + 0.5
    ss = torch.zeros(S, 6, device="cuda")
    ss[:, :6] = (src.float() ** 2).view(S, 6, 256).sum(-1)
    import math

    def tab(npos, npair, axis_dim):
        j = torch.arange(npair, dtype=torch.float64)
        th = torch.pow(torch.tensor(10000.0, dtype=torch.float64), -2.0 * j / axis_dim)
        a = torch.arange(npos, dtype=torch.float64)[:, None] * th[None]
        return torch.stack([a.cos(), a.sin()], -1).float().cuda().contiguous()

    tf, th_, tw = tab(f, 22, 44), tab(hh, 21, 42), tab(ww, 21, 42)
    dst = torch.zeros(S, D, device="cuda", dtype=torch.bfloat16)
    ops.rmsnorm_rope(src, ss, 0, 6, w, dst, rope=(tf, th_, tw))
    xn = src.float() * torch.rsqrt((src.float() ** 2).mean(-1, keepdim=True) + 1e-6) * w
    tok = torch.arange(S, device="cuda")
    pf, ph, pw = tok // (hh * ww), (tok // ww) % hh, tok % ww
    cs = torch.cat([tf[pf], th_[ph], tw[pw]], 1)  # [S, 64, 2]
    xr = xn.view(S, D // 128, 64, 2)
    a, b = xr[..., 0], xr[..., 1]
    c, s = cs[:, None, :, 0], cs[:, None, :, 1]
    ref = torch.stack([a * c - b * s, a * s + b * c], -1).reshape(S, D)
    res["rmsnorm_rope"] = _err(dst, ref)
    # patchify / unpatchify
    C_, F_, H_, W_ = 16, 3, 8, 12
    lat = torch.randn(C_, F_, H_, W_, device="cuda")
    ntok = F_ * (H_ // 2) * (W_ // 2)
    pa = torch.zeros(ntok, 64, device="cuda", dtype=torch.bfloat16)
    ops.patchify(lat, pa)
    ref = lat.view(C_, F_, H_ // 2, 2, W_ // 2, 2).permute(1, 2, 4, 0, 3, 5).reshape(ntok, 64)
    res["patchify"] = _err(pa, ref)
    hp = torch.randn(ntok, 64, device="cuda")
    hn = torch.randn(ntok, 64, device="cuda")
    lat2 = lat.clone()
    ops.unpatchify_cfg_step(lat2, hp, hn, (C_, F_, H_, W_), 5.0, -0.01)

    def unp(h):
        return h.view(F_, H_ // 2, W_ // 2, 2, 2, C_).permute(5, 0, 1, 3, 2, 4).reshape(C_, F_, H_, W_)

    ref = lat + (unp(hn) + 5.0 * (unp(hp) - unp(hn))) * -0.01
    res["unpatchify_cfg_step"] = _err(lat2, ref)
    return res


CASES = {
    "gemm_1tile": lambda: case_gemm(128, 256, 64, "f32"),
    "gemm_ktiles": lambda: case_gemm(128, 256, 512, "f32"),
    "gemm_multi": lambda: case_gemm(1000, 1536, 1536, "f32"),
    "gemm_n64": lambda: case_gemm(300, 64, 1536, "f32"),
    "gemm_n128": lambda: case_gemm(300, 128, 256, "f32"),
    "gemm_k64": lambda: case_gemm(2048, 1536, 64, "f32"),
    "gemm_bias_ss": lambda: case_gemm(2048, 3072, 1536, "bf16_bias_ss"),
    "gemm_gelu": lambda: case_gemm(512, 8960, 1536, "gelu"),
    "gemm_resid": lambda: case_gemm(700, 1536, 8960, "resid_gate"),
    "gemm_rowbias": lambda: case_gemm(1536, 2048, 1536, "rowbias_add"),
    "gemm_persist": lambda: case_gemm(37440, 1536, 256, "f32"),
    "fmha_1tile": lambda: case_fmha(256, 128, 1),
    "fmha_2tile": lambda: case_fmha(256, 256, 2),
    "fmha_tails": lambda: case_fmha(300, 520, 2),
    "fmha_2048": lambda: case_fmha(2048, 2048, 12),
    "fmha_cross": lambda: case_fmha(1000, 512, 12),
    "fmha_seg2": lambda: case_fmha(512, 520, 2, n_seg=2),
    "elementwise": case_elementwise,
    "perf_gemm_qk": lambda: case_gemm_perf(37440, 3072, 1536),
    "perf_gemm_ffn1": lambda: case_gemm_perf(37440, 8960, 1536),
    "perf_gemm_ffn2": lambda: case_gemm_perf(37440, 1536, 8960),
    "gemm_resid_ragged": lambda: case_gemm(37440 // 8 + 72, 1536, 512, "resid_gate"),
    "gemm_resid_full": lambda: case_gemm(37440, 1536, 1536, "resid_gate"),
    "gemm_resid_k256": lambda: case_gemm(37440, 1536, 256, "resid_gate"),
    "gemm_resid_clip": lambda: case_gemm(37440 - 8, 1536, 256, "resid_gate"),
    "perf_gemm_oproj_resid": lambda: case_gemm_resid_perf(37440, 1536, 1536),
    "perf_gemm_ffn2_resid": lambda: case_gemm_resid_perf(37440, 1536, 8960),
    "perf_rmsnorm": case_rmsnorm_rope_perf,
    "perf_rmsnorm_rope": case_rmsnorm_rope_perf_rope,
    "perf_fmha_full": lambda: case_fmha_perf(37440, 12),
}


def case_conv_perf(C, T, H, W):
    """Our implicit-GEMM causal 3x3x3 conv (C -> C, channels-last bf16) against cuDNN conv3d through torch on the same
    shape (channels_last_3d bf16, symmetric padding: same FLOPs), same process."""
    import torch
    from infinicube_b200.videogen import vae as V
    torch.manual_seed(0)
    w = (torch.randn(C, C, 3, 3, 3) / (27 * C) ** 0.5).to(torch.bfloat16)
    b = torch.zeros(C, dtype=torch.bfloat16)
    m = V.WanVideoVAE.__new__(V.WanVideoVAE)
    m.device = torch.device("cuda:0")
    m.sd = {"c.weight": w, "c.bias": b}
    m.convs = {}
    m._add_conv("c")
    x = torch.randn(T, H, W, C, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(x)
    ms = _time(lambda: m._conv(x, "c", out=out), iters=3, warm=1)
    fl = 2.0 * T * H * W * 27 * C * C
    res = {"ms": ms, "tflops": fl / ms / 1e9, "shape": [T, H, W, C]}
    try:
        xc = x.permute(3, 0, 1, 2)[None].contiguous(memory_format=torch.channels_last_3d)
        wc = w.cuda().contiguous(memory_format=torch.channels_last_3d)
        ms_t = _time(lambda: torch.nn.functional.conv3d(xc, wc, padding=1), iters=3, warm=1)
        res.update({"cudnn_ms": ms_t, "cudnn_tflops": fl / ms_t / 1e9})
    except Exception as ex:  # noqa: BLE001
        res["cudnn_error"] = repr(ex)[:300]
    return res


CASES["perf_conv96_fullres"] = lambda: case_conv_perf(96, 24, 480, 832)
CASES["perf_conv192_halfres"] = lambda: case_conv_perf(192, 24, 240, 416)


def case_gemm_shard_perf():
    """The DiT block's GEMM shapes at the single-GPU row count and at the per-rank row count of the 8-GPU default
    layout (37 440 / 4): what a launch costs beyond its FLOPs when the problem is 4x smaller."""
    import torch
    from infinicube_b200 import ops
    out = {}
    for M in (37440, 9360):
        for name, N, K, resid in (("qkv", 4608, 1536, False), ("o_proj", 1536, 1536, True), ("cross_q", 1536, 1536, False),
                                  ("ffn1", 8960, 1536, False), ("ffn2", 1536, 8960, True)):
            a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
            b = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
            if resid:
                bias = torch.randn(N, device="cuda")
                gate = torch.randn(N, device="cuda") * 0.1
                x = torch.randn(M, N, device="cuda")
                ms = _time(lambda: ops.gemm(a, b, bias=bias, resid=x, gate=gate), iters=20)
            else:
                o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
                ms = _time(lambda: ops.gemm(a, b, out_bf16=o), iters=20)
            fl = 2.0 * M * N * K
            fn = (lambda: ops.gemm(a, b, bias=bias, resid=x, gate=gate)) if resid else (lambda: ops.gemm(a, b, out_bf16=o))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                fn()
            enq = (time.perf_counter() - t0) / 20 * 1e6   # host enqueue cost: the event timing is only valid above it
            torch.cuda.synchronize()
            out[f"{name}_M{M}"] = {"us": ms * 1e3, "tflops": fl / ms / 1e9, "host_enqueue_us": enq}
    for name in ("qkv", "o_proj", "cross_q", "ffn1", "ffn2"):
        out[f"{name}_fixed_us"] = (4 * out[f"{name}_M9360"]["us"] - out[f"{name}_M37440"]["us"]) / 3.0
    return out


CASES["perf_gemm_shard"] = case_gemm_shard_perf


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--one" and sys.argv[2] in CASES:
        print("RESULT " + json.dumps(CASES[sys.argv[2]]()))
        return
    if len(sys.argv) == 2 and sys.argv[1] in CASES:  # single exact case: run in-process (for ncu)
        print("RESULT " + json.dumps(CASES[sys.argv[1]]()))
        return
    names = [n for n in CASES if not sys.argv[1:] or any(n.startswith(p) for p in sys.argv[1:])]
    out = {}
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=300)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                out[n] = json.loads(lines[-1][7:])
            else:
                out[n] = {"error": (r.stdout[-1500:] + "\n" + r.stderr[-1500:])}
        except subprocess.TimeoutExpired:
            out[n] = {"error": "timeout"}
        out[n]["wall_s"] = round(time.time() - t0, 1)
        print(n, json.dumps(out[n])[:600], flush=True)
    od = ROOT / "gpurun_out"
    od.mkdir(exist_ok=True)
    (od / "kernel_check.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
