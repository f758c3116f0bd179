"""Size-independent properties at BASELINE.json's full size (Wan2.1-1.3B, 93 x 480 x 832 -> 37 440 tokens), where the
fp32 CPU oracle would take hours: softmax normalisation, key-permutation invariance, GEMM linearity, determinism
and the zero-init guidance invariant, all through the C ABI."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
S, H, D = 37440, 12, 1536


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def dev():
    from infinicube_b200 import _lib
    _lib.require_device()
    return torch.device("cuda:0")


def test_attention_rows_sum_to_one_and_keys_commute(dev):
    from infinicube_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(S, D, device=dev, generator=g).bfloat16()
    k = torch.randn(S, D, device=dev, generator=g).bfloat16()
    scale = 1.0 / math.sqrt(128)
    # V = const per channel -> softmax(QK^T) V = that constant for every query, whatever the scores
    vt = (torch.arange(D, device=dev, dtype=torch.float32) % 7 - 3).bfloat16()[:, None].expand(D, S).contiguous()
    out = torch.empty(S, D, device=dev, dtype=torch.bfloat16)
    ops.fmha(q, k, vt, out, H, scale)
    assert torch.allclose(out.float(), vt[:, 0].float()[None, :].expand(S, D), atol=2e-2)
    # permuting the keys (and V^T columns the same way) cannot change the result beyond accumulation order
    v = torch.randn(S, D, device=dev, generator=g).bfloat16()
    o1 = torch.empty_like(out)
    o2 = torch.empty_like(out)
    ops.fmha(q, k, v.t().contiguous(), o1, H, scale)
    perm = torch.randperm(S, device=dev, generator=g)
    ops.fmha(q, k[perm].contiguous(), v[perm].t().contiguous(), o2, H, scale)
    # both sides carry independent bf16 roundings of P (2^-9 relative per probability) and of the output
    assert rel_l2(o2, o1) < 8e-3
    # 256 sampled queries against an fp32 softmax over all 37 440 keys
    idx = torch.randperm(S, device=dev, generator=g)[:256]
    qs = q[idx].float().view(256, H, 128).transpose(0, 1)             # [H, 256, 128]
    kh = k.float().view(S, H, 128).transpose(0, 1)                    # [H, S, 128]
    vh = v.float().view(S, H, 128).transpose(0, 1)
    ref = torch.softmax(qs @ kh.transpose(1, 2) * scale, dim=-1) @ vh  # [H, 256, 128]
    ref = ref.transpose(0, 1).reshape(256, D)
    assert rel_l2(o1[idx], ref) < 6e-3
    # determinism: same launch, same bits
    o3 = torch.empty_like(out)
    ops.fmha(q, k, v.t().contiguous(), o3, H, scale)
    assert torch.equal(o1, o3)


def test_gemm_linearity_and_identity(dev):
    from infinicube_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(S, D, device=dev, generator=g).bfloat16()
    # identity weights reproduce the input exactly (bf16 values are exact in the fp32 accumulator)
    eye = torch.eye(D, device=dev).bfloat16()
    out = torch.empty(S, D, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, eye, out_bf16=out)
    assert torch.equal(out, a)
    # linearity in the weights with an fp32 output: A(W1 + W2) = A W1 + A W2 for weights whose sum is exact in bf16
    w1 = (torch.randint(-8, 8, (D, D), device=dev, generator=g).float() / 64).bfloat16()
    w2 = (torch.randint(-8, 8, (D, D), device=dev, generator=g).float() / 64).bfloat16()
    o12 = torch.empty(S, D, device=dev)
    o1 = torch.empty(S, D, device=dev)
    o2 = torch.empty(S, D, device=dev)
    ops.gemm(a, (w1.float() + w2.float()).bfloat16(), out_f32=o12)
    ops.gemm(a, w1, out_f32=o1)
    ops.gemm(a, w2, out_f32=o2)
    assert rel_l2(o12, o1 + o2) < 1e-5
    # residual epilogue: x += 1 * (A I) twice == x + 2A
    x = torch.zeros(S, D, device=dev)
    ops.gemm(a, eye, resid=x)
    ops.gemm(a, eye, resid=x)
    assert torch.equal(x, 2 * a.float())


def test_full_size_layer_invariants(dev):
    """One full-size layer (37 440 tokens): determinism and the zero-init buffer-embedder invariant."""
    from infinicube_b200.videogen.pipeline import WanDiTEngine, WanModelConfig, synthetic_context, synthetic_state_dict
    cfg = WanModelConfig(num_layers=1)
    eng = WanDiTEngine(cfg, 24, 60, 104, guide_channels=32, device=dev)
    sd = synthetic_state_dict(cfg, 32, dev, seed=5, zero_guidance=True)
    eng.load_state_dict(sd)
    eng.set_context(0, synthetic_context("a street", cfg, dev))
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(16, 24, 60, 104, generator=g).to(dev)
    guide = torch.randn(32, 24, 60, 104, generator=g).to(dev)
    outs = []
    for use_guide in (True, False, True):
        eng.set_guidance(guide if use_guide else None)
        o = torch.empty(S, 64, device=dev)
        eng.forward(lat, 900.0, 0, o)
        outs.append(o)
    assert not torch.isnan(outs[0]).any() and float(outs[0].abs().max()) < 1e3
    assert torch.equal(outs[0], outs[1])   # zero-initialised embedder == no guidance (inference.py:86-88)
    assert torch.equal(outs[0], outs[2])   # deterministic
    assert eng.flops_per_forward > 1e13
