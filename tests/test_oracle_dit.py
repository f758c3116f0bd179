"""Pins the DiT oracle against what the reference offers for this path: the flow-match schedule known
answers, the zero-init buffer-embedder invariant (videogen/inference.py:86-88), and the
algorithmic FLOP table of SURVEY §8(d)."""
import math

import pytest
import torch

from oracle import wan_dit_oracle as o


def test_schedule_known_answers():
    s = o.flow_match_sigmas(50, 5.0)
    assert s.dtype == torch.float32 and len(s) == 51 and s[-1] == 0
    for i, v in {0: 1.0, 1: 0.995935, 2: 0.991736, 3: 0.987395, 24: 0.844156, 48: 0.172414, 49: 0.092593}.items():
        assert abs(float(s[i]) - v) < 1e-6, (i, float(s[i]))
    assert abs(float(s[1] - s[0]) + 0.0040650) < 1e-6
    assert abs(float(s[50] - s[49]) + 0.0925926) < 1e-6
    t = s[:3] * 1000
    for a, b in zip(t.tolist(), (1000.0, 995.935, 991.735)):
        assert abs(a - b) < 2e-3


def test_product_scheduler_matches_oracle():
    from infinicube_b200.videogen.pipeline import FlowMatchScheduler
    sch = FlowMatchScheduler().set_timesteps(50, shift=5.0)
    s = o.flow_match_sigmas(50, 5.0)
    assert torch.equal(sch.sigmas, s[:-1])
    assert all(abs(sch.delta_sigma(i) - float(s[i + 1] - s[i])) == 0 for i in range(50))


def test_flop_model_matches_survey_table():
    cfg = o.WanConfig.wan_1_3b()
    fwd = o.dit_flops_per_forward(cfg, 37440)
    assert abs(fwd / 3.557e14 - 1) < 2e-3
    assert abs(o.dit_flops_per_forward(o.WanConfig.wan_14b(), 37440) / 2.061e15 - 1) < 2e-3
    blk = o.dit_flops_per_forward(o.WanConfig(num_layers=1), 2048)
    assert abs(blk / 2.175e11 - 1) < 0.05


def test_patchify_roundtrip_and_layout():
    lat = torch.arange(16 * 2 * 4 * 6, dtype=torch.float32).view(16, 2, 4, 6)
    tok = o.patchify(lat)
    assert tok.shape == (2 * 2 * 3, 64)
    # column = c*4 + py*2 + px
    assert tok[0, 5 * 4 + 1 * 2 + 0] == lat[5, 0, 1, 0]
    head = torch.randn(12, 64)
    v = o.unpatchify(head, 16, 2, 4, 6)
    # column = (py*2+px)*C + c
    assert v[3, 1, 2 * 1 + 1, 2 * 2 + 0] == head[(1 * 2 + 1) * 3 + 2, (1 * 2 + 0) * 16 + 3]


def test_rope_is_identity_at_origin_and_norm_preserving():
    ang = o.rope_angles(2, 3, 4, 128)
    assert ang.shape == (24, 64) and torch.all(ang[0] == 0)
    x = torch.randn(24, 256)
    y = o.rope_apply(x, ang, 2)
    assert torch.allclose(y[0], x[0])
    assert torch.allclose(y.norm(dim=1), x.norm(dim=1), rtol=1e-5)
    # split 22 | 21 | 21: frame index only drives the first 22 pairs
    assert torch.all(ang[12, 22:] == 0) and ang[12, 0] == 1.0


@pytest.fixture(scope="module")
def tiny():
    cfg = o.WanConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16)
    g = torch.Generator().manual_seed(0)
    return cfg, torch.randn(16, 2, 4, 4, generator=g), torch.randn(16, 64, generator=g), torch.randn(32, 2, 4, 4, generator=g)


def test_zero_init_buffer_embedder_reproduces_plain_wan(tiny):
    cfg, lat, ctx, guide_lat = tiny
    sd0 = o.make_weights(cfg, zero_guidance=True)
    g = o.guidance_tokens(guide_lat, sd0, cfg)
    assert torch.count_nonzero(g) == 0
    a = o.dit_forward(lat, 700.0, ctx, sd0, cfg, guide=g)
    b = o.dit_forward(lat, 700.0, ctx, sd0, cfg, guide=None)
    assert torch.equal(a, b)
    sd1 = o.make_weights(cfg, zero_guidance=False)
    c = o.dit_forward(lat, 700.0, ctx, sd1, cfg, guide=o.guidance_tokens(guide_lat, sd1, cfg))
    assert not torch.allclose(c, o.dit_forward(lat, 700.0, ctx, sd1, cfg, guide=None))


def test_temporal_shard_equivalence(tiny):
    """Every op except self-attention is token-local: a frame shard with global RoPE offsets and
    the full K/V set reproduces the unsharded block."""
    cfg, lat, ctx, _ = tiny
    sd = o.make_weights(cfg)
    full, xfull = o.dit_forward(lat, 500.0, ctx, sd, cfg, layers=0, return_tokens=True)
    ang_full = o.rope_angles(2, 2, 2, cfg.head_dim)
    ang_1 = o.rope_angles(1, 2, 2, cfg.head_dim, frame0=1)
    assert torch.equal(ang_full[4:], ang_1)


def test_denoise_one_step_is_euler(tiny):
    cfg, lat, ctx, _ = tiny
    sd = o.make_weights(cfg)
    ctx2 = ctx.flip(0)
    out = o.denoise(lat, ctx, ctx2, sd, cfg, None, steps_to_run=1)
    sig = o.flow_match_sigmas()
    vp = o.dit_forward(lat, 1000.0, ctx, sd, cfg)
    vn = o.dit_forward(lat, 1000.0, ctx2, sd, cfg)
    v = o.unpatchify(vn + 5.0 * (vp - vn), 16, 2, 4, 4)
    assert torch.allclose(out, lat + v * float(sig[1] - sig[0]), atol=1e-6)
