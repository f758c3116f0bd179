"""GPU parity of the Wan VAE path (tcgen05 implicit-GEMM convolutions + helper kernels, through the C ABI) against
the fp32 CPU oracle.  Activations are bf16 on the GPU (like the reference's torch_dtype=bfloat16) through ~60
convolutions, so the bar is relative L2 <= 3e-2 on the decoded video / encoded latents and <= 6 / 255 mean
absolute error on the uint8 frames."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 3e-2


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def vae():
    from infinicube_b200 import _lib
    _lib.require_device()
    from oracle import wan_vae_oracle as o
    from infinicube_b200.videogen.vae import WanVideoVAE
    sd = o.make_weights()
    return o, sd, WanVideoVAE(dict(sd), device="cuda:0")


def test_conv_kernel_vs_conv3d(vae):
    """The implicit-GEMM kernel alone: causal 3x3x3, ragged spatial size (not a multiple of the 16x8 pixel tile)."""
    import ctypes as C
    from infinicube_b200._lib import check, lib
    g = torch.Generator().manual_seed(3)
    # (96, 96) / (96, 8) take the multi-frame kernel (4 output frames per tile): T = 3 (one partial frame group),
    # 5 and 9 (full + partial groups); the others the generic one
    for cin, cout, T in ((96, 192, 3), (384, 384, 3), (32, 96, 3), (96, 96, 3), (96, 96, 5), (96, 96, 9), (96, 8, 6)):
        H, W = 11, 21
        x = torch.randn(T, H, W, cin, generator=g).bfloat16()
        w = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5).bfloat16()
        b = torch.randn(cout, generator=g)
        res = torch.randn(T, H, W, cout, generator=g).bfloat16()
        xr = x.float().permute(3, 0, 1, 2)
        ref = torch.nn.functional.conv3d(torch.nn.functional.pad(xr, (1, 1, 1, 1, 2, 0))[None], w.float(), b)[0]
        ref = ref.permute(1, 2, 3, 0) + res.float()
        taps = [(it - 2, ih - 1, iw - 1) for it in range(3) for ih in range(3) for iw in range(3)]
        w2 = w.float().reshape(cout, cin, 27).permute(0, 2, 1).reshape(cout, -1).bfloat16().cuda()
        out = torch.zeros(T, H, W, cout, dtype=torch.bfloat16, device="cuda")
        tp = (C.c_int * 81)(*[v for t in taps for v in t])
        xd, bd, rd = x.cuda(), b.cuda(), res.cuda()
        check(lib().ic_conv_cl(C.c_void_p(xd.data_ptr()), T, H, W, cin, C.c_void_p(w2.data_ptr()), C.c_void_p(bd.data_ptr()),
                               tp, 27, C.c_void_p(out.data_ptr()), T, H, W, cout, cout, C.c_void_p(rd.data_ptr()), cout,
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        assert rel_l2(out, ref) < 5e-3, (cin, cout, rel_l2(out, ref))


def test_decode_parity(vae):
    o, sd, m = vae
    g = torch.Generator().manual_seed(0)
    z = torch.randn(16, 3, 8, 10, generator=g)
    ref = o.decode_full(z, sd).clamp(-1, 1)
    frames, vid = m.decode(z, tiled=False, want_f32=True)
    assert vid.shape == (3, 9, 64, 80) and frames.shape == (9, 64, 80, 3)
    assert rel_l2(vid, ref) < TOL, rel_l2(vid, ref)
    ref_u8 = o.video_to_frames(ref)
    assert (frames.cpu().int() - ref_u8.int()).abs().float().mean() < 6.0


def test_tiled_decode_parity(vae):
    o, sd, m = vae
    g = torch.Generator().manual_seed(1)
    z = torch.randn(16, 2, 9, 12, generator=g)
    ref = o.tiled_decode(z, sd, tile_size=(6, 8), tile_stride=(3, 4))
    _, vid = m.decode(z, tiled=True, tile_size=(6, 8), tile_stride=(3, 4), want_f32=True)
    assert vid.shape == ref.shape == (3, 5, 72, 96)
    assert rel_l2(vid, ref) < TOL, rel_l2(vid, ref)


def test_encode_parity(vae):
    o, sd, m = vae
    g = torch.Generator().manual_seed(2)
    fr = torch.randint(0, 256, (9, 64, 80, 3), generator=g, dtype=torch.uint8)
    ref = o.encode_full(o.frames_to_video(fr), sd)
    lat = m.encode(fr.cuda(), tiled=False)
    assert lat.shape == ref.shape == (16, 3, 8, 10)
    assert rel_l2(lat, ref) < TOL, rel_l2(lat, ref)
    ref_t = o.tiled_encode(o.frames_to_video(fr[:5, :, :]), sd, tile_size=(4, 6), tile_stride=(2, 3))
    # 64x80 px with 32x48-px tiles, 16x24 stride -> ragged last tiles
    lat_t = m.encode(fr[:5].cuda(), tiled=True, tile_size=(4, 6), tile_stride=(2, 3))
    assert rel_l2(lat_t, ref_t) < TOL, rel_l2(lat_t, ref_t)


def test_encode_validation(vae):
    _, _, m = vae
    with pytest.raises(ValueError):
        m.encode(torch.zeros(8, 64, 80, 3, dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        m.encode(torch.zeros(9, 60, 80, 3, dtype=torch.uint8, device="cuda"))


def test_encode_many_sharded_tiles_equal_single_rank(vae):
    """The two guidance buffers of a call are encoded together; with several ranks their tiles are dealt jointly.  Two
    'ranks' are run on one device with the all-reduce replaced by the sum of their recorded partial accumulators: the
    result must equal the single-rank encode to fp32 summation order (same tiles, same blending arithmetic)."""
    from infinicube_b200.videogen.vae import WanVideoVAE
    o, sd, m = vae
    g = torch.Generator().manual_seed(5)
    a = torch.randint(0, 256, (5, 64, 80, 3), generator=g, dtype=torch.uint8).cuda()
    b = torch.randint(0, 256, (5, 64, 80, 3), generator=g, dtype=torch.uint8).cuda()
    kw = dict(tiled=True, tile_size=(4, 6), tile_stride=(2, 3))
    ref = m.encode_many([a, b], **kw)
    assert torch.equal(ref[0], m.encode(a, **kw)) and torch.equal(ref[1], m.encode(b, **kw))
    world = 3
    ranks = [WanVideoVAE(dict(sd), device="cuda:0", world_size=world, rank=r) for r in range(world)]
    partial = []
    for r in ranks:      # pass 1: record every rank's partial accumulators
        r._all_reduce = lambda values, weight: partial.append((values.clone(), weight.clone()))
        r.encode_many([a, b], **kw)
    assert len(partial) == world

    def summed(values, weight):
        values.copy_(sum(p[0] for p in partial))
        weight.copy_(sum(p[1] for p in partial))

    ranks[1]._all_reduce = summed   # pass 2: any rank, with the "all-reduce" delivering the sum
    got = ranks[1].encode_many([a, b], **kw)
    for x, y in zip(got, ref):
        assert torch.isfinite(x).all() and torch.allclose(x, y, atol=1e-5, rtol=1e-5)   # fp32 sums in another order
