"""VAE oracle self-consistency: the whole-sequence formulation the CUDA path implements equals the reference's
frame-chunked formulation with its explicit feature cache, tile enumeration matches the
9-tile layout the survey documents, and masks follow DiffSynth's linear ramps."""
import torch

from oracle import wan_vae_oracle as v


def test_full_sequence_equals_chunked_feature_cache():
    sd = v.make_weights()
    g = torch.Generator().manual_seed(0)
    z = torch.randn(16, 3, 4, 6, generator=g)
    a, b = v.decode_full(z, sd), v.decode_chunked(z, sd)
    assert a.shape == (3, 9, 32, 48)
    assert (a - b).abs().max() < 2e-5
    vid = torch.rand(3, 9, 32, 48, generator=g) * 2 - 1
    e1, e2 = v.encode_full(vid, sd), v.encode_chunked(vid, sd)
    assert e1.shape == (16, 3, 4, 6)
    assert (e1 - e2).abs().max() < 2e-5
    # single-frame case: no temporal resampling at all
    assert v.decode_full(z[:, :1], sd).shape == (3, 1, 32, 48)
    assert (v.decode_full(z[:, :1], sd) - a[:, :1]).abs().max() < 2e-5  # causality: frame 0 ignores the future


def test_tile_layout_and_masks():
    tasks = v.tile_tasks(60, 104, (30, 52), (15, 26))
    assert len(tasks) == 9 and tasks[0] == (0, 30, 0, 52) and tasks[-1] == (30, 60, 52, 104)
    m = v.build_mask(240, 416, (True, False, False, True), (120, 208))
    assert m[0, 415] == 1.0 and abs(float(m[239, 415]) - 1 / 120) < 1e-7 and abs(float(m[0, 0]) - 1 / 208) < 1e-7
    from infinicube_b200.videogen.vae import tile_tasks
    mine = tile_tasks(60, 104, (30, 52), (15, 26))
    assert [(a, b, c, d) for a, b, c, d, _, _ in mine] == tasks
    assert [t[4] for t in mine] == [False] * 6 + [True] * 3 and [t[5] for t in mine][:3] == [False, False, True]


def test_frames_conversion_roundtrip():
    fr = torch.randint(0, 256, (2, 4, 4, 3), dtype=torch.uint8)
    vid = v.frames_to_video(fr)
    assert vid.min() >= -1 and vid.max() <= 1
    back = v.video_to_frames(vid)
    assert (back.int() - fr.int()).abs().max() <= 1  # ((x+1)*127.5) truncation
