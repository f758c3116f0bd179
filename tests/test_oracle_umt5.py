"""Pins oracle/umt5_oracle.py (the prompt encoder) against the published umT5 implementation in
`transformers` (live, when importable) and against the committed fixture that implementation produced."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import umt5_oracle as o

GOLD = Path(__file__).parent / "golden"


def test_bucket_known_answers():
    # bidirectional, 32 buckets, max distance 128: 16 buckets per direction, 8 exact + 8 logarithmic
    rel = torch.tensor([0, -1, -7, -8, -11, -12, -127, -128, -500, 1, 7, 8, 15, 16, 90, 91, 128, 511])
    want = torch.tensor([0, 1, 7, 8, 8, 9, 15, 15, 15, 17, 23, 24, 25, 26, 30, 31, 31, 31])
    assert torch.equal(o.relative_position_bucket(rel, 32, 128), want)


def test_position_bias_layout():
    tab = torch.arange(32 * 3, dtype=torch.float32).view(32, 3)
    b = o.position_bias(tab, 5, 7, 32, 128)
    assert b.shape == (3, 5, 7)
    assert b[2, 4, 0] == tab[4, 2] and b[1, 0, 6] == tab[16 + 6, 1] and b[0, 3, 3] == tab[0, 0]


def test_oracle_matches_hf_fixture():
    z = np.load(GOLD / "umt5_small.npz")
    keys = ("vocab_size", "dim", "dim_attn", "dim_ffn", "num_heads", "num_layers")
    cfg = o.T5Config(**{k: int(v) for k, v in zip(keys, z["cfg"])})
    sd = o.make_weights(cfg, seed=4321)
    wsum = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(wsum - float(z["weight_abs_sum"])) < 1e-6 * wsum, "torch's CPU generator stream changed: regenerate"
    out = o.encode(torch.from_numpy(z["ids"]), torch.from_numpy(z["mask"]), sd, cfg)
    ref = torch.from_numpy(z["hidden"])
    assert float((out - ref).abs().max()) <= 2e-5
    # prompter post-processing: rows past the prompt length are zero, the others untouched
    n = int(z["mask"].sum())
    pe = o.encode_prompt_ids(torch.from_numpy(z["ids"]), torch.from_numpy(z["mask"]), sd, cfg)
    assert torch.equal(pe[:n], out[:n]) and float(pe[n:].abs().max()) == 0.0


@pytest.mark.parametrize("L,n_valid,heads,dim", [(64, 64, 4, 256), (300, 211, 2, 128)])
def test_oracle_matches_transformers_live(L, n_valid, heads, dim):
    pytest.importorskip("transformers")
    sys.path.insert(0, str(GOLD))
    import gen_umt5_golden as gen
    cfg = o.T5Config(vocab_size=97, dim=dim, dim_attn=dim, dim_ffn=2 * dim + 8, num_heads=heads, num_layers=2)
    sd = o.make_weights(cfg, seed=L)
    g = torch.Generator().manual_seed(L)
    ids = torch.randint(1, cfg.vocab_size, (L,), generator=g)
    mask = torch.zeros(L, dtype=torch.long)
    mask[:n_valid] = 1
    ids[n_valid:] = 0
    with torch.no_grad():
        ref = gen.hf_model(cfg, sd)(input_ids=ids[None], attention_mask=mask[None]).last_hidden_state[0]
    out = o.encode(ids, mask, sd, cfg)
    assert float((out - ref).abs().max()) <= 2e-5


def test_hf_key_bridge_round_trip():
    sys.path.insert(0, str(GOLD))
    import gen_umt5_golden as gen
    cfg = o.T5Config(vocab_size=16, dim=32, dim_attn=32, dim_ffn=64, num_heads=2, num_layers=2)
    sd = o.make_weights(cfg, seed=1)
    back = o.hf_to_wan_keys(gen.wan_to_hf_keys(sd, 2), 2)
    assert back.keys() == sd.keys() and all(torch.equal(back[k], sd[k]) for k in sd)
