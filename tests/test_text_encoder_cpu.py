"""Host logic of the umT5 prompt encoder (row A11) that needs no GPU: the offset-indexed position-bias table, prompt
cleaning, the key-name bridge, and the no-fallback rule."""
import pytest
import torch

from oracle import umt5_oracle as o


def test_bias_by_offset_matches_oracle_position_bias():
    from infinicube_b200.videogen.text_encoder import bias_by_offset
    tab = torch.randn(32, 5, generator=torch.Generator().manual_seed(0))
    for L in (1, 7, 130, 512):
        full = o.position_bias(tab, L, L, 32, 128)          # [H, L, L]
        b = bias_by_offset(tab, L, 32, 128)                 # [H, 2L-1]
        assert tuple(b.shape) == (5, 2 * L - 1)
        i = torch.arange(L)[:, None]
        j = torch.arange(L)[None, :]
        assert torch.equal(b[:, (j - i + L - 1)], full)


def test_product_bucket_equals_oracle_bucket():
    from infinicube_b200.videogen.text_encoder import relative_position_bucket
    rel = torch.arange(-600, 601)
    assert torch.equal(relative_position_bucket(rel, 32, 128), o.relative_position_bucket(rel, 32, 128))


def test_prompt_cleaning():
    from infinicube_b200.videogen.text_encoder import WanPrompter
    p = WanPrompter(text_len=512)
    assert p.process_prompt("  The video\n is   about &amp;amp; a\tdriving scene. ") == "The video is about & a driving scene."


def test_no_cpu_fallback():
    from infinicube_b200._lib import ICError
    from infinicube_b200.videogen.text_encoder import WanPrompter, WanTextEncoder
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ICError):
        WanTextEncoder(device="cuda:0")
    p = WanPrompter()
    with pytest.raises(ICError):
        p.tokenize("a prompt")
    with pytest.raises(ICError):
        p.encode_ids(torch.ones(4, dtype=torch.long), torch.ones(4, dtype=torch.long))


def test_hf_key_bridge_matches_oracle_bridge():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).parent / "golden"))
    import gen_umt5_golden as gen
    from infinicube_b200.videogen.text_encoder import _from_hf_keys
    cfg = o.T5Config(vocab_size=16, dim=32, dim_attn=32, dim_ffn=64, num_heads=2, num_layers=2)
    sd = o.make_weights(cfg, seed=1)
    back = _from_hf_keys(gen.wan_to_hf_keys(sd, 2), 2)
    assert back.keys() == sd.keys() and all(torch.equal(back[k], sd[k]) for k in sd)


def test_synthetic_state_dict_has_reference_key_names():
    from infinicube_b200.videogen.text_encoder import T5Config, synthetic_t5_state_dict
    kw = dict(vocab_size=16, dim=64, dim_attn=64, dim_ffn=128, num_heads=1, num_layers=2)
    assert synthetic_t5_state_dict(T5Config(**kw)).keys() == o.make_weights(o.T5Config(**kw)).keys()


def test_prompter_tokenizes_like_wans_huggingface_tokenizer(tmp_path):
    """fetch_tokenizer + tokenize with a tiny unigram tokenizer built on the spot (the real google/umt5-xxl
    sentencepiece file is not in this image): cleaned prompt, EOS appended, right padding to text_len, mask."""
    tr = pytest.importorskip("transformers")
    from infinicube_b200.videogen.text_encoder import WanPrompter
    words = ["the", "video", "is", "about", "a", "driving", "scene", "weather", "clear"]
    vocab = ([("<pad>", 0.0), ("</s>", 0.0), ("<unk>", 0.0), ("▁", -2.0)] + [("▁" + w, -3.0 - 0.1 * i) for i, w in enumerate(words)] +
             [(c, -8.0) for c in "abcdefghijklmnopqrstuvwxyz."])
    tr.T5Tokenizer(vocab=vocab, extra_ids=0).save_pretrained(str(tmp_path / "umt5"))
    p = WanPrompter(tokenizer_path=str(tmp_path / "umt5"), text_len=12)
    ids, mask = p.tokenize("  the   weather\n is clear ")
    assert ids.shape == (12,) and mask.shape == (12,)
    n = int(mask.sum())
    assert ids[:n].tolist() == [4, 11, 6, 12, 1]            # ▁the ▁weather ▁is ▁clear </s>
    assert ids[n:].eq(0).all() and mask[:n].eq(1).all() and mask[n:].eq(0).all()
    long_ids, long_mask = p.tokenize(" ".join(["the"] * 40))  # truncation keeps the EOS in the last slot
    assert int(long_mask.sum()) == 12 and int(long_ids[-1]) == 1
