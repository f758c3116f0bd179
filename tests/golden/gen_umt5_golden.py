"""Generates tests/golden/umt5_small.npz with `transformers.UMT5EncoderModel` (the published umT5 implementation
that ships in this image) on the deterministic weights of oracle/umt5_oracle.make_weights:

    python tests/golden/gen_umt5_golden.py

The fixture pins oracle/umt5_oracle.py (row A11) without needing transformers at test time.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import umt5_oracle as o  # noqa: E402

CFG = dict(vocab_size=128, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2)
L, N_VALID = 160, 141   # relative distances beyond max_dist=128 exercise the last (clamped) bucket


def wan_to_hf_keys(sd, num_layers):
    out = {"shared.weight": sd["token_embedding.weight"], "encoder.embed_tokens.weight": sd["token_embedding.weight"],
           "encoder.final_layer_norm.weight": sd["norm.weight"]}
    for i in range(num_layers):
        s, d = f"blocks.{i}.", f"encoder.block.{i}.layer."
        out[d + "0.layer_norm.weight"] = sd[s + "norm1.weight"]
        for n in "qkvo":
            out[d + f"0.SelfAttention.{n}.weight"] = sd[s + f"attn.{n}.weight"]
        out[d + "0.SelfAttention.relative_attention_bias.weight"] = sd[s + "pos_embedding.embedding.weight"]
        out[d + "1.layer_norm.weight"] = sd[s + "norm2.weight"]
        out[d + "1.DenseReluDense.wi_0.weight"] = sd[s + "ffn.gate.0.weight"]
        out[d + "1.DenseReluDense.wi_1.weight"] = sd[s + "ffn.fc1.weight"]
        out[d + "1.DenseReluDense.wo.weight"] = sd[s + "ffn.fc2.weight"]
    return out


def hf_model(cfg: o.T5Config, sd):
    from transformers import UMT5Config, UMT5EncoderModel
    hc = UMT5Config(vocab_size=cfg.vocab_size, d_model=cfg.dim, d_kv=cfg.head_dim, d_ff=cfg.dim_ffn,
                    num_layers=cfg.num_layers, num_heads=cfg.num_heads,
                    relative_attention_num_buckets=cfg.num_buckets, relative_attention_max_distance=cfg.max_dist,
                    dropout_rate=0.0, layer_norm_epsilon=cfg.eps, feed_forward_proj="gated-gelu")
    m = UMT5EncoderModel(hc).eval().float()
    m.load_state_dict(wan_to_hf_keys(sd, cfg.num_layers), strict=True)
    return m


def main():
    cfg = o.T5Config(**CFG)
    sd = o.make_weights(cfg, seed=4321)
    g = torch.Generator().manual_seed(11)
    ids = torch.randint(1, cfg.vocab_size, (L,), generator=g)
    mask = torch.zeros(L, dtype=torch.long)
    mask[:N_VALID] = 1
    ids[N_VALID:] = 0
    with torch.no_grad():
        out = hf_model(cfg, sd)(input_ids=ids[None], attention_mask=mask[None]).last_hidden_state[0]
    wsum = float(sum(v.double().abs().sum() for v in sd.values()))
    np.savez_compressed(Path(__file__).with_name("umt5_small.npz"), ids=ids.numpy(), mask=mask.numpy(),
                        hidden=out.numpy().astype(np.float32), weight_abs_sum=np.float64(wsum),
                        cfg=np.array([CFG[k] for k in ("vocab_size", "dim", "dim_attn", "dim_ffn", "num_heads", "num_layers")]))
    print("wrote umt5_small.npz", out.shape, "weight |sum|", wsum)


if __name__ == "__main__":
    main()
