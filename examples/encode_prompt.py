"""The umT5 prompt encoder on its own: load the Wan2.1 text-encoder checkpoint (or run with random weights), encode a
prompt and hand the context to a DiT engine.  Needs a B200.

    python examples/encode_prompt.py [models/Wan-AI/Wan2.1-T2V-1.3B]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from infinicube_b200.videogen.text_encoder import (T5Config, WanPrompter, WanTextEncoder,  # noqa: E402
                                                   synthetic_t5_state_dict)


def main():
    dev = torch.device("cuda:0")
    model_dir = Path(sys.argv[1]) if len(sys.argv) > 1 else None
    prompter = WanPrompter(text_len=512)
    if model_dir and (model_dir / "models_t5_umt5-xxl-enc-bf16.pth").exists():
        enc = WanTextEncoder(T5Config(), dev)
        enc.load_state_dict(torch.load(model_dir / "models_t5_umt5-xxl-enc-bf16.pth", map_location="cpu", weights_only=True))
        prompter.fetch_models(enc)
        prompter.fetch_tokenizer(str(model_dir / "google" / "umt5-xxl"))
        ctx = prompter.encode_prompt("The video is about a driving scene captured at daytime. The weather is clear.")
    else:   # no checkpoint on disk: a 4-block encoder with random weights and hand-made token ids
        cfg = T5Config(vocab_size=4096, num_layers=4)
        enc = WanTextEncoder(cfg, dev)
        enc.load_state_dict(synthetic_t5_state_dict(cfg))
        prompter.fetch_models(enc)
        ids = torch.zeros(512, dtype=torch.long)
        ids[:17] = torch.randint(2, 4096, (17,))
        ids[16] = 1                                   # </s>
        ctx = prompter.encode_ids(ids, (ids > 0).long())
    n = int((ctx.float().abs().sum(-1) > 0).sum())
    print(f"context {tuple(ctx.shape)} {ctx.dtype}, {n} non-zero rows, {enc.launch_count} kernel launches")
    # ctx is what WanDiTEngine.set_context(slot, ctx) / ic_dit_set_context expects


if __name__ == "__main__":
    main()
