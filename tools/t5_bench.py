"""Times the umT5-XXL prompt encoder (row A11) at its real size (24 blocks, 4096 wide, 64 heads, 512 tokens) with
device-generated random weights: ms per prompt, algorithmic TFLOP/s, weight bytes streamed per prompt vs HBM peak
(M = 512 rows makes the encoder weight-bandwidth bound: 11.4 GB of weights are read once per prompt).
Writes gpurun_out/t5_bench.json."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from infinicube_b200.videogen.text_encoder import T5Config, WanTextEncoder  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    cfg = T5Config(num_layers=layers, vocab_size=32768)  # the gather reads 512 rows whatever the table height
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, s):
        return (torch.randn(*shape, generator=g, device=dev) * s).bfloat16()

    sd = {"token_embedding.weight": rnd(cfg.vocab_size, cfg.dim, s=1.0), "norm.weight": torch.ones(cfg.dim, device=dev)}
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = torch.ones(cfg.dim, device=dev)
        sd[p + "norm2.weight"] = torch.ones(cfg.dim, device=dev)
        for n in "qkv":
            sd[p + f"attn.{n}.weight"] = rnd(cfg.dim_attn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "attn.o.weight"] = rnd(cfg.dim, cfg.dim_attn, s=cfg.dim_attn ** -0.5)
        sd[p + "ffn.gate.0.weight"] = rnd(cfg.dim_ffn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "ffn.fc1.weight"] = rnd(cfg.dim_ffn, cfg.dim, s=cfg.dim ** -0.5)
        sd[p + "ffn.fc2.weight"] = rnd(cfg.dim, cfg.dim_ffn, s=cfg.dim_ffn ** -0.5)
        sd[p + "pos_embedding.embedding.weight"] = rnd(cfg.num_buckets, cfg.num_heads, s=0.5).float().cpu()
    enc = WanTextEncoder(cfg, dev)
    enc.load_state_dict(sd)
    del sd
    ids = torch.randint(1, cfg.vocab_size, (cfg.text_len,))
    mask = torch.zeros(cfg.text_len, dtype=torch.long)
    mask[:40] = 1
    for _ in range(3):
        out = enc(ids, mask, zero_from_row=40)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        out = enc(ids, mask, zero_from_row=40)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    wbytes = cfg.num_layers * 2 * (4 * cfg.dim * cfg.dim_attn + 3 * cfg.dim * cfg.dim_ffn)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    res = {"what": "umT5-XXL encoder, one 512-token prompt", "layers": cfg.num_layers, "ms_per_prompt": ms,
           "launches": enc.launch_count, "tflops": enc.flops_per_prompt / (ms * 1e-3) / 1e12,
           "weight_bytes": wbytes, "weight_gbs": wbytes / (ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs"),
           "finite": bool(torch.isfinite(out.float()).all()), "zero_rows_ok": bool((out[40:] == 0).all())}
    print(json.dumps(res))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "t5_bench.json").write_text(json.dumps(res))


if __name__ == "__main__":
    main()
