"""Times the Wan VAE at the bench size (93 x 480 x 832, tiled like the reference): encode of one uint8 buffer video
and decode of one latent, with algorithmic conv FLOPs.  Writes gpurun_out/vae_bench.json."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from infinicube_b200.videogen.vae import WanVideoVAE, synthetic_vae_state_dict  # noqa: E402


def main():
    tiled = "--untiled" not in sys.argv
    dev = torch.device("cuda:0")
    vae = WanVideoVAE(synthetic_vae_state_dict(), device=dev)
    fr = torch.randint(0, 256, (93, 480, 832, 3), dtype=torch.uint8, device=dev)
    z = torch.randn(16, 24, 60, 104, device=dev)
    out = {}
    cases = (("decode", lambda: vae.decode(z, tiled=tiled)), ("encode", lambda: vae.encode(fr, tiled=tiled)))
    if "--decode-only" in sys.argv:   # for an ncu launch list of one decode
        cases = cases[:1]
    if "--encode-only" in sys.argv:
        cases = cases[1:]
    for name, fn in cases:
        if "--once" not in sys.argv:
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        out[name + "_s"] = time.perf_counter() - t0
    out["tiled"] = tiled
    # ~0.32 PFLOP untiled decode, x2.25 tile overlap when tiled
    if "decode_s" in out:
        out["decode_pflops_algorithmic"] = 0.32 * (2.25 if tiled else 1.0)
        out["decode_tflops"] = out["decode_pflops_algorithmic"] * 1e3 / out["decode_s"]
    print(json.dumps(out))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "vae_bench.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
