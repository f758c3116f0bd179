"""A/B of the RMSNorm + RoPE pass (v1 vs the opt-in v2, ICB_RMSROPE_V2) in one process: bit-level agreement of the
outputs at the bench shape, a ragged shape with column groups off, the no-RoPE path and 14B width, then timings.
Writes gpurun_out/ab_rmsrope.json."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from infinicube_b200 import ops  # noqa: E402


def run(variant, src, ss, cnt, w, dst, tabs, f0=0):
    os.environ["ICB_RMSROPE_V2"] = str(variant)
    ops.rmsnorm_rope(src, ss, 0, cnt, w, dst, 1e-6, tabs, f0)


def main():
    dev = torch.device("cuda:0")
    res = {}
    for name, (f, hh, ww, D, rope, rows_cut) in {"bench_1p3b": (24, 30, 52, 1536, True, 0), "ragged": (3, 7, 11, 1536, True, 5),
                                                   "no_rope": (2, 30, 52, 1536, False, 3), "w14b": (2, 30, 52, 5120, True, 1)}.items():
        rows = f * hh * ww - rows_cut
        cnt = D // 256
        src = torch.randn(rows, 2 * D, device=dev).bfloat16()
        ss = (src[:, :D].float() ** 2).view(rows, cnt, 256).sum(-1).contiguous()
        w = torch.randn(D, device=dev)
        tabs = tuple(torch.randn(n, p, 2, device=dev).contiguous() for n, p in ((f + 1, 22), (hh, 21), (ww, 21))) if rope else None
        d1 = torch.zeros(rows, D, device=dev, dtype=torch.bfloat16)
        d2 = torch.zeros_like(d1)
        run(0, src[:, :D], ss, cnt, w, d1, tabs, 1)
        run(1, src[:, :D], ss, cnt, w, d2, tabs, 1)
        torch.cuda.synchronize()
        diff = (d1.float() - d2.float()).abs()
        res[name] = {"rows": rows, "D": D, "max_abs_diff": float(diff.max()), "mismatch_frac": float((diff > 0).float().mean()),
                     "ref_absmax": float(d1.float().abs().max())}
        if name in ("bench_1p3b", "w14b"):
            for v in (0, 1):
                for _ in range(3):
                    run(v, src[:, :D], ss, cnt, w, d2, tabs, 1)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    run(v, src[:, :D], ss, cnt, w, d2, tabs, 1)
                b.record()
                torch.cuda.synchronize()
                us = a.elapsed_time(b) / 20 * 1e3
                res[name][f"v{v + 1}_us"] = us
                res[name][f"v{v + 1}_gbs"] = rows * D * 4 / us / 1e3
    os.environ["ICB_RMSROPE_V2"] = "0"
    print(json.dumps(res))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "ab_rmsrope.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
