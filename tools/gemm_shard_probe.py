"""Runs the DiT block's GEMM shapes at several row counts a few times each, for an ncu launch list
(`ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max -k regex:gemm`)."""
import sys

import torch

sys.path.insert(0, ".")
from infinicube_b200 import ops  # noqa: E402

torch.manual_seed(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for M in (37440, 18720, 9472, 9360, 9216):
    for name, N, K, resid in (("qkv", 4608, 1536, False), ("o_proj", 1536, 1536, True), ("cross_q", 1536, 1536, False),
                              ("ffn1", 8960, 1536, False), ("ffn2", 1536, 8960, True)):
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        b = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        bias = torch.randn(N, device="cuda")
        gate = torch.randn(N, device="cuda") * 0.1
        x = torch.randn(M, N, device="cuda") if resid else None
        o = None if resid else torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(reps):
            if i == 1:
                s.record()
            if resid:
                ops.gemm(a, b, bias=bias, resid=x, gate=gate)
            else:
                ops.gemm(a, b, bias=bias, out_bf16=o)
        e.record()
        torch.cuda.synchronize()
        print(f"SHAPE M={M} {name} N={N} K={K} event_us={s.elapsed_time(e) / (reps - 1) * 1e3:.1f}", flush=True)
