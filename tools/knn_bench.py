"""Times the nearest-neighbour label transfer at a stage-1 chunk-merge size: ~2 M voxel centres
(0.2 m lattice, rigidly moved) against ~2 M.  Writes gpurun_out/knn_bench.json."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from infinicube_b200.voxelgen.utils.color_util import KnnIndex  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    ijk = torch.unique(torch.randint(0, 400, (2_400_000, 3), generator=g) * torch.tensor([1, 1, 0]) +
                       torch.randint(0, 40, (2_400_000, 3), generator=g) * torch.tensor([0, 0, 1]), dim=0)
    ref = (ijk.float() * 0.2 + 0.1).to(dev)
    m = ref.shape[0]
    sem = (torch.arange(m, device=dev) % 19).long()
    c, s = np.cos(0.03), np.sin(0.03)
    R = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=torch.float32, device=dev)
    q = (ref @ R.T + torch.tensor([0.07, -0.04, 0.02], device=dev)).contiguous()
    res = {"ref_points": m, "queries": q.shape[0]}
    for cell in (0.2, 0.4, 0.0):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        KnnIndex(ref, cell).query(q, sem)          # warm-up
        torch.cuda.synchronize()
        e[0].record()
        index = KnnIndex(ref, cell)
        e[1].record()
        for _ in range(5):
            index.query(q, sem, want_dist=False, want_idx=False)
        e[2].record()
        torch.cuda.synchronize()
        info = index.info()
        q_ms = e[1].elapsed_time(e[2]) / 5
        res[f"cell_{cell}"] = {"build_ms": e[0].elapsed_time(e[1]), "query_ms": q_ms, "cells": info["cells"],
                               "cell_size": info["cell_size"], "mqueries_per_s": q.shape[0] / q_ms / 1e3,
                               # queries 12 B in + 8 B label out + one 16 B record and one 8 B label of the hit
                               "algorithmic_gbs": q.shape[0] * 44 / (q_ms * 1e-3) / 1e9}
    print(json.dumps(res))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "knn_bench.json").write_text(json.dumps(res))


if __name__ == "__main__":
    main()
