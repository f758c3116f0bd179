"""Summarises an ncu launch list of tools/gemm_shard_probe.py: median duration per (M, shape)."""
import collections
import csv
import json
import statistics
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ix = {n: H.index(n) for n in ("ID", "Metric Name", "Metric Value")}
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) >= len(H):
        d.setdefault(int(r[ix["ID"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
shapes = [(M, n) for M in (37440, 18720, 9472, 9360, 9216) for n in ("qkv", "o_proj", "cross_q", "ffn1", "ffn2")]
ids = list(d)
out = collections.OrderedDict()
for si, (M, n) in enumerate(shapes):
    ls = [d[ids[si * reps + i]] for i in range(1, reps)]
    out[f"M{M}_{n}"] = {"us": round(statistics.median(l["gpu__time_duration.sum"] for l in ls) / 1e3, 1),
                        "sm_active_kclk_avg": round(ls[-1]["sm__cycles_active.avg"] / 1e3, 1),
                        "sm_active_kclk_max": round(ls[-1]["sm__cycles_active.max"] / 1e3, 1)}
for M in (37440, 18720, 9472, 9360, 9216):
    g = lambda n: out[f"M{M}_{n}"]["us"]
    out[f"M{M}_per_layer_us"] = round(g("qkv") + 2 * g("o_proj") + g("cross_q") + g("ffn1") + g("ffn2"), 1)
print(json.dumps(out, indent=1))
