"""Per-kernel shares from an ncu `--metrics gpu__time_duration.sum --csv` launch list:
   python tools/launch_shares.py <launches.csv> "<command that produced it>" [exclude-regex] > shares.json
(the exclude pattern drops set-up kernels, e.g. the weight loader's copy_convert_kernel, from the total)"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    name = re.sub(r"\(.*$", "", r[kn]).strip()
    if len(sys.argv) > 3 and re.search(sys.argv[3], name):
        continue
    v = float(r[mv].replace(",", ""))
    v_ms = v / 1e6 if r[mu] in ("ns", "nsecond") else (v / 1e3 if r[mu] in ("us", "usecond") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v_ms
total = sum(a[1] for a in agg.values())
out = {"source": sys.argv[2] if len(sys.argv) > 2 else "", "total_ms": total,
       "note": "cold-cache, serialised per-launch times: compare SHARES with bench.py's live kernel_shares, not absolutes",
       "kernels": [{"kernel": k, "launches": a[0], "total_ms": round(a[1], 3), "share": round(a[1] / total, 4),
                    "avg_us": round(a[1] / a[0] * 1e3, 2)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
