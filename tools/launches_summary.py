"""Per-kernel table of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`).

    python tools/launches_summary.py gpurun_out/s1/launches.csv [n_rows] > profiles/rNN_launches_ncu.md
"""
import csv
import gzip
import re
import sys
from collections import defaultdict

path = sys.argv[1]
n_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 32
opener = gzip.open if path.endswith(".gz") else open
with opener(path, "rt", errors="replace") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rows = list(csv.reader(lines))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
for r in rows[hdr + 1:]:
    if len(r) != len(h) or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\b(at|cusrl_b200)::", lambda m: m.group(0) if m.group(1) == "cusrl_b200" else "", r[ki])
    us = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print(f"{sum(cnt.values())} kernel launches, {total / 1e3:.1f} ms of kernel time\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for name in sorted(tot, key=lambda k: -tot[k])[:n_rows]:
    print(f"| `{name[:110]}` | {cnt[name]} | {tot[name] / 1e3:.2f} | {100 * tot[name] / total:.1f}% | {tot[name] / cnt[name]:.1f} |")
ours = sum(v for k, v in tot.items() if "cusrl_b200" in k)
print(f"\ncusrl_b200 kernels: {100 * ours / total:.1f}% of kernel time, {sum(c for k, c in cnt.items() if 'cusrl_b200' in k)} launches")
