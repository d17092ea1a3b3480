"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel totals."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 6]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
data = [r for r in rows[1:] if r[h.index("Metric Name")] == "gpu__time_duration.sum"]
if len(sys.argv) > 3:                      # keep only the last N launches (e.g. the second, warm forward of two)
    data = data[-int(sys.argv[3]):]
for r in data:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*$", "", r[ki])
    name = re.sub(r"^void ", "", name)[:64]
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"{'ms':>10s} {'share':>6s} {'launches':>8s}  kernel   (total {s:.3f} ms, {sum(cnt.values())} launches)")
for k in sorted(tot, key=tot.get, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{tot[k]:10.3f} {100 * tot[k] / s:5.1f}% {cnt[k]:8d}  {k}")
