"""Aggregates an ncu `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list into
per-kernel totals:  python tools/launch_summary.py launches.csv [top_n [last_n_launches]]"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 6]
h = rows[0]
ki, vi, ui, mi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Metric Name"), h.index("ID")
T = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}
Bm = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
MERGE = len(sys.argv) > 4 and sys.argv[4] == "merge"      # gemm_tc_kernel<160, 4, 2, 0> -> gemm_tc_kernel<...>
launch = {}                                 # ncu ID -> [name, ms, dram MB]
for r in rows[1:]:
    name = re.sub(r"^void ", "", re.sub(r"\(.*$", "", r[ki]))
    if MERGE:
        name = re.sub(r"<[^<>]*>$", "<...>", name)
    e = launch.setdefault(int(r[ii]), [name[:64], 0.0, 0.0])
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        e[1] += v * T.get(r[ui], 1e-6)
    elif r[mi].startswith("dram__bytes"):
        e[2] += v * Bm.get(r[ui], 1e-6)
ids = sorted(launch)
if len(sys.argv) > 3:                      # keep only the last N launches (e.g. the second, warm forward of two)
    ids = ids[-int(sys.argv[3]):]
tot, cnt, mb = defaultdict(float), defaultdict(int), defaultdict(float)
for i in ids:
    n, ms, d = launch[i]
    tot[n] += ms
    cnt[n] += 1
    mb[n] += d
s = sum(tot.values())
print(f"{'ms':>10s} {'share':>6s} {'launches':>8s} {'dram MB':>9s}  kernel   (total {s:.3f} ms, {sum(cnt.values())} launches)")
for k in sorted(tot, key=tot.get, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{tot[k]:10.3f} {100 * tot[k] / s:5.1f}% {cnt[k]:8d} {mb[k]:9.1f}  {k}")
