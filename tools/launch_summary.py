#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and share."""
import csv, re, sys
from collections import OrderedDict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
# optional second argument "engine": keep only the engine's own launches (its kernels and the CUB sorts / scans of
# its plan builders; the torch kernels of bench.py's synthetic-data generator are dropped)
engine_only = len(sys.argv) > 2 and sys.argv[2] == "engine"
for r in rd:
    if len(r) <= vi: continue
    if engine_only and re.search(r"^(void )?(at::|native::|at_cuda_detail::|cuda::|c10::|compute_cuda_kernel)", r[ki]): continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mfb::", "")
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':90s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:90s} {a[0]:8d} {a[1]:10.3f} {100*a[1]/tot:6.1f}%")
print(f"{'TOTAL':90s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")
