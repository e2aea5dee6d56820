#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i rep --page source --csv` (first kernel of the report,
or the N-th with argv[2]); prints the dominant stall reason per instruction and a coarse histogram over the program."""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
blocks = out.split('"Kernel Name"')
blk = '"Kernel Name"' + blocks[1 + which]
lines = blk.splitlines()
print(lines[0][:150])
rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rd[0]
si = hdr.index("# Samples"); src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
rows = [r for r in rd[1:] if len(r) == len(hdr)]
tot = sum(int(r[si] or 0) for r in rows)
print("instructions", len(rows), "samples", tot)
stot = {hdr[i]: sum(int(r[i] or 0) for r in rows) for i in stall_cols}
print("stall totals:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.1f}%" for k, v in sorted(stot.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][si] or 0))[:35]
for i in sorted(order):
    r = rows[i]
    top = max(stall_cols, key=lambda c: int(r[c] or 0))
    print(f"{i:5d} {100*int(r[si])/max(tot,1):5.1f}% {hdr[top][6:]:12s} {r[src].strip()[:90]}")
# histogram over 20 program segments
n = len(rows); seg = max(1, n // 20)
print("segment histogram (instruction index range: share of samples)")
for s in range(0, n, seg):
    sh = sum(int(r[si] or 0) for r in rows[s:s + seg])
    print(f"  {s:5d}-{min(s+seg,n)-1:5d}: {100*sh/max(tot,1):5.1f}%")
