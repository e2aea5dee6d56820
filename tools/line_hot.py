#!/usr/bin/env python
"""Warp-stall samples per SOURCE LINE for one kernel of an .ncu-rep: joins the SASS-level source page of ncu
with the line table nvdisasm prints for the same function of the built library (instruction order is identical).
usage: python tools/line_hot.py <report.ncu-rep> <kernel-name-substring> [launch index in report] [top N]"""
import csv, io, os, re, subprocess, sys, tempfile, glob
rep, needle = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "matfac_b200", "libmfb.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines_of = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    cur, fn, table = None, None, {}
    for l in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            fn = m.group(1); table[fn] = []; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if fn and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
            table[fn].append(cur)
    for fn, t in table.items():
        if needle in fn and t:
            lines_of = (fn, t)
if not lines_of:
    sys.exit("function not found in the library: " + needle)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
blocks = [b for b in out.split('"Kernel Name"')[1:] if needle.split("ILi")[0][-12:] in b.splitlines()[0] or True]
blk = blocks[which]
rd = list(csv.reader(io.StringIO("\n".join(blk.splitlines()[1:]))))
hdr = rd[0]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
rows = [r for r in rd[1:] if len(r) == len(hdr)]
fn, table = lines_of
print("kernel:", blk.splitlines()[0][:120], "| function:", fn, "| sass", len(rows), "line-table", len(table))
agg = {}
tot = 0
for i, r in enumerate(rows):
    n = int(r[si] or 0); tot += n
    key = table[i] if i < len(table) else ("?", 0)
    a = agg.setdefault(key, [0, {}])
    a[0] += n
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: a[1][hdr[c][6:]] = a[1].get(hdr[c][6:], 0) + v
src_cache = {}
def src(f, ln):
    for root in (os.path.join(ROOT, "matfac_b200", "csrc"),):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln - 1].strip()[:80] if 0 < ln <= len(L) else ""
    return ""
for (f, ln), (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ", ".join(f"{k} {100*v/max(n,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"{100*n/max(tot,1):5.1f}%  {f}:{ln:<5d} [{top}]  {src(f, ln)}")
