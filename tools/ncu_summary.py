#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py <report.ncu-rep> [regex of extra metric names]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = [r"^Kernel Name$", r"^Block Size$", r"^Grid Size$", r"gpu__time_duration\.sum$", r"dram__bytes_read\.sum$", r"dram__bytes_write\.sum$",
        r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"lts__t_bytes\.sum$", r"lts__t_sector_hit_rate\.pct$",
        r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"launch__registers_per_thread$", r"launch__occupancy_limit", r"sm__pipe_tensor.*cycles_active.*pct", r"sm__inst_executed_pipe_tensor.*sum$",
        r"smsp__average_warp.*issue_stalled.*_per_warp_active\.pct$", r"smsp__warp_issue_stalled.*per_warp_active\.pct$",
        r"sm__inst_executed_pipe_(fma|alu|lsu|fp64).*pct", r"smsp__inst_executed\.sum$", r"launch__shared_mem_per_block", r"launch__waves_per_multiprocessor"]
if extra:
    want.append(extra)
for r in rows[2:]:
    print("=" * 100)
    seen = set()
    for pat in want:
        for i, h in enumerate(hdr):
            base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            if re.search(pat, h) and base not in seen and "Triage" not in h:
                seen.add(base)
                v = r[i]
                if v in ("", "n/a"):
                    continue
                print(f"{h:90s} {v:>22s} {units[i]}")
