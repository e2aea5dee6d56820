#!/usr/bin/env python
"""DRAM traffic of one SGD epoch from an ncu launch list (read here, no GPU needed):
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:"sgd_flat|sgd_hot" --csv --log-file gpurun_out/<tag>.csv python tools/ncu_target.py --shape netflix|yahoo --algo sgd
usage: python tools/ncu_traffic.py <key> <csv> [<key> <csv> ...]   ->  profiles/r2_ncu_traffic.json (merged)
The last launch of every kernel name in the list is one epoch's launch of that kernel (the target runs whole epochs)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
for key, path in zip(sys.argv[1::2], sys.argv[2::2]):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iid, ik, im, iu, iv = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = {}
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        d = launches.setdefault(int(r[iid]), {"kernel": r[ik]})
        d[r[im]] = float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
    last = {}
    for i in sorted(launches):
        name = launches[i]["kernel"].split("(")[0]
        last[name] = launches[i]
    kernels = []
    dram = ms = 0.0
    for name, d in last.items():
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        dram += b
        ms += d.get("gpu__time_duration.sum", 0.0)
        kernels.append({"kernel": name, "dram_bytes": b, "ms_under_ncu": d.get("gpu__time_duration.sum", 0.0)})
    out[key] = {"dram_bytes": dram, "kernels": kernels, "ms_under_ncu_serialised": ms,
                "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, last launch of each kernel in {os.path.basename(path)} (profiles/)"}
    print(key, f"{dram/1e9:.2f} GB per epoch over", [k["kernel"][:40] for k in kernels])
json.dump(out, open(out_path, "w"), indent=1)
