#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time (diagnostic): the same calls with a sync after every one,
then the overlapped step as bench.py runs it.  usage: python tools/e2e_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from matfac_b200 import engine as E  # noqa: E402

prob = bench.make_problem(bench.SHAPE, 1.0, "cuda:0")
nu, ni = prob["n_users"], prob["n_items"]
ptr, ind, val = prob["train"]
U0, V0 = bench.init_factors(nu, ni, bench.RANK)
bad_u, bad_i, _ = bench.masks_of(prob)
eng = E.Engine(nu, ni, bench.RANK)
eng.upload_csr(E.VAL, bench.Mat(nu, ni, prob["val"]), with_csc=False)
eng.set_masks(bad_u, bad_i)
h, keep = {}, []
for name, a in (("ptr", ptr), ("ind", ind), ("val", val), ("U", U0.copy()), ("V", V0.copy())):
    h[name], t = bench.pinned(a); keep.append(t)
Uo, tU = bench.pinned(np.empty_like(U0)); Vo, tV = bench.pinned(np.empty_like(V0))
trp = bench.Mat(nu, ni, (h["ptr"], h["ind"], h["val"]))
hp = bench.HP


def seg(f):
    eng.sync(); t = time.perf_counter(); f(); eng.sync(); return (time.perf_counter() - t) * 1e3


for mode in (0, 1):
    eng.set_option("copy_overlap", mode)
    for rep in range(3):
        t = {}
        t["upload_csr"] = seg(lambda: eng.upload_csr(E.TRAIN, trp, with_csc=False))
        t["upload_factors"] = seg(lambda: eng.upload_factors(h["U"], h["V"]))
        t["plan"] = seg(lambda: eng.sgd_plan(1))
        t["epoch"] = seg(lambda: eng.sgd_epoch_flat(E.MF, hp["lr"], hp["ureg"], hp["ireg"], 1, rep))
        t["eval_train"] = seg(lambda: eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True))
        t["eval_val"] = seg(lambda: eng.eval(E.VAL))
        t["download"] = seg(lambda: eng.L.mfb_download_factors(eng.h, E.CURRENT, Uo.ctypes.data, bench.RANK, Vo.ctypes.data, bench.RANK))
        print("copy_overlap", mode, "serialised segments ms:", {k: round(v, 2) for k, v in t.items()}, "sum", round(sum(t.values()), 2), flush=True)
eng.set_option("copy_overlap", 1)
for rep in range(4):
    eng.sync(); t1 = time.perf_counter()
    eng.upload_csr(E.TRAIN, trp, with_csc=False)
    eng.upload_factors(h["U"], h["V"])
    eng.sgd_plan(1)
    t2 = time.perf_counter()
    eng.sgd_epoch_flat(E.MF, hp["lr"], hp["ureg"], hp["ireg"], 1, rep)
    t3 = time.perf_counter()
    eng.L.mfb_download_factors(eng.h, E.CURRENT, Uo.ctypes.data, bench.RANK, Vo.ctypes.data, bench.RANK)
    t4 = time.perf_counter()
    eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
    t5 = time.perf_counter()
    eng.eval(E.VAL)
    t6 = time.perf_counter()
    eng.sync()
    t7 = time.perf_counter()
    print("overlapped step ms: upload+plan %.2f | epoch call %.2f | download call %.2f | eval train %.2f | eval val %.2f | sync %.2f | total %.2f"
          % tuple((b - a) * 1e3 for a, b in ((t1, t2), (t2, t3), (t3, t4), (t4, t5), (t5, t6), (t6, t7), (t1, t7))), flush=True)
eng.close()
