"""Times mfb_upload_csr and mfb_sgd_plan on the bench matrix (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E
n_users, n_items, nnz = bench.SHAPE
prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
ptr, ind, val = prob["train"]
eng = E.Engine(n_users, n_items, bench.RANK)
h = {}
keep = []
for name, a in (("ptr", ptr), ("ind", ind), ("val", val)):
    h[name], t = bench.pinned(a)
    keep.append(t)
    print(name, "pinned:", t.is_pinned(), flush=True)
trp = bench.Mat(n_users, n_items, (h["ptr"], h["ind"], h["val"]))
for it in range(4):
    eng.sync(); t0 = time.perf_counter()
    eng.upload_csr(E.TRAIN, trp, with_csc=False)
    t1 = time.perf_counter()
    eng.sgd_plan(1)
    eng.sync(); t2 = time.perf_counter()
    eng.sgd_epoch_flat(E.MF, 0.002, 0.05, 0.05, 1, it)
    eng.sync(); t3 = time.perf_counter()
    print(f"iter {it}: upload {1e3*(t1-t0):.2f} ms, plan {1e3*(t2-t1):.2f} ms, epoch {1e3*(t3-t2):.2f} ms", flush=True)
