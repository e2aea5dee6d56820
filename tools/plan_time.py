"""Times the phases of the bench's end-to-end step on the bench matrix with a synchronisation after each phase
(diagnostic): upload CSR + factors, plan, epoch, objective + validation pass, factor download."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E
n_users, n_items, nnz = bench.SHAPE
prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
ptr, ind, val = prob["train"]
R = bench.RANK
rng = np.random.default_rng(1)
U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
eng = E.Engine(n_users, n_items, R)
eng.upload_csr(E.VAL, bench.Mat(n_users, n_items, prob["val"]), with_csc=False)
eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))
h = {}
keep = []
for name, a in (("ptr", ptr), ("ind", ind), ("val", val), ("U", U0), ("V", V0)):
    h[name], t = bench.pinned(a)
    keep.append(t)
Uo, tU = bench.pinned(np.empty_like(U0)); Vo, tV = bench.pinned(np.empty_like(V0))
trp = bench.Mat(n_users, n_items, (h["ptr"], h["ind"], h["val"]))
for hot in [int(x) for x in os.environ.get("HOT", "1,0").split(",")]:
    eng.set_option("sgd_hot", hot)
    for it in range(int(os.environ.get("ITERS", "4"))):
        eng.sync(); t0 = time.perf_counter()
        eng.upload_csr(E.TRAIN, trp, with_csc=False)
        eng.upload_factors(h["U"], h["V"])
        eng.sync(); t1 = time.perf_counter()
        eng.sgd_plan(1)
        eng.sync(); t2 = time.perf_counter()
        eng.sgd_epoch_flat(E.MF, 0.002, 0.05, 0.05, 1, it)
        eng.sync(); t3 = time.perf_counter()
        obj = eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
        vr = eng.eval(E.VAL)
        eng.sync(); t4 = time.perf_counter()
        eng.L.mfb_download_factors(eng.h, E.CURRENT, Uo.ctypes.data, R, Vo.ctypes.data, R)
        eng.sync(); t5 = time.perf_counter()
        print(f"hot {hot} iter {it}: upload {1e3*(t1-t0):.2f} ms, plan {1e3*(t2-t1):.2f} ms, epoch {1e3*(t3-t2):.2f} ms, eval {1e3*(t4-t3):.2f} ms, "
              f"download {1e3*(t5-t4):.2f} ms, total {1e3*(t5-t0):.2f} ms", flush=True)
