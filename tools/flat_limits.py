"""Where does the shuffled SGD kernel's time go?  Times the bench epoch with parts of the memory traffic switched off
(option sgd_flat_debug; results are meaningless, only the timings count)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E
n_users, n_items, nnz = bench.SHAPE
prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
ptr, ind, val = prob["train"]
rng = np.random.default_rng(1)
R = bench.RANK
U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
eng = E.Engine(n_users, n_items, R)
eng.upload_csr(E.TRAIN, bench.Mat(n_users, n_items, prob["train"]), with_csc=False)
eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))
eng.sgd_plan(1)
names = {0: "full", 1: "no U write", 2: "no V write", 3: "no writes", 4: "no u load", 8: "no v load", 7: "v load only", 11: "u load only",
         15: "rating records only", 5: "no u traffic", 10: "no v traffic"}
for wps in (32, 64):
    eng.set_option("sgd_warps_per_sm", wps)
    for dbg in (0, 1, 2, 3, 4, 8, 5, 10, 7, 11, 15):
        eng.set_option("sgd_flat_debug", dbg)
        eng.upload_factors(U0, V0)
        ms = []
        for ep in range(4):
            eng.event_record(0)
            eng.sgd_epoch_flat(E.MF, 0.002, 0.05, 0.05, 1, ep)
            eng.event_record(1)
            ms.append(eng.event_elapsed_ms(0, 1))
        print(f"warps/SM {wps} debug {dbg:2d} {names[dbg]:22s} ms/epoch {np.median(ms[1:]):7.3f}", flush=True)
