"""Emulates the N-rank DSGD bench on ONE GPU (diagnostic): the P x P blocks of the bench matrix are run one
block per launch, exactly as rank g would run block (g, sigma_t(g)); reports the emulated parallel epoch time
sum_t max_g time(g, t), the validation curve, and how both change with the hot-row concurrency cap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E, dsgd
P = int(os.environ.get("P", "8"))
epochs = int(os.environ.get("EPOCHS", "6"))
n_users, n_items, nnz = bench.SHAPE
prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
ptr, ind, val = prob["train"]
rng = np.random.default_rng(1)
R = bench.RANK
U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
eng = E.Engine(n_users, n_items, R)
eng.upload_csr(E.TRAIN, bench.Mat(n_users, n_items, prob["train"]), with_csc=False)
eng.upload_csr(E.VAL, bench.Mat(n_users, n_items, prob["val"]), with_csc=False)
eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))
prng = np.random.default_rng(7)
user_part = prng.integers(0, P, n_users).astype(np.int32)
item_part = prng.integers(0, P, n_items).astype(np.int32)
eng.sgd_plan(P, user_part, item_part)
eng.set_option("sgd_block_order", 1)
sched = dsgd.rotation_schedule(P, epochs * P)
cnt = np.bincount(ind, minlength=n_items)
print("P", P, "hot item global share", cnt.max() / cnt.sum(), flush=True)
for cap in (0.15, 0.5, 2.0, 1e9):
    eng.set_option("sgd_flat_hot_lr", cap)
    eng.upload_factors(U0, V0)
    curve, par_ms, tot_ms = [], [], []
    for ep in range(epochs):
        par = tot = 0.0
        for t in range(ep * P, (ep + 1) * P):
            worst = 0.0
            for g in range(P):
                eng.event_record(0)
                eng.sgd_subepoch(np.array([[g, sched[t, g]]], np.int32), E.MF, 0.002, 0.05, 0.05, 1, t)
                eng.event_record(1)
                ms = eng.event_elapsed_ms(0, 1)
                worst = max(worst, ms); tot += ms
            par += worst
        par_ms.append(par); tot_ms.append(tot)
        curve.append(eng.rmse(E.VAL))
    print(f"hot_lr cap {cap:8.2f}: emulated parallel ms/epoch {np.median(par_ms[1:]):7.3f} (sum of all blocks {np.median(tot_ms[1:]):7.3f}) val " +
          " ".join(f"{x:.4f}" for x in curve), flush=True)
