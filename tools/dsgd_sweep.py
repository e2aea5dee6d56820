"""One-GPU emulation of the 8-rank DSGD bench (every (g, sigma_t(g)) block of the balanced partition as its own
launch): emulated parallel epoch time and validation curve against the in-flight bound and the hot-row CTAs.
The engine here holds all 8 user strata, so its in-flight budget is sgd_flat_inflight_frac x 100 M; a real rank
holds one stratum (12.5 M): FRAC below is the per-rank fraction, the option is set to FRAC / 8."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E, dsgd

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
LR = 0.002
epochs = int(os.environ.get("EPOCHS", "10"))
P = 8
user_part = dsgd.balanced_partition(np.diff(ptr), P)
item_part = dsgd.balanced_partition(np.bincount(ind, minlength=n_items), P)
sched = dsgd.rotation_schedule(P, epochs * P)
for hot, frac in ((0, 2e-4), (0, 8e-4), (0, 3.2e-3), (1, 2e-4), (1, 8e-4), (1, 3.2e-3), (1, 1.6e-2)):
    eng.set_option("sgd_hot", hot)
    eng.set_option("sgd_flat_inflight_frac", frac / 8)
    eng.sgd_plan(P, user_part, item_part)
    eng.set_option("sgd_block_order", 1)
    nl = len(eng.debug_sgd_records(0, 0, with_records=False)[2])
    eng.upload_factors(U0, V0)
    curve, par_ms = [], []
    for ep in range(epochs):
        par = 0.0
        for t in range(ep * P, (ep + 1) * P):
            worst = 0.0
            for g in range(P):
                eng.event_record(0)
                eng.sgd_subepoch(np.array([[g, sched[t, g]]], np.int32), E.MF, LR, 0.05, 0.05, 1, t)
                eng.event_record(1)
                worst = max(worst, eng.event_elapsed_ms(0, 1))
            par += worst
        par_ms.append(par)
        curve.append(eng.rmse(E.VAL))
    st = eng.debug_sgd_hot_batch()
    print(f"hot {hot} per-rank frac {frac:g}: lists/block {nl:3d} batch {st[2]:.0f}  emulated parallel ms/epoch {np.median(par_ms[1:]):7.3f}  val " +
          " ".join(f"{x:.4f}" for x in curve), flush=True)
eng.close()
