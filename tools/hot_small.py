"""Diagnostic: the bench workload at 1/20 scale (tests/test_gpu_parity.py::test_sgd_netflix_shaped_rank64_matches_oracle),
validation curve of the shuffled kernel for several hot-row settings next to the oracle's serial SGD."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from matfac_b200 import engine as E, synth
import test_gpu_parity as T
from gpu_driver import make_engine

n_users, n_items, nnz = int(bench.SHAPE[0] * 0.05), bench.SHAPE[1], int(bench.SHAPE[2] * 0.05)
prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
tr = synth.Csr(n_users, n_items, *prob["train"]); va = synth.Csr(n_users, n_items, *prob["val"])
splits = (tr, va, va)
epochs = 4
LR = float(os.environ.get("LR", "0.005"))
om = T.oracle_model(splits, "mf", 64, maxiter=epochs, seed=1, learnrate=LR)
eng, variant = make_engine(splits, om, 64, with_csc=False)
U0, V0 = om.factors()
U0 = U0.copy(); V0 = V0.copy()
om.train("sgd", keep_history=True)
print("oracle                                   val " + " ".join(f"{h[3]:.4f}" for h in om.history()), flush=True)
for hot, lists, batch, stages, pace, frac in ((0, 64, 0, 8, 1, 2e-4), (1, 64, 0, 8, 1, 2e-4), (1, 64, 0, 4, 1, 2e-4), (1, 64, 1, 4, 1, 2e-4), (1, 8, 0, 8, 1, 2e-4),
                                              (1, 64, 0, 8, 0, 2e-4), (1, 127, 64, 8, 1, 1e-3), (1, 127, 0, 8, 1, 2e-4)):
    for k, v in (("sgd_hot", hot), ("sgd_hot_max_lists", lists), ("sgd_hot_batch", batch), ("sgd_hot_stages", stages), ("sgd_hot_pace", pace),
                 ("sgd_flat_inflight_frac", frac)):
        eng.set_option(k, v)
    eng.upload_factors(U0, V0)
    eng.sgd_plan(1)
    _, cold, ls = eng.debug_sgd_records(0, 0, with_records=False)
    got = []
    for ep in range(epochs):
        eng.sgd_epoch_flat(variant, LR, 0.05, 0.05, 1, ep)
        got.append(eng.rmse(E.VAL))
    st = eng.debug_sgd_hot_batch()
    print(f"hot {hot} lists {len(ls):3d} batch {batch:2d} stages {stages} pace {pace} frac {frac:g}: val " + " ".join(f"{x:.4f}" for x in got) +
          f"   mean|u|^2 {st[0] / max(st[1], 1):.3f} batch used {st[2]:.0f}", flush=True)
