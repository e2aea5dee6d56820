"""Shuffled SGD kernel on the bench matrix with a given user-band size (env BAND_MB): ms/epoch and val RMSE curve."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E
mb = float(os.environ.get("BAND_MB", "32"))
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
ustore = int(os.environ.get("USTORE", "0"))
for wps in (32, 64):
    eng.set_option("sgd_flat_user_store", ustore)
    eng.set_option("sgd_flat_band_mb", mb)
    eng.set_option("sgd_warps_per_sm", wps)
    eng.sgd_plan(1)
    eng.upload_factors(U0, V0)
    ms, rm = [], []
    for ep in range(8):
        eng.event_record(0)
        eng.sgd_epoch_flat(E.MF, 0.002, 0.05, 0.05, 1, ep)
        eng.event_record(1)
        ms.append(eng.event_elapsed_ms(0, 1))
        rm.append(eng.rmse(E.VAL))
    print(f"ustore {ustore} band_mb {mb:5.0f} warps/SM {wps} ms/epoch {np.median(ms[1:]):7.3f} G/s {int(ptr[-1])/np.median(ms[1:])/1e6:6.2f} val " + " ".join(f"{x:.4f}" for x in rm), flush=True)
