import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E
R = 64
for scale in (0.05, 0.25, 1.0):
    n_users, n_items, nnz = int(bench.SHAPE[0] * scale), bench.SHAPE[1], int(bench.SHAPE[2] * scale)
    prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
    ptr, ind, val = prob["train"]
    print("scale", scale, "nnz", int(ptr[-1]), "max user", int(np.diff(ptr).max()), "max item", int(np.bincount(ind).max()), flush=True)
    rng = np.random.default_rng(1)
    U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
    V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
    eng = E.Engine(n_users, n_items, R)
    eng.upload_csr(E.TRAIN, bench.Mat(n_users, n_items, prob["train"]), with_csc=False)
    eng.upload_csr(E.VAL, bench.Mat(n_users, n_items, prob["val"]), with_csc=False)
    eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))
    eng.sgd_plan(1)
    print("  init val rmse", eng.rmse(E.VAL), "train", eng.rmse(E.TRAIN), flush=True)
    for kind, opts, lr in (("flat", {}, 0.005), ("flat", {"sgd_max_hot_inflight": 1}, 0.005), ("flat", {}, 0.001), ("run", {}, 0.005), ("run", {}, 0.001)):
        for k, v in dict(sgd_workers=0, sgd_warps_per_sm=16, sgd_max_hot_inflight=8, sgd_atomic=1).items():
            eng.set_option(k, opts.get(k, v))
        eng.upload_factors(U0, V0)
        out = []
        for ep in range(4):
            if kind == "flat":
                eng.sgd_epoch_flat(E.MF, lr, 0.05, 0.05, 1, ep)
            else:
                eng.sgd_subepoch(np.array([[0, 0]], np.int32), E.MF, lr, 0.05, 0.05, 1, ep)
            U, V = eng.download_factors()
            out.append((eng.rmse(E.VAL), int(np.isnan(U).sum()), int(np.isnan(V).sum()), float(np.nanmax(np.abs(U))), float(np.nanmax(np.abs(V)))))
        print("  ", kind, opts, lr, " | ".join(f"{o[0]:.4f} nanU={o[1]} nanV={o[2]} maxU={o[3]:.2f} maxV={o[4]:.2f}" for o in out), flush=True)
    eng.close()
