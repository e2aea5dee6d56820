"""Full-scale convergence vs concurrency cap (diagnostic)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E
R = 64
scale = float(os.environ.get("SCALE", "1.0"))
lr = float(os.environ.get("LR", "0.002"))
EPOCHS = int(os.environ.get("EPOCHS", "8"))
n_users, n_items, nnz = int(bench.SHAPE[0] * scale), bench.SHAPE[1], int(bench.SHAPE[2] * scale)
prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
ptr, ind, val = prob["train"]
print("scale", scale, "lr", lr, "nnz", int(ptr[-1]), "max user", int(np.diff(ptr).max()), "max item", int(np.bincount(ind).max()), flush=True)
rng = np.random.default_rng(1)
U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
eng = E.Engine(n_users, n_items, R)
eng.upload_csr(E.TRAIN, bench.Mat(n_users, n_items, prob["train"]), with_csc=False)
eng.upload_csr(E.VAL, bench.Mat(n_users, n_items, prob["val"]), with_csc=False)
eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))

def run(kind, P=1, **opts):
    for k, v in dict(sgd_workers=0, sgd_warps_per_sm=32, sgd_max_hot_inflight=8, sgd_atomic=1).items():
        eng.set_option(k, opts.get(k, v))
    eng.upload_factors(U0, V0)
    ms, rm = [], []
    prng = np.random.default_rng(5)
    for ep in range(EPOCHS):
        eng.event_record(0)
        if kind == "flat":
            eng.sgd_epoch_flat(E.MF, lr, 0.05, 0.05, 1, ep)
        elif P == 1:
            eng.sgd_subepoch(np.array([[0, 0]], np.int32), E.MF, lr, 0.05, 0.05, 1, ep)
        else:
            for k in range(P):
                perm = prng.permutation(P)
                blocks = np.stack([prng.permutation(P), perm], 1).astype(np.int32)
                eng.sgd_subepoch(blocks, E.MF, lr, 0.05, 0.05, 1, ep * P + k)
        eng.event_record(1)
        ms.append(eng.event_elapsed_ms(0, 1))
        rm.append(eng.rmse(E.VAL))
    print(f"{kind:4s} P={P} {json.dumps(opts):45s} ms {np.median(ms[1:]):7.2f} G/s {int(ptr[-1])/np.median(ms[1:])/1e6:5.2f} val " +
          " ".join(f"{x:.4f}" for x in rm), flush=True)

eng.sgd_plan(1)
for cap in (1, 8, 32, 128, 1e9):
    run("flat", sgd_max_hot_inflight=cap)
run("run", sgd_max_hot_inflight=1e9)
prng = np.random.default_rng(7)
P = 8
eng.sgd_plan(P, prng.integers(0, P, n_users).astype(np.int32), prng.integers(0, P, n_items).astype(np.int32))
for cap in (8, 32, 128, 1e9):
    run("run", P=P, sgd_max_hot_inflight=cap)
