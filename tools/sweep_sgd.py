"""Full-size sweep of the SGD kernels' concurrency knobs (diagnostic): ms/epoch and val RMSE."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n_users, n_items, nnz = int(bench.SHAPE[0] * scale), bench.SHAPE[1], int(bench.SHAPE[2] * scale)
prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
ptr, ind, val = prob["train"]
print("train nnz", int(ptr[-1]), "max user deg", int(np.diff(ptr).max()), "max item count", int(np.bincount(ind).max()), flush=True)
rng = np.random.default_rng(1)
R = bench.RANK
U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
eng = E.Engine(n_users, n_items, R)
eng.upload_csr(E.TRAIN, bench.Mat(n_users, n_items, prob["train"]), with_csc=False)
eng.upload_csr(E.VAL, bench.Mat(n_users, n_items, prob["val"]), with_csc=False)
eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))
eng.sgd_plan(1)
lr = float(os.environ.get("LR", "0.005"))

def run(kind, epochs=6, **opts):
    for k, v in dict(sgd_workers=0, sgd_warps_per_sm=16, sgd_max_hot_inflight=8, sgd_atomic=1).items():
        eng.set_option(k, opts.get(k, v))
    eng.upload_factors(U0, V0)
    ms, rm = [], []
    for ep in range(epochs):
        eng.event_record(0)
        if kind == "flat":
            eng.sgd_epoch_flat(E.MF, lr, 0.05, 0.05, 1, ep)
        else:
            eng.sgd_subepoch(np.array([[0, 0]], np.int32), E.MF, lr, 0.05, 0.05, 1, ep)
        eng.event_record(1)
        ms.append(eng.event_elapsed_ms(0, 1))
        rm.append(eng.rmse(E.VAL))
    print(f"{kind:4s} {json.dumps(opts):60s} ms/epoch {np.median(ms[1:]):8.3f}  G/s {int(ptr[-1])/np.median(ms[1:])/1e6:6.2f}  val " +
          " ".join(f"{x:.4f}" for x in rm), flush=True)

for kind in ("run", "flat"):
    run(kind)
    for hot in (2, 32, 128, 1e9):
        run(kind, sgd_max_hot_inflight=hot)
    for w in (256, 1024, 4096, 16384):
        run(kind, sgd_workers=w)
    for wps in (4, 8, 32, 64):
        run(kind, sgd_max_hot_inflight=1e9, sgd_warps_per_sm=wps)
run("run", sgd_atomic=0)
run("run", sgd_atomic=0, sgd_max_hot_inflight=1e9, sgd_warps_per_sm=64)
