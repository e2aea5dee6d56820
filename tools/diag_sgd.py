"""GPU-vs-oracle RMSE curves for the SGD trainers (diagnostic; prints a table)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
from common import small_problem
from gpu_driver import make_engine, run_sgd
from matfac_b200 import engine as E, synth

def curve(splits, algo, method, P, rank, epochs, lr=0.005, reg=0.05, kernel="run", opts=None, **flags):
    od = ol.OracleData(*splits)
    om = ol.OracleModel(od, algo=algo, facdim=rank, maxiter=epochs, seed=3, nthreads=P, ureg=reg, ireg=reg, learnrate=lr, **flags)
    eng, variant = make_engine(splits, om, rank, algo, rho=flags.get("rhorms", 0.0), with_csc=False)
    for k, v in (opts or {}).items():
        eng.set_option(k, v)
    up, ip, sched = om.dsgd_plan(P, epochs * P)
    eng.sgd_plan(P, up, ip) if P > 1 else eng.sgd_plan(1)
    g = []
    for ep in range(epochs):
        if kernel == "flat":
            eng.sgd_epoch_flat(variant, lr, reg, reg, 3, ep)
        elif P == 1:
            eng.sgd_subepoch(np.array([[0, 0]], np.int32), variant, lr, reg, reg, 3, ep)
        else:
            for k in range(P):
                eng.sgd_subepoch(sched[ep * P + k], variant, lr, reg, reg, 3, ep * P + k)
        g.append(eng.rmse(E.VAL, E.CURRENT, variant))
    om.train(method, keep_history=True)
    o = [h[3] for h in om.history()]
    eng.close()
    sel = [e for e in range(epochs) if e < 3 or e % 5 == 4 or e == epochs - 1]
    print(f"{algo:10s} {method:7s} {kernel:4s} P={P} r={rank:3d} nnz={splits[0].nnz} {opts or ''} lr={om.learn_rate:.4f}")
    print("   ep     " + " ".join(f"{e:8d}" for e in sel))
    print("   gpu    " + " ".join(f"{g[e]:8.4f}" for e in sel))
    print("   oracle " + " ".join(f"{o[e]:8.4f}" for e in sel))
    print("   rel%   " + " ".join(f"{100*(g[e]-o[e])/o[e]:+8.2f}" for e in sel), flush=True)

if __name__ == "__main__":
    s1 = small_problem(3000, 1500, 300000, seed=21)
    ml1m = synth.make_splits(6040, 3706, 1000209, seed=20260101)
    for sp in (s1, ml1m):
        curve(sp, "mf", "sgd", 1, 10, 30, kernel="flat")
        curve(sp, "mf", "sgd", 1, 10, 30, kernel="flat", opts=dict(sgd_max_hot_inflight=2))
        curve(sp, "mf", "sgd", 1, 10, 30, kernel="flat", opts=dict(sgd_max_hot_inflight=32))
        curve(sp, "mf", "sgd", 1, 10, 30, kernel="run")
        curve(sp, "mf", "sgdpar", 8, 10, 30)
        curve(sp, "mf", "sgdpar", 8, 10, 30, opts=dict(sgd_max_hot_inflight=2))
        curve(sp, "mf", "sgdpar", 8, 10, 30, opts=dict(sgd_max_hot_inflight=32))
        curve(sp, "mf", "sgdpar", 8, 10, 30, opts=dict(sgd_atomic=0))
        curve(sp, "mf", "sgdpar", 8, 64, 20)
        curve(sp, "IFWMF", "sgd", 1, 10, 20, kernel="flat", rhorms=100.0)
        curve(sp, "IFWMF", "sgdpar", 4, 10, 20, rhorms=100.0)
        curve(sp, "TMF", "sgdpar", 8, 16, 20, rhorms=20.0, alpha=0.5)
        curve(sp, "TMFDropout", "sgdpar", 8, 16, 20, rhorms=20.0, alpha=0.5)
