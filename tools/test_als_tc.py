"""ALS tensor-core Gram check: tcgen05 path vs fp32 CUDA-core path vs float64 (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
from gpu_driver import make_engine
from matfac_b200 import engine as E, synth

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))

splits = synth.make_splits(900, 600, 380000, seed=13, user_s=0.2, item_s=0.2)
tr = splits[0]
for rank in (128, 96, 72):
    od = ol.OracleData(*splits)
    om = ol.OracleModel(od, algo="mf", facdim=rank, maxiter=1, seed=3, nthreads=8, ureg=0.1, ireg=0.1)
    eng, _ = make_engine(splits, om, rank)
    U0, V0 = om.factors()
    res = {}
    for tc in (0, 1):
        eng.set_option("als_tensor_cores", tc)
        eng.upload_factors(U0, V0)
        eng.sync(); t0 = time.time()
        eng.als_half_step(E.USER, 0.1)
        eng.sync(); t1 = time.time()
        U1, _ = eng.download_factors()
        eng.als_half_step(E.ITEM, 0.1)
        _, V1 = eng.download_factors()
        res[tc] = (U1, V1, t1 - t0)
    Ut = np.zeros(U0.shape, np.float64)
    for u in range(tr.nrows):
        s, e = tr.rowptr[u], tr.rowptr[u + 1]
        Vs = V0[tr.rowind[s:e]].astype(np.float64)
        Ut[u] = np.linalg.solve(Vs.T @ Vs + 0.1 * np.eye(rank), Vs.T @ tr.rowval[s:e].astype(np.float64))
    om.train("als")
    Uo, Vo = om.factors()
    print(f"rank {rank}: U err vs f64: cuda-core {rel(res[0][0], Ut):.2e} tcgen05 {rel(res[1][0], Ut):.2e} oracle {rel(Uo, Ut):.2e}; "
          f"V vs oracle: cuda-core {rel(res[0][1], Vo):.2e} tcgen05 {rel(res[1][1], Vo):.2e}; nan {int(np.isnan(res[1][0]).sum())}", flush=True)
    eng.close()
