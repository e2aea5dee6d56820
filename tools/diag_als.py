"""ALS half-step diagnostics: device vs oracle vs float64 solve, per row."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
from common import small_problem
from gpu_driver import make_engine
from matfac_b200 import engine as E

for rank in (10, 64, 128):
    splits = small_problem(500, 300, 40000, seed=13)
    tr = splits[0]
    od = ol.OracleData(*splits)
    om = ol.OracleModel(od, algo="mf", facdim=rank, maxiter=1, seed=3, nthreads=4, ureg=0.1, ireg=0.1)
    eng, _ = make_engine(splits, om, rank)
    U0, V0 = om.factors()
    eng.als_half_step(E.USER, 0.1)
    U1, _ = eng.download_factors()
    # float64 truth for the user half-step
    Ut = np.zeros_like(U0, dtype=np.float64)
    for u in range(tr.nrows):
        s, e = tr.rowptr[u], tr.rowptr[u + 1]
        Vs = V0[tr.rowind[s:e]].astype(np.float64)
        r = tr.rowval[s:e].astype(np.float64)
        A = Vs.T @ Vs + 0.1 * np.eye(rank)
        Ut[u] = np.linalg.solve(A, Vs.T @ r)
    om.train("als")
    Uo, Vo = om.factors()
    eng.als_half_step(E.ITEM, 0.1)
    _, V1 = eng.download_factors()
    def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
    rowerr = np.linalg.norm(U1 - Ut, axis=1) / np.maximum(np.linalg.norm(Ut, axis=1), 1e-30)
    rowerr_o = np.linalg.norm(Uo - Ut, axis=1) / np.maximum(np.linalg.norm(Ut, axis=1), 1e-30)
    worst = np.argsort(-rowerr)[:5]
    print(f"rank {rank}: U gpu-vs-f64 {rel(U1, Ut):.2e}  oracle-vs-f64 {rel(Uo, Ut):.2e}  gpu-vs-oracle {rel(U1, Uo):.2e}  V gpu-vs-oracle {rel(V1, Vo):.2e}")
    print("   worst gpu rows", [(int(u), int(tr.rowptr[u+1]-tr.rowptr[u]), float(rowerr[u]), float(rowerr_o[u])) for u in worst])
    print("   median row err gpu", float(np.median(rowerr)), "oracle", float(np.median(rowerr_o)), flush=True)
    eng.close()
