import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
from gpu_driver import make_engine
from matfac_b200 import engine as E, synth
np.set_printoptions(linewidth=200, precision=4, suppress=True)
splits = synth.make_splits(200, 150, 12000, seed=13, user_s=0.2, item_s=0.2)
tr = splits[0]
rank = 128
od = ol.OracleData(*splits)
om = ol.OracleModel(od, algo="mf", facdim=rank, maxiter=1, seed=3, nthreads=2, ureg=0.1, ireg=0.1)
eng, _ = make_engine(splits, om, rank)
# structured factors: V[i, j] = (i+1) * 1e-3 + j  -> easy to recognise rows/cols
V = (np.arange(od.n_items)[:, None] * 0.01 + np.arange(rank)[None, :] * 1.0 + 1.0).astype(np.float32)
U = np.zeros((od.n_users, rank), np.float32)
eng.upload_factors(U, V)
row = 5
s, e = tr.rowptr[row], tr.rowptr[row + 1]
Vs = V[tr.rowind[s:e]].astype(np.float64)
G = Vs.T @ Vs
b = Vs.T @ tr.rowval[s:e].astype(np.float64)
for tc in (0, 1):
    eng.set_option("als_tensor_cores", tc)
    Gd, bd = eng.debug_als_gram(E.USER, row)
    print("tc", tc, "n ratings", e - s, "G rel err", np.linalg.norm(Gd - G) / np.linalg.norm(G), "b rel err", np.linalg.norm(bd - b) / np.linalg.norm(b))
    if tc:
        print("true G[0:4,0:6]\n", G[0:4, 0:6]); print("dev  G[0:4,0:6]\n", Gd[0:4, 0:6])
        print("true G[30:34,30:36]\n", G[30:34, 30:36]); print("dev\n", Gd[30:34, 30:36])
        # block structure: per 32x32 block relative error
        for bi in range(4):
            print([f"{np.linalg.norm(Gd[bi*32:(bi+1)*32, bj*32:(bj+1)*32]-G[bi*32:(bi+1)*32, bj*32:(bj+1)*32])/np.linalg.norm(G[bi*32:(bi+1)*32, bj*32:(bj+1)*32]):.3f}" for bj in range(4)])
        # try to find where dev G[0, :] values appear in true G
        print("dev row0 first 16", Gd[0, :16]); print("true row0 first 16", G[0, :16])
        print("dev col0 first 16", Gd[:16, 0]); print("ratio dev/true diag", (np.diag(Gd) / np.diag(G))[:16])
# random factors: symmetric? 
rng = np.random.default_rng(0)
V = rng.normal(size=(od.n_items, rank)).astype(np.float32)
eng.upload_factors(U, V)
Vs = V[tr.rowind[s:e]].astype(np.float64); G = Vs.T @ Vs
eng.set_option("als_tensor_cores", 1)
Gd, bd = eng.debug_als_gram(E.USER, row)
print("random: rel err", np.linalg.norm(Gd - G) / np.linalg.norm(G), "symmetry", np.linalg.norm(Gd - Gd.T) / np.linalg.norm(Gd))
# does Gd match G under a permutation of indices? test candidate: units swizzle not undone etc.
best = []
for name, perm in (("identity", np.arange(128)),):
    pass
# correlation of each dev row with each true row
C = np.abs(np.corrcoef(np.vstack([Gd, G]))[:128, 128:])
print("argmax true-row for dev rows 0..31:", C.argmax(1)[:32])
print("max corr", C.max(1)[:8])
