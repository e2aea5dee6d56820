"""BASELINE.json configs[3]: IFWMF weighted SGD, TMF and TMF+Dropout (modelPoissonDropout) on an ML-20M-shaped synthetic
matrix (138 493 x 26 744, 20 M ratings), rank 64, one B200 — epoch time and rating updates/s of the shuffled SGD path
per model variant, with the mean effective rank the truncated models update (algorithmic bytes per update are
16 k + 12 with k the mean effective rank, + 8 for the two IFWMF weights).

The per-id inputs are computed here with numpy from the reference's formulas (modelInvPopMF.cpp:98-114,
modelDropoutSigmoid.h:67-95 / .cpp:158-170, modelPoissonDropout.cpp:25-47): this tool times kernels, the parity of
those inputs is the business of tests/ (oracle) and matfac_b200/host (bit-exact host code)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from matfac_b200 import engine as E, synth
from gpu_driver import poisson_cdf_table

n_users, n_items, nnz = synth.SHAPES["ml20m"]
R = 64
prob = bench.gen_problem(n_users, n_items, nnz, 20260104, "cuda:0")
ptr, ind, val = prob["train"]
train_nnz = int(ptr[-1])
rng = np.random.default_rng(1)
U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
eng = E.Engine(n_users, n_items, R)
FRAC = float(os.environ.get("INFLIGHT_FRAC", "2e-4"))  # sgd_flat_inflight_frac (default 2e-4)
eng.set_option("sgd_flat_inflight_frac", FRAC)
eng.upload_csr(E.TRAIN, bench.Mat(n_users, n_items, prob["train"]), with_csc=False)
eng.upload_csr(E.VAL, bench.Mat(n_users, n_items, prob["val"]), with_csc=False)
ufreq = np.diff(ptr).astype(np.int32)
ifreq = np.bincount(ind, minlength=n_items).astype(np.int32)
eng.set_masks((ufreq == 0).astype(np.uint8), (ifreq == 0).astype(np.uint8))
# IFWMF: p = normalised popularity, wt = 1 / (1 + rho p)
rho = 1e5
pu = ufreq / n_items; pu = pu / pu.sum()
pi = ifreq / n_users; pi = pi / pi.sum()
wu = (1.0 / (1.0 + rho * pu)).astype(np.float32); wi = (1.0 / (1.0 + rho * pi)).astype(np.float32)
# TMF: z-score of the rarer side's frequency over the concatenated frequencies, k = clamp(ceil(r sigma(rho (z - alpha))), 1, r)
allf = np.concatenate([ufreq, ifreq]).astype(np.float64)
mean, std = allf.mean(), allf.std()
def tmf_rank(f, rho_t=20.0, alpha=0.5):
    z = (f.astype(np.float64) - mean) / std
    return np.clip(np.ceil(R / (1.0 + np.exp(-rho_t * (z - alpha)))), 1, R).astype(np.int32)
ur, ir = tmf_rank(ufreq), tmf_rank(ifreq)
peak, _ = bench.measured_hbm_gbs()
users = np.repeat(np.arange(n_users), np.diff(ptr))
k_per_rating = np.where(ufreq[users] < ifreq[ind], ur[users], ir[ind])   # the rarer side decides
out = {"workload": "ML-20M-shaped synthetic ratings, rank 64 (BASELINE.json configs[3])", "n_users": n_users, "n_items": n_items,
       "train_nnz": train_nnz, "sgd_flat_inflight_frac": FRAC, "variants": {}}
for name, variant in (("MF", E.MF), ("IFWMF", E.IFWMF), ("TMF", E.TMF), ("TMFDropout", E.TMFDROPOUT)):
    if variant == E.MF:
        eng.set_aux(E.MF, ufreq, ifreq)
    elif variant == E.IFWMF:
        eng.set_aux(variant, ufreq, ifreq, wu, wi)
    else:
        eng.set_aux(variant, ufreq, ifreq, ur, ir, ur, ir, poisson_cdf_table(R) if variant == E.TMFDROPOUT else None)
    eng.upload_factors(U0, V0)
    eng.sgd_plan(1)
    _, cold, lists = eng.debug_sgd_records(0, 0, with_records=False)
    ms, curve = [], []
    for ep in range(8):
        eng.event_record(0)
        eng.sgd_epoch_flat(variant, 0.005, 0.05, 0.05, 1, ep)
        eng.event_record(1)
        ms.append(eng.event_elapsed_ms(0, 1))
        curve.append(round(eng.rmse(E.VAL, E.CURRENT, variant), 4))
    m = float(np.median(ms[2:]))
    kbar = float(k_per_rating.mean()) if variant in (E.TMF, E.TMFDROPOUT) else float(R)
    bytes_per = 16 * kbar + 12 + (8 if variant == E.IFWMF else 0)
    out["variants"][name] = {"ms_per_epoch": m, "updates_per_sec": train_nnz / (m * 1e-3), "mean_effective_rank": kbar,
                             "algorithmic_bytes_per_update": bytes_per, "algorithmic_gbs": bytes_per * train_nnz / (m * 1e-3) / 1e9,
                             "frac_of_hbm": bytes_per * train_nnz / (m * 1e-3) / 1e9 / peak, "hot_lists": int(len(lists)),
                             "hot_share": 1.0 - cold / train_nnz, "val_rmse_by_epoch": curve}
    print(name, json.dumps(out["variants"][name]), file=sys.stderr, flush=True)
print(json.dumps(out))
eng.close()
