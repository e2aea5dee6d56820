"""Drives the engine through the C ABI the way the host trainers do (tests only).

The frequency-derived inputs (invalid masks, partitions, schedules, TMF ranks, IFWMF
popularities, CCD++ dimension order) come from the oracle here, so that a device-vs-oracle
difference can only come from the kernels.  The product's own host code for the same quantities
lives in matfac_b200/host/ and is checked bit-for-bit against the oracle in test_host_*.py.
"""
from __future__ import annotations

import numpy as np

import oracle_lib as ol
from matfac_b200 import engine as E


def poisson_cdf_table(r):
    """row l-1: P(Poisson(l) <= k), k = 0..r-1 (float64 recurrence, stored fp32)."""
    t = np.zeros((r, r), np.float64)
    for lam in range(1, r + 1):
        p = np.exp(-float(lam))
        c = p
        for k in range(r):
            t[lam - 1, k] = c
            p = p * lam / (k + 1)
            c += p
    return np.minimum(t, 1.0).astype(np.float32)


def make_engine(splits, om: ol.OracleModel, rank, algo="mf", rho=0.0, with_csc=True):
    tr, va, te = splits
    od = om.data
    eng = E.Engine(od.n_users, od.n_items, rank)
    eng.upload_csr(E.TRAIN, tr, with_csc=with_csc)
    eng.upload_csr(E.VAL, va, with_csc=False)
    eng.upload_csr(E.TEST, te, with_csc=False)
    om.compute_invalid()
    bu, bi = om.invalid()
    eng.set_masks(bu, bi)
    U, V = om.factors()
    eng.upload_factors(U, V)
    variant = E.VARIANT[algo]
    ufreq = np.diff(tr.rowptr).astype(np.int32)
    ifreq = np.bincount(tr.rowind, minlength=tr.ncols).astype(np.int32)
    if algo == "IFWMF":
        pu, pi = om.ifw_weights()
        f32 = np.float32
        # float wt = invPop; wt = 1.0/(1.0 + rhoRMS*wt)   (modelInvPopMF.cpp:163-168)
        wu = (1.0 / (1.0 + (f32(rho) * pu.astype(f32)).astype(np.float64))).astype(f32)
        wi = (1.0 / (1.0 + (f32(rho) * pi.astype(f32)).astype(np.float64))).astype(f32)
        eng.set_aux(variant, ufreq, ifreq, wu, wi)
    elif algo in ("TMF", "TMFDropout"):
        ur, ir = om.tmf_ranks(for_prediction=False)
        up, ip = om.tmf_ranks(for_prediction=True)
        cdf = poisson_cdf_table(rank) if algo == "TMFDropout" else None
        eng.set_aux(variant, ufreq, ifreq, ur, ir, up, ip, cdf)
    elif algo == "mf":
        eng.set_aux(E.MF, ufreq, ifreq)
    return eng, variant


def run_sgd(eng, variant, epochs, lr, ureg, ireg, P=1, schedule=None, seed=1, on_epoch=None):
    """schedule: [epochs*P][P][2] block pairs (oracle's dsgd_plan) or None for P == 1."""
    for ep in range(epochs):
        if P == 1:
            eng.sgd_subepoch(np.array([[0, 0]], np.int32), variant, lr, ureg, ireg, seed, ep)
        else:
            for k in range(P):
                eng.sgd_subepoch(schedule[ep * P + k], variant, lr, ureg, ireg, seed, ep * P + k)
        if on_epoch:
            on_epoch(ep)
