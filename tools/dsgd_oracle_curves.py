#!/usr/bin/env python
"""Validation-RMSE curves of the ORACLE's stratified SGD (ModelMF::trainSGDPar restated, modelMF.cpp:154-350) for
P = 1, 2, 4, 8 threads and of its serial SGD on the 1/20-scale Netflix-shaped matrix of bench.py (lr 0.002, rank 64) —
the reference side of the DSGD-at-N-ranks parity check (tests/test_gpu_parity.py::test_dsgd_ranks_match_oracle, bench.py's
N > 1 lines).  CPU only; writes profiles/r2_dsgd_oracle_curves.json.

usage: python tools/dsgd_oracle_curves.py [--epochs 25] [--seeds 1 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from matfac_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=25)
    ap.add_argument("--seeds", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--parts", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_dsgd_oracle_curves.json"))
    a = ap.parse_args()
    nu, ni, nnz = synth.SHAPES["netflix"]
    nu, ni, nnz = int(nu * a.scale), int(ni * max(a.scale, 0.05)), int(nnz * a.scale)
    prob = synth.skewed_problem(nu, ni, nnz, 20260102)
    tr = synth.Csr(nu, ni, *prob["train"])
    va = synth.Csr(nu, ni, *prob["val"])
    od = ol.OracleData(tr, va, va)
    out = {"matrix": {"n_users": nu, "n_items": ni, "train_nnz": tr.nnz, "crc": prob["crc"]}, "rank": 64,
           "learnrate": 0.002, "ureg": 0.05, "ireg": 0.05, "epochs": a.epochs, "curves": {}}
    runs = [("sgd", 1)] + [("sgdpar", p) for p in a.parts]
    for method, P in runs:
        for seed in a.seeds:
            t0 = time.time()
            om = ol.OracleModel(od, algo="mf", facdim=64, maxiter=a.epochs, seed=seed, nthreads=P, ureg=0.05, ireg=0.05,
                                learnrate=0.002)
            om.train(method, keep_history=True)
            curve = [h[3] for h in om.history()]
            key = f"{method}_P{P}_seed{seed}"
            out["curves"][key] = curve
            print(key, f"{time.time()-t0:.0f}s", " ".join(f"{v:.4f}" for v in curve), flush=True)
            del om
            json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
