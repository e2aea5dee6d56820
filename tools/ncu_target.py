#!/usr/bin/env python
"""A small, single-purpose workload for ncu: one kernel family on one of the bench shapes.

  python tools/ncu_target.py --shape netflix|yahoo --algo sgd|als|ccdpp|eval [--rank 64] [--scale 1.0] [--reps 2]

Run it plainly first (it must exit 0), then under ncu, e.g.
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:"sgd_flat|sgd_hot" --csv --log-file gpurun_out/x.csv python tools/ncu_target.py --shape yahoo --algo sgd
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from matfac_b200 import engine as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="netflix", choices=["netflix", "yahoo"])
    ap.add_argument("--algo", default="sgd", choices=["sgd", "als", "ccdpp", "eval"])
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--warm", type=int, default=3)
    a = ap.parse_args()
    prob = bench.make_problem(bench.SHAPE if a.shape == "netflix" else bench.YAHOO_SHAPE, a.scale, "cuda:0")
    nu, ni = prob["n_users"], prob["n_items"]
    bad_u, bad_i, cnt_i = bench.masks_of(prob)
    U0, V0 = bench.init_factors(nu, ni, a.rank)
    eng = E.Engine(nu, ni, a.rank)
    eng.upload_csr(E.TRAIN, bench.Mat(nu, ni, prob["train"]), with_csc=False)
    eng.upload_csr(E.VAL, bench.Mat(nu, ni, prob["val"]), with_csc=False)
    eng.set_masks(bad_u, bad_i)
    eng.set_aux(E.MF, np.diff(prob["train"][0]).astype(np.int32), cnt_i.astype(np.int32))
    eng.upload_factors(U0, V0)
    hp = bench.HP
    if a.algo == "sgd":
        eng.sgd_plan(1)
        for ep in range(a.warm + a.reps):  # the warm-up epochs bring the factors to their steady scale
            eng.sgd_epoch_flat(E.MF, hp["lr"], hp["ureg"], hp["ireg"], 1, ep)
    elif a.algo == "als":
        eng.build_csc(E.TRAIN)
        for _ in range(1 + a.reps):
            eng.als_half_step(E.USER, 0.1)
            eng.als_half_step(E.ITEM, 0.1)
    elif a.algo == "ccdpp":
        eng.build_csc(E.TRAIN)
        eng.ccdpp_begin()
        for k in range(2):
            eng.ccdpp_rank1(k, True, 5, 0.05, 0.05, 75)
        for k in range(a.reps):
            eng.ccdpp_rank1(k, False, 5, 0.05, 0.05, 75)
        eng.ccdpp_end()
    else:
        for _ in range(1 + a.reps):
            eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
    eng.sync()
    print("ncu_target ok: launches", E.launch_count(), "train_nnz", int(prob["train"][0][-1]), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
