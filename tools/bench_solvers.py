#!/usr/bin/env python
"""ALS / CCD++ / evaluation epoch timings on the Netflix-shaped bench matrix (BASELINE.json configs[2]
and the CCD++ row of SURVEY.md §8d), one GPU, CUDA events on the engine's stream.

  python tools/bench_solvers.py --algo als --rank 128 [--scale 1.0] [--tc 0|1] [--epochs 3]
  python tools/bench_solvers.py --algo ccdpp --rank 64
  python tools/bench_solvers.py --algo eval --rank 64

Prints one JSON line per algorithm with the epoch time and the roofline figures of SURVEY.md §8(d).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def csc_on_device(n_users, n_items, ptr, ind, val, device):
    """gk_csr_CreateIndex(mat, GK_CSR_COL) as a stable sort by column (torch, set-up only)."""
    import torch
    dev = torch.device(device)
    ind_d = torch.from_numpy(ind).to(dev)
    deg = torch.from_numpy(np.diff(ptr)).to(dev)
    rows = torch.repeat_interleave(torch.arange(n_users, device=dev, dtype=torch.int32), deg)
    order = torch.sort(ind_d.long() * n_users + rows.long()).indices  # (col, row) ascending == stable by col
    colind = rows[order].cpu().numpy()
    colval = torch.from_numpy(val).to(dev)[order].cpu().numpy()
    cnt = torch.bincount(ind_d.long(), minlength=n_items)
    colptr = np.zeros(n_items + 1, np.int64)
    colptr[1:] = torch.cumsum(cnt, 0).cpu().numpy()
    del ind_d, deg, rows, order, cnt
    torch.cuda.empty_cache()
    return colptr, colind, colval


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="als", choices=["als", "ccdpp", "eval"])
    ap.add_argument("--rank", type=int, default=128)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--tc", type=int, default=1)
    ap.add_argument("--reg", type=float, default=0.1)
    args = ap.parse_args()
    import torch
    from matfac_b200 import engine as E
    n_users, n_items, nnz = int(bench.SHAPE[0] * args.scale), bench.SHAPE[1], int(bench.SHAPE[2] * args.scale)
    t0 = time.time()
    prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
    ptr, ind, val = prob["train"]
    train_nnz = int(ptr[-1])
    tr = bench.Mat(n_users, n_items, prob["train"])
    tr.colptr, tr.colind, tr.colval = csc_on_device(n_users, n_items, ptr, ind, val, "cuda:0")
    va = bench.Mat(n_users, n_items, prob["val"])
    bench.log(f"data {train_nnz} ratings in {time.time()-t0:.1f}s")
    r = args.rank
    rng = np.random.default_rng(1)
    U0 = rng.uniform(-0.01, 0.01, size=(n_users, r)).astype(np.float32)
    V0 = rng.uniform(-0.01, 0.01, size=(n_items, r)).astype(np.float32)
    eng = E.Engine(n_users, n_items, r)
    eng.upload_csr(E.TRAIN, tr, with_csc=True)
    eng.upload_csr(E.VAL, va, with_csc=False)
    eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))
    eng.set_aux(E.MF, np.diff(ptr).astype(np.int32), np.bincount(ind, minlength=n_items).astype(np.int32))
    eng.upload_factors(U0, V0)
    peak_gbs, _ = bench.measured_hbm_gbs()
    out = {"algo": args.algo, "rank": r, "n_users": n_users, "n_items": n_items, "train_nnz": train_nnz}
    if args.algo == "als":
        eng.set_option("als_tensor_cores", args.tc)
        eng.set_option("als_dual", int(os.environ.get("ALS_DUAL", "1")))
        lens = np.diff(ptr)
        out.update(als_dual=int(os.environ.get("ALS_DUAL", "1")), users_le16=int((lens <= 16).sum()), users_le32=int((lens <= 32).sum()), users_le64=int((lens <= 64).sum()))
        eng.als_half_step(E.USER, args.reg)
        eng.als_half_step(E.ITEM, args.reg)  # warm-up epoch (plans are built here)
        eng.sync()
        ms_u, ms_i = [], []
        for ep in range(args.epochs):
            eng.event_record(0)
            eng.als_half_step(E.USER, args.reg)
            eng.event_record(1)
            eng.als_half_step(E.ITEM, args.reg)
            eng.event_record(2)
            eng.sync()
            ms_u.append(eng.event_elapsed_ms(0, 1))
            ms_i.append(eng.event_elapsed_ms(1, 2))
        rp = 16 if r <= 16 else 32 if r <= 32 else 64 if r <= 64 else 128
        ms = float(np.median(ms_u) + np.median(ms_i))
        gram_flop = 4.0 * r * r * train_nnz
        out.update(tensor_cores=args.tc, ms_user=float(np.median(ms_u)), ms_item=float(np.median(ms_i)), epoch_ms=ms,
                   gram_tflops_alg=gram_flop / (ms * 1e-3) / 1e12,
                   tensor_tflops_issued=(3 if (args.tc and r > 64) else 1) * 4.0 * rp * rp * train_nnz / (ms * 1e-3) / 1e12,
                   gather_gbs=2.0 * train_nnz * r * 4 / (ms * 1e-3) / 1e9, val_rmse=eng.rmse(E.VAL))
    elif args.algo == "ccdpp":
        eng.ccdpp_begin()
        dims = min(r, 8)
        for k in range(dims):  # iter 0 (no add-back)
            eng.ccdpp_rank1(k, True, 5, 0.05, 0.05, 75)
        eng.sync()
        eng.event_record(0)
        for k in range(dims):
            eng.ccdpp_rank1(k, False, 5, 0.05, 0.05, 75)
        eng.event_record(1)
        eng.sync()
        ms_k = eng.event_elapsed_ms(0, 1) / dims
        out.update(ms_per_rank1=ms_k, epoch_ms=ms_k * r, algorithmic_gbs=128.0 * train_nnz / (ms_k * 1e-3) / 1e9,
                   frac_of_hbm=128.0 * train_nnz / (ms_k * 1e-3) / 1e9 / peak_gbs, dims_timed=dims)
        eng.ccdpp_end()
    else:
        for _ in range(2):
            eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
        ms = []
        for _ in range(5):
            eng.event_record(0)
            eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
            eng.event_record(1)
            eng.sync()
            ms.append(eng.event_elapsed_ms(0, 1))
        m = float(np.median(ms))
        gbs = (8.0 + 8.0 * r) * train_nnz / (m * 1e-3) / 1e9
        out.update(ms_objective_pass=m, algorithmic_gbs=gbs, frac_of_hbm=gbs / peak_gbs)
    print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
