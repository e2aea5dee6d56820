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


def main():
    real_stdout = bench._protect_stdout()  # only the JSON line goes to stdout (NCCL prints its banner on fd 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="als", choices=["als", "ccdpp", "eval", "rank"])
    ap.add_argument("--rank", type=int, default=128)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--tc", type=int, default=1)
    ap.add_argument("--reg", type=float, default=0.1)
    ap.add_argument("--shape", default="netflix", choices=["netflix", "yahoo"])
    args = ap.parse_args()
    import torch
    from matfac_b200 import engine as E
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = bench.YAHOO_SHAPE if args.shape == "yahoo" else bench.SHAPE
    n_users, n_items, nnz = int(shape[0] * args.scale), shape[1], int(shape[2] * args.scale)
    t0 = time.time()
    prob = bench.gen_problem(n_users, n_items, nnz, 20260102, f"cuda:{local}")  # the same matrix on every rank
    ptr, ind, val = prob["train"]
    train_nnz = int(ptr[-1])
    tr = bench.Mat(n_users, n_items, prob["train"])
    va = bench.Mat(n_users, n_items, prob["val"])
    bench.log(f"data {train_nnz} ratings in {time.time()-t0:.1f}s")
    r = args.rank
    rng = np.random.default_rng(1)
    U0 = rng.uniform(-0.01, 0.01, size=(n_users, r)).astype(np.float32)
    V0 = rng.uniform(-0.01, 0.01, size=(n_items, r)).astype(np.float32)
    eng = E.Engine(n_users, n_items, r, device=local)
    eng.upload_csr(E.TRAIN, tr, with_csc=False)
    eng.build_csc(E.TRAIN)  # gk_csr_CreateIndex on the device
    eng.upload_csr(E.VAL, va, with_csc=False)
    eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))
    eng.set_aux(E.MF, np.diff(ptr).astype(np.int32), np.bincount(ind, minlength=n_items).astype(np.int32))
    eng.upload_factors(U0, V0)
    peak_gbs, _ = bench.measured_hbm_gbs()
    out = {"algo": args.algo, "rank": r, "n_users": n_users, "n_items": n_items, "train_nnz": train_nnz, "n_gpus": world}
    if world > 1:
        # rows sharded in contiguous ranges of (nearly) equal rating counts; every solved row / updated entry
        # is stored into all peers by the kernel that produces it (matfac_b200/csrc/comm.cu)
        blobs = [None] * world
        dist.all_gather_object(blobs, eng.comm_init(rank, world))
        eng.comm_connect(blobs)
        ucut = np.searchsorted(ptr, np.linspace(0, train_nnz, world + 1)).astype(int)
        colptr = np.zeros(n_items + 1, np.int64)
        np.cumsum(np.bincount(ind, minlength=n_items), out=colptr[1:])
        icut = np.searchsorted(colptr, np.linspace(0, train_nnz, world + 1)).astype(int)
        ucut[0] = icut[0] = 0
        ucut[-1], icut[-1] = n_users, n_items
        eng.set_row_range(E.USER, int(ucut[rank]), int(ucut[rank + 1]))
        eng.set_row_range(E.ITEM, int(icut[rank]), int(icut[rank + 1]))

    def finish(ms_list):
        """max over ranks of the per-rank median"""
        m = float(np.median(ms_list))
        if dist is not None:
            t = torch.tensor([m], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            m = float(t.item())
        return m

    if args.algo == "als":
        eng.set_option("als_tensor_cores", args.tc)
        eng.set_option("als_dual", int(os.environ.get("ALS_DUAL", "1")))
        eng.set_option("als_ws_split", int(os.environ.get("ALS_WS_SPLIT", "0")))
        if "ALS_DEBUG" in os.environ:
            eng.set_option("als_debug", int(os.environ["ALS_DEBUG"]))
            out["als_debug"] = int(os.environ["ALS_DEBUG"])
        if "ALS_CHOL_WARPS" in os.environ:
            eng.set_option("als_chol_warps", int(os.environ["ALS_CHOL_WARPS"]))
            out["als_chol_warps"] = int(os.environ["ALS_CHOL_WARPS"])
        out["als_ws_split"] = int(os.environ.get("ALS_WS_SPLIT", "0"))
        if "ALS_CHUNK" in os.environ:
            eng.set_option("als_chunk", int(os.environ["ALS_CHUNK"]))
            out["als_chunk"] = int(os.environ["ALS_CHUNK"])
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
        ms_u, ms_i = [finish(ms_u)], [finish(ms_i)]
        ms = float(ms_u[0] + ms_i[0])
        gram_flop = 4.0 * r * r * train_nnz
        out.update(tensor_cores=args.tc, ms_user=float(np.median(ms_u)), ms_item=float(np.median(ms_i)), epoch_ms=ms,
                   gram_tflops_alg=gram_flop / (ms * 1e-3) / 1e12,
                   tensor_tflops_issued=(3 if (args.tc and r > 64) else 1) * 4.0 * rp * rp * train_nnz / (ms * 1e-3) / 1e12,
                   gather_gbs=2.0 * train_nnz * r * 4 / (ms * 1e-3) / 1e9)
        ev = eng.eval(E.VAL)
        if dist is not None:
            t = torch.tensor([ev[0], ev[1]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            ev = [float(t[0]), float(t[1])]
        out.update(val_rmse=float(np.sqrt(ev[0] / max(ev[1], 1))))
    elif args.algo == "ccdpp":
        eng.set_option("ccd_fuse", int(os.environ.get("MFB_CCD_FUSE", "1")))
        eng.set_option("ccd_smem", int(os.environ.get("MFB_CCD_SMEM", "0")))
        out["ccd_smem"] = int(os.environ.get("MFB_CCD_SMEM", "0"))
        dims = min(r, 8)

        def run_ccdpp(stream, cap=4096, stage=0):
            eng.set_option("ccd_stage", stage)
            eng.set_option("ccd_stream", stream)
            eng.set_option("ccd_cap", cap)
            eng.ccdpp_begin()
            for k in range(dims):  # iter 0 (no add-back)
                eng.ccdpp_rank1(k, True, 5, 0.05, 0.05, 75)
            eng.sync()
            eng.event_record(0)
            for k in range(dims):
                eng.ccdpp_rank1(k, False, 5, 0.05, 0.05, 75)
            eng.event_record(1)
            eng.sync()
            ms = eng.event_elapsed_ms(0, 1) / dims
            eng.ccdpp_end()
            return ms

        # MFB_CCD_SWEEP = "stream[:cap[:stage]],..." (engine options ccd_stream / ccd_cap / ccd_stage) times every listed configuration in this process (this rank's time)
        sweep = os.environ.get("MFB_CCD_SWEEP", "")
        if sweep:
            out["sweep_ms_per_rank1"] = {}
            for cfg in sweep.split(","):
                out["sweep_ms_per_rank1"][cfg] = run_ccdpp(*(int(x) for x in cfg.split(":")))
                bench.log(f"ccd sweep {cfg}: {out['sweep_ms_per_rank1'][cfg]:.3f} ms")
        stream, cap, stage = int(os.environ.get("MFB_CCD_STREAM", "0")), int(os.environ.get("MFB_CCD_CAP", "4096")), int(os.environ.get("MFB_CCD_STAGE", "0"))
        out.update(ccd_stream=stream, ccd_cap=cap, ccd_stage=stage)
        ms_k = finish([run_ccdpp(stream, cap, stage) * dims]) / dims
        out.update(ms_per_rank1=ms_k, epoch_ms=ms_k * r, algorithmic_gbs=128.0 * train_nnz / (ms_k * 1e-3) / 1e9,
                   frac_of_hbm=128.0 * train_nnz / (ms_k * 1e-3) / 1e9 / peak_gbs, dims_timed=dims)
    elif args.algo == "rank":
        # hit-rate positions of every user against the validation matrix: dense U V^T (users x items x rank) with the
        # count fused into the epilogue; wall time of the whole call (split, sparse pass, GEMM, download of the positions)
        eng.upload_factors(rng.standard_normal((n_users, r)).astype(np.float32), rng.standard_normal((n_items, r)).astype(np.float32))
        res = {}
        for tc in ([1, 0] if args.scale <= 0.05 else [1]):
            eng.set_option("rank_tensor_cores", tc)
            eng.rank_positions(E.VAL)
            t = []
            for _ in range(3):
                eng.sync(); t0 = time.perf_counter()
                pos, tst = eng.rank_positions(E.VAL)
                t.append(time.perf_counter() - t0)
            sec = float(np.median(t))
            res["tcgen05" if tc else "cuda_cores"] = {"seconds": sec, "dense_tflops": 2.0 * n_users * n_items * r / sec / 1e12,
                                                     "users_counted": int((pos != -1).sum()), "hit_rate": float(((pos >= 0) & (pos < 10)).sum() / max((pos != -1).sum(), 1))}
        out.update(rank_positions=res, flops_counted="2 x users x items x rank (one product per pair; the 3xTF32 split issues three)")
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
        m = finish(ms)
        gbs = (8.0 + 8.0 * r) * train_nnz / (m * 1e-3) / 1e9
        out.update(ms_objective_pass=m, algorithmic_gbs=gbs, frac_of_hbm=gbs / peak_gbs)
    if world > 1:
        out["comm_error"] = eng.comm_error()
        eng.comm_barrier(); eng.sync(); dist.barrier()
    if rank == 0:
        print(json.dumps(out), file=real_stdout, flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
