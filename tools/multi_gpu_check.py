#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/multi_gpu_check.py

  1. row-sharded ALS over peer memory == the same epochs on one unsharded engine (to rounding);
  2. row-sharded CCD++ == unsharded (to rounding);
  3. DSGD with item blocks pushed between ranks: validation RMSE within 0.5 % of the same schedule run
     on one engine.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from matfac_b200 import dsgd, engine as E, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tr, va, te = synth.make_splits(3000, 1500, 300000, seed=21)
    n_users, n_items = tr.nrows, max(tr.ncols, va.ncols, te.ncols)
    bad_u = (np.diff(tr.rowptr) == 0).astype(np.uint8)
    bad_i = np.ones(n_items, np.uint8)
    bad_i[: tr.ncols] = (np.bincount(tr.rowind, minlength=tr.ncols) == 0)
    ok = True

    def rel(a, b):
        return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))

    def fresh(r, sharded):
        eng = E.Engine(n_users, n_items, r, device=local)
        eng.upload_csr(E.TRAIN, tr, with_csc=True)
        eng.upload_csr(E.VAL, va, with_csc=False)
        eng.set_masks(bad_u, bad_i)
        eng.set_aux(E.MF, np.diff(tr.rowptr).astype(np.int32), np.bincount(tr.rowind, minlength=n_items).astype(np.int32))
        rng = np.random.default_rng(3)
        eng.upload_factors(rng.uniform(-0.01, 0.01, (n_users, r)).astype(np.float32),
                           rng.uniform(-0.01, 0.01, (n_items, r)).astype(np.float32))
        if sharded:
            blobs = [None] * world
            dist.all_gather_object(blobs, eng.comm_init(rank, world))
            eng.comm_connect(blobs)
            eng.set_row_range(E.USER, n_users * rank // world, n_users * (rank + 1) // world)
            eng.set_row_range(E.ITEM, n_items * rank // world, n_items * (rank + 1) // world)
        return eng

    # ---- 1. ALS ----
    for r in (16, 128):
        one, sh = fresh(r, False), fresh(r, True)
        for ep in range(2):
            for eng in (one, sh):
                eng.als_half_step(E.USER, 0.1)
                eng.als_half_step(E.ITEM, 0.1)
        U1, V1 = one.download_factors()
        U2, V2 = sh.download_factors()
        # rows split over several CTAs sum their partial Grams with atomics (order varies run to run)
        du, dv = rel(U2, U1), rel(V2, V1)
        same = du < 1e-5 and dv < 1e-5
        print(f"[rank {rank}] ALS rank {r}: sharded vs single engine: rel diff U {du:.2e} V {dv:.2e}; comm_error {sh.comm_error()}", flush=True)
        ok &= same and not sh.comm_error()
        sh.comm_barrier(); sh.sync(); dist.barrier()
        one.close(); sh.close()

    # ---- 2. CCD++ ----
    one, sh = fresh(8, False), fresh(8, True)
    for eng in (one, sh):
        eng.ccdpp_begin()
        for it in range(2):
            for k in range(8):
                eng.ccdpp_rank1(k, it == 0, 5, 0.05, 0.05, 75)
        eng.ccdpp_end()
    U1, V1 = one.download_factors()
    U2, V2 = sh.download_factors()
    du, dv = rel(U2, U1), rel(V2, V1)
    same = du < 1e-5 and dv < 1e-5
    print(f"[rank {rank}] CCD++: sharded vs single engine: rel diff U {du:.2e} V {dv:.2e}", flush=True)
    ok &= same and not sh.comm_error()
    sh.comm_barrier(); sh.sync(); dist.barrier()
    one.close(); sh.close()

    # ---- 3. DSGD ----
    r, P, epochs = 16, world, 30
    prng = np.random.default_rng(7)
    user_part = prng.integers(0, P, n_users).astype(np.int32)
    item_part = prng.integers(0, P, n_items).astype(np.int32)
    item_part[bad_i != 0] = -1
    user_part[bad_u != 0] = -1
    sched = dsgd.random_schedule(P, epochs * P + 1, seed=11)
    one = fresh(r, False)
    one.sgd_plan(P, user_part, item_part)
    one.set_option("sgd_block_order", 1)
    for t in range(epochs * P):
        blocks = np.stack([np.arange(P), sched[t]], 1).astype(np.int32)
        one.sgd_subepoch(blocks, E.MF, 0.005, 0.05, 0.05, 1, t)
    want = one.rmse(E.VAL)
    sh = E.Engine(n_users, n_items, r, device=local)
    mine = user_part == rank

    def local_rows(m):
        rows = np.repeat(mine, np.diff(m.rowptr))
        lptr = np.zeros(n_users + 1, np.int64)
        np.cumsum(np.where(mine, np.diff(m.rowptr), 0), out=lptr[1:])
        return synth.Csr(n_users, m.ncols, lptr, m.rowind[rows], m.rowval[rows])

    sh.upload_csr(E.TRAIN, local_rows(tr), with_csc=False)
    sh.upload_csr(E.VAL, local_rows(va), with_csc=False)
    sh.set_masks(bad_u, bad_i)
    rng = np.random.default_rng(3)
    sh.upload_factors(rng.uniform(-0.01, 0.01, (n_users, r)).astype(np.float32),
                      rng.uniform(-0.01, 0.01, (n_items, r)).astype(np.float32))
    sh.sgd_plan(P, np.where(mine, user_part, -1).astype(np.int32), item_part)
    sh.set_option("sgd_block_order", 1)
    blobs = [None] * world
    dist.all_gather_object(blobs, sh.comm_init(rank, world))
    sh.comm_connect(blobs)
    tp = dsgd.EngineTransport(sh, rank)
    for ep in range(epochs):
        dsgd.run_steps(sched, ep * P, (ep + 1) * P, rank, tp,
                       lambda block, t: sh.sgd_subepoch(np.array([[rank, block]], np.int32), E.MF, 0.005, 0.05, 0.05, 1, t))
    dsgd.publish(sched, epochs * P - 1, rank, tp)
    ev = sh.eval(E.VAL)
    t = torch.tensor([ev[0], ev[1]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    got = float(np.sqrt(t[0].item() / t[1].item()))
    good = abs(got - want) <= 0.005 * want and not sh.comm_error()
    print(f"[rank {rank}] DSGD val RMSE {got:.5f} vs single engine {want:.5f}: {'ok' if good else 'MISMATCH'}", flush=True)
    ok &= good
    # the item matrix is identical on all ranks after publish
    _, Vmine = sh.download_factors()
    vt = torch.from_numpy(Vmine).cuda()
    v0 = vt.clone()
    dist.broadcast(v0, 0)
    same = bool(torch.equal(vt, v0))
    print(f"[rank {rank}] V identical to rank 0 after publish: {same}", flush=True)
    ok &= same
    sh.comm_barrier(); sh.sync(); dist.barrier()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if flag.item() == 1.0 else "FAIL", flush=True)
    one.close(); sh.close()
    dist.destroy_process_group()
    return 0 if flag.item() == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
