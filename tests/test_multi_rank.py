"""The multi-rank paths on whatever GPUs the box has — one is enough: the ranks are engines of ONE process connected
with mfb_comm_connect_local (the host classes' way of using several GPUs), spread round-robin over the visible devices.
Kernels, peer stores, sequence flags and barriers are exactly those of the one-process-per-GPU path (bench.py --gpus N).

  * row-sharded ALS and CCD++ (deliberately UNBALANCED shards: a rank that finishes early must not overwrite what a slow
    peer still reads) == one unsharded engine;
  * DSGD over N ranks == the oracle's trainSGDPar with P = N, value for value, on a conflict-free matrix;
  * DSGD over 2 / 4 / 8 ranks on the 1/20-scale bench matrix, reference partitions + update sequences, lr 0.002:
    validation RMSE against the oracle's trainSGDPar with the same P (committed curves: profiles/r2_dsgd_*).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    nd = torch.cuda.device_count()
    return [r % nd for r in range(n)]


def _cuts(ptr, fracs):
    nnz = int(ptr[-1])
    c = [0] + [int(np.searchsorted(ptr, f * nnz)) for f in np.cumsum(fracs)[:-1]] + [len(ptr) - 1]
    return c


def _sharded_group(E, tr, va, n_users, n_items, r, fracs, bad_u, bad_i, seed=3):
    devs = _devices(len(fracs))
    colptr = np.zeros(n_items + 1, np.int64)
    np.cumsum(np.bincount(tr.rowind, minlength=n_items), out=colptr[1:])
    ucut, icut = _cuts(tr.rowptr, fracs), _cuts(colptr, fracs)
    rng = np.random.default_rng(seed)
    U0 = rng.uniform(-0.01, 0.01, (n_users, r)).astype(np.float32)
    V0 = rng.uniform(-0.01, 0.01, (n_items, r)).astype(np.float32)
    engs = []
    for k, d in enumerate(devs):
        eng = E.Engine(n_users, n_items, r, device=d)
        eng.upload_csr(E.TRAIN, tr, with_csc=True)
        eng.upload_csr(E.VAL, va, with_csc=False)
        eng.set_masks(bad_u, bad_i)
        eng.set_aux(E.MF, np.diff(tr.rowptr).astype(np.int32), np.bincount(tr.rowind, minlength=n_items).astype(np.int32))
        eng.upload_factors(U0, V0)
        engs.append(eng)
    if len(engs) > 1:
        E.connect_local(engs)
        for k, eng in enumerate(engs):
            eng.set_row_range(E.USER, ucut[k], ucut[k + 1])
            eng.set_row_range(E.ITEM, icut[k], icut[k + 1])
    return engs


def _close(engs):
    for e in engs:
        e.sync()
    if len(engs) > 1:
        for e in engs:
            assert not e.comm_error()
            e.comm_disconnect()
    for e in engs:
        e.close()


def _problem():
    from matfac_b200 import synth
    tr, va, te = synth.make_splits(3000, 1500, 300000, seed=21)
    n_users, n_items = tr.nrows, max(tr.ncols, va.ncols, te.ncols)
    bad_u = (np.diff(tr.rowptr) == 0).astype(np.uint8)
    bad_i = np.ones(n_items, np.uint8)
    bad_i[: tr.ncols] = (np.bincount(tr.rowind, minlength=tr.ncols) == 0)
    return tr, va, n_users, n_items, bad_u, bad_i


@pytest.mark.parametrize("r,fracs", [(16, (0.5, 0.5)), (128, (0.08, 0.22, 0.7)), (64, (0.7, 0.1, 0.1, 0.1))])
def test_sharded_als_equals_single_engine(r, fracs):
    from common import rel_err
    from matfac_b200 import engine as E
    tr, va, n_users, n_items, bad_u, bad_i = _problem()
    one = _sharded_group(E, tr, va, n_users, n_items, r, (1.0,), bad_u, bad_i)
    sh = _sharded_group(E, tr, va, n_users, n_items, r, fracs, bad_u, bad_i)
    for ep in range(2):
        for side in (E.USER, E.ITEM):
            for eng in one + sh:
                eng.als_half_step(side, 0.1)
    U1, V1 = one[0].download_factors()
    for eng in sh:  # every rank holds the complete factors after the fused all-gather
        U2, V2 = eng.download_factors()
        assert rel_err(U2, U1) < 1e-5 and rel_err(V2, V1) < 1e-5, (rel_err(U2, U1), rel_err(V2, V1))
    _close(one); _close(sh)


@pytest.mark.parametrize("fracs", [(0.5, 0.5), (0.05, 0.15, 0.8), (0.85, 0.05, 0.05, 0.05)])
def test_sharded_ccdpp_equals_single_engine_with_unbalanced_shards(fracs):
    from common import rel_err
    from matfac_b200 import engine as E
    tr, va, n_users, n_items, bad_u, bad_i = _problem()
    r = 8
    one = _sharded_group(E, tr, va, n_users, n_items, r, (1.0,), bad_u, bad_i)
    sh = _sharded_group(E, tr, va, n_users, n_items, r, fracs, bad_u, bad_i)
    for eng in one + sh:
        eng.ccdpp_begin()
    for it in range(2):
        for k in range(r):
            for eng in one + sh:  # step by step for all ranks: the barriers between passes run on the device
                eng.ccdpp_rank1(k, it == 0, 5, 0.05, 0.05, 75)
    for eng in one + sh:
        eng.ccdpp_end()
    U1, V1 = one[0].download_factors()
    for eng in sh:
        U2, V2 = eng.download_factors()
        assert rel_err(U2, U1) < 1e-5 and rel_err(V2, V1) < 1e-5, (rel_err(U2, U1), rel_err(V2, V1))
    _close(one); _close(sh)


@pytest.mark.parametrize("world", [2, 3])
def test_dsgd_ranks_match_oracle_value_for_value(world):
    """One user per item: blocks do not interact, so N ranks with the reference plan must reproduce the oracle's
    trainSGDPar factors to fp32 rounding — partitions, update sequences, routing and pushes all have to be right."""
    import oracle_lib as ol
    from common import disjoint_problem, rel_err
    from matfac_b200 import dsgd
    from matfac_b200 import engine as E
    tr, va, te = disjoint_problem()
    od = ol.OracleData(tr, va, te)
    epochs = 3
    om = ol.OracleModel(od, algo="mf", facdim=16, maxiter=epochs, seed=5, nthreads=world, ureg=0.05, ireg=0.05, learnrate=0.01)
    om.compute_invalid()
    bu, bi = om.invalid()
    U0, V0 = om.factors()
    devs = _devices(world)
    d = dsgd.Dsgd(od.n_users, od.n_items, 16, world, dict(enumerate(devs)), (tr.rowptr, tr.rowind, tr.rowval),
                  (va.rowptr, va.rowind, va.rowval), U0, V0, bu, bi, epochs * world, plan="reference", seed=5,
                  block_order=0)  # the reference's visiting order inside a block: value-for-value comparable
    d.run(0, epochs * world, 0.01, 0.05, 0.05, 5)
    d.publish()
    om.train("sgdpar")
    Uo, Vo = om.factors()
    # user rows live on their owner rank, item rows are complete everywhere after publish()
    for r, eng in d.engines.items():
        U, V = eng.download_factors()
        mine = d.user_part == r
        assert rel_err(U[mine], Uo[mine]) < 2e-5, (r, rel_err(U[mine], Uo[mine]))
        assert rel_err(V, Vo) < 2e-5, (r, rel_err(V, Vo))
        assert not eng.comm_error()
    s = d.eval_sums(E.VAL)
    assert abs(np.sqrt(s[0] / s[1]) - om.rmse(1)) <= 1e-5 * om.rmse(1)
    d.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dsgd_ranks_rmse_matches_oracle_on_bench_shape(world):
    """bench.py's N > 1 path at 1/20 scale (24 k x 888, 5 M ratings, rank 64, lr 0.002): N ranks, the reference's
    partitions and update sequences (seed 1) against the oracle's trainSGDPar with P = N for seeds 1 and 2.  Past the
    knee of the curve the device must sit within 0.5 % of the band the oracle's two seeds span; the final value within
    0.5 % of that band as well (band = the two oracle curves widened by their own distance, see below)."""
    import oracle_lib as ol
    from matfac_b200 import dsgd, synth
    from matfac_b200 import engine as E
    nu, ni, nnz = int(480_189 * 0.05), int(17_770 * 0.05), int(100_480_507 * 0.05)
    prob = synth.skewed_problem(nu, ni, nnz, 20260102, device="cuda:0")
    tr, va = synth.Csr(nu, ni, *prob["train"]), synth.Csr(nu, ni, *prob["val"])
    od = ol.OracleData(tr, va, va)
    epochs, knee = 14, 8
    oracle = []
    U0 = V0 = bu = bi = None
    for seed in (1, 2):
        om = ol.OracleModel(od, algo="mf", facdim=64, maxiter=epochs, seed=seed, nthreads=world, ureg=0.05, ireg=0.05,
                            learnrate=0.002)
        if seed == 1:
            om.compute_invalid()
            bu, bi = om.invalid()
            U0, V0 = om.factors()
        om.train("sgdpar", keep_history=True)
        oracle.append([h[3] for h in om.history()])
        del om
    devs = _devices(world)
    d = dsgd.Dsgd(nu, ni, 64, world, dict(enumerate(devs)), prob["train"], prob["val"], U0, V0, bu, bi, epochs * world,
                  plan="reference", seed=1)
    got = []
    for ep in range(epochs):
        d.run(ep * world, (ep + 1) * world, 0.002, 0.05, 0.05, 1)
        d.publish()
        s = d.eval_sums(E.VAL)
        d.barrier()
        got.append(float(np.sqrt(s[0] / s[1])))
    assert not any(e.comm_error() for e in d.engines.values())
    d.close()
    lo = [min(a, b) for a, b in zip(*oracle)]
    hi = [max(a, b) for a, b in zip(*oracle)]
    msg = (world, got, oracle)
    # The stratified trainer's curve is noisy while it still falls (an epoch draws its P schedules at random; at P = 2
    # the oracle's own two seeds are up to 15 % apart at equal epochs, at P = 8 about 1 %: profiles/r2_dsgd_parity.md), so
    # the band is the oracle's two curves widened by their own distance at that epoch, plus the 0.5 % bar; the device must
    # never be worse than that, and at the last epoch not worse than the oracle's worse seed + 0.5 %.
    # Below the band the bar is 5 %: with the ratings of a block in shuffled order (block_order = 1, the only order that
    # converges on this matrix at every N: profiles/r2_dsgd_parity.md) the device falls FASTER than the reference's
    # user-major sweep through the steep part of the curve (2 - 4 % lower at P = 4) and meets it again when both flatten.
    for ep in range(knee, epochs):
        spread = hi[ep] - lo[ep]
        assert (lo[ep] - spread) * 0.95 <= got[ep] <= (hi[ep] + spread) * 1.005, (ep,) + msg
    assert got[-1] <= hi[-1] * 1.005 + (hi[-1] - lo[-1]), msg
