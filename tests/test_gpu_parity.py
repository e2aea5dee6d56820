"""Device-vs-oracle parity tests (run on the B200 box: pytest -m gpu).  Everything goes through
the C ABI of matfac_b200/libmfb.so; the oracle (oracle/libmf_oracle.so) is only the checker.

Tolerances (BASELINE.json north_star): ALS / CCD++ factors within 1e-4 relative per epoch;
SGD validation/test RMSE within 0.5 % of the reference at equal epochs; evaluation sums 1e-6.
"""
import numpy as np
import pytest

import oracle_lib as ol
from common import disjoint_problem, rel_err, small_problem
from gpu_driver import make_engine, run_sgd
from matfac_b200 import engine as E

pytestmark = pytest.mark.gpu

HP = dict(ureg=0.05, ireg=0.05, learnrate=0.01)
ALGO_FLAGS = {"mf": {}, "IFWMF": dict(rhorms=1000.0), "TMF": dict(rhorms=20.0, alpha=0.5),
              "TMFDropout": dict(rhorms=20.0, alpha=0.5)}


def oracle_model(splits, algo, rank, maxiter=1, seed=3, nthreads=1, **kw):
    od = ol.OracleData(*splits)
    hp = dict(HP); hp.update(ALGO_FLAGS[algo]); hp.update(kw)
    return ol.OracleModel(od, algo=algo, facdim=rank, maxiter=maxiter, seed=seed, nthreads=nthreads, **hp)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["mf", "IFWMF", "TMF", "TMFDropout"])
@pytest.mark.parametrize("rank", [10, 64])
def test_eval_matches_oracle(algo, rank):
    """rmse_reduce (K10+K11): objective and masked RMSE, at init and after oracle training."""
    splits = small_problem()
    om = oracle_model(splits, algo, rank, maxiter=3, nthreads=2)
    eng, variant = make_engine(splits, om, rank, algo, rho=ALGO_FLAGS[algo].get("rhorms", 0.0))
    for phase in ("init", "trained"):
        if phase == "trained":
            om.train("sgdpar" if algo != "mf" else "sgd")
            eng.upload_factors(*om.factors())
        for which in (E.TRAIN, E.VAL, E.TEST):
            got, want = eng.rmse(which, E.CURRENT, variant), om.rmse(which)
            assert abs(got - want) <= 1e-6 * want, (phase, which, got, want)
        got, want = eng.objective(HP["ureg"], HP["ireg"], variant), om.objective()
        assert abs(got - want) <= 1e-6 * abs(want), (phase, got, want)
    eng.close()


def test_eval_masks_and_ragged_inputs():
    """Invalid users/items are skipped; val/test files may be shorter than the train matrix and
    may hold items that never occur in train (datastruct.cpp:91, model.cpp:235)."""
    from matfac_b200 import synth
    tr, va, te = small_problem(300, 200, 8000, seed=5)
    # drop all training ratings of a few users and of a few items -> invalid ids
    keep = ~np.isin(np.repeat(np.arange(tr.nrows), np.diff(tr.rowptr)), [3, 17, 150]) & ~np.isin(tr.rowind, [5, 9])
    users = np.repeat(np.arange(tr.nrows, dtype=np.int32), np.diff(tr.rowptr))[keep]
    tr2 = synth.coo_to_csr(users, tr.rowind[keep], tr.rowval[keep], tr.nrows).build_csc()
    # a val matrix with an item beyond every train column, and a test matrix with fewer rows
    vu = np.repeat(np.arange(va.nrows, dtype=np.int32), np.diff(va.rowptr))
    order = np.argsort(np.append(vu, 7), kind="stable")
    va2 = synth.coo_to_csr(np.append(vu, 7)[order].astype(np.int32), np.append(va.rowind, 260)[order].astype(np.int32),
                           np.append(va.rowval, 4.0)[order].astype(np.float32), va.nrows).build_csc()
    tu = np.repeat(np.arange(te.nrows, dtype=np.int32), np.diff(te.rowptr))
    m = tu < 250
    te2 = synth.coo_to_csr(tu[m], te.rowind[m], te.rowval[m], 250).build_csc()
    splits = (tr2, va2, te2)
    om = oracle_model(splits, "mf", 8)
    assert om.data.n_items == 261
    eng, variant = make_engine(splits, om, 8)
    bu, bi = om.invalid()
    assert bu[[3, 17, 150]].all() and bi[[5, 9, 260]].all()
    for which in (E.TRAIN, E.VAL, E.TEST):
        o = eng.eval(which)
        got, want = np.sqrt(o[0] / o[1]), om.rmse(which)
        assert abs(got - want) <= 1e-6 * want
    assert abs(eng.objective(0.05, 0.05) - om.objective()) <= 1e-6 * om.objective()
    eng.close()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["mf", "IFWMF", "TMF"])
@pytest.mark.parametrize("rank,P", [(10, 1), (64, 3), (128, 2), (5, 2)])
def test_sgd_conflict_free_matches_oracle(algo, rank, P):
    """With one user per item no two runs touch the same row, so the device epoch must equal
    the oracle's stratified epoch up to fp32-vs-double gradient rounding."""
    splits = disjoint_problem()
    epochs = 3
    om = oracle_model(splits, algo, rank, maxiter=epochs, nthreads=P)
    eng, variant = make_engine(splits, om, rank, algo, rho=ALGO_FLAGS[algo].get("rhorms", 0.0), with_csc=False)
    up, ip, sched = om.dsgd_plan(P, epochs * P)
    if P == 1:
        eng.sgd_plan(1)
    else:
        eng.sgd_plan(P, up, ip)
    assert eng.sgd_block_nnz(np.array([[a, b] for a in range(P) for b in range(P)], np.int32)) == splits[0].nnz
    run_sgd(eng, variant, epochs, HP["learnrate"], HP["ureg"], HP["ireg"], P=P, schedule=sched)
    om.train("sgdpar")
    U, V = eng.download_factors()
    Uo, Vo = om.factors()
    assert rel_err(U, Uo) < 2e-5 and rel_err(V, Vo) < 2e-5
    eng.close()


def test_sgd_zero_learning_rate_is_identity():
    splits = small_problem()
    om = oracle_model(splits, "mf", 64)
    eng, variant = make_engine(splits, om, 64, with_csc=False)
    eng.sgd_plan(1)
    U0, V0 = eng.download_factors()
    run_sgd(eng, variant, 2, 0.0, 0.0, 0.0)
    U1, V1 = eng.download_factors()
    assert np.array_equal(U0, U1) and np.array_equal(V0, V1)
    eng.close()


@pytest.mark.parametrize("algo,method,P,rank", [
    ("mf", "sgd", 1, 10), ("mf", "sgdpar", 4, 10), ("mf", "sgdpar", 8, 64), ("IFWMF", "sgd", 1, 10),
    ("IFWMF", "sgdpar", 4, 16), ("TMF", "sgdpar", 4, 16), ("TMFDropout", "sgdpar", 4, 16)])
def test_sgd_rmse_parity(algo, method, P, rank):
    """RMSE within 0.5 % of the oracle at equal epochs (north_star): the device visits ratings in
    the stratified trainers' order (user-major) with item rows shared Hogwild-style."""
    splits = small_problem(3000, 1500, 300000, seed=21)
    epochs = 12
    flags = dict(ALGO_FLAGS[algo])
    if algo == "IFWMF":
        flags["rhorms"] = 100.0
    om = oracle_model(splits, algo, rank, maxiter=epochs, nthreads=P, **flags)
    eng, variant = make_engine(splits, om, rank, algo, rho=flags.get("rhorms", 0.0), with_csc=False)
    up, ip, sched = om.dsgd_plan(P, epochs * P)
    if P == 1:
        eng.sgd_plan(1)
    else:
        eng.sgd_plan(P, up, ip)
    curve = []
    run_sgd(eng, variant, epochs, HP["learnrate"], HP["ureg"], HP["ireg"], P=P, schedule=sched, seed=3,
            on_epoch=lambda ep: curve.append((eng.rmse(E.VAL, E.CURRENT, variant), eng.rmse(E.TEST, E.CURRENT, variant))))
    om.train(method, keep_history=True)
    hist = om.history()
    assert len(hist) == epochs
    for ep in (3, 7, epochs - 1):
        want_val = hist[ep][3]
        assert abs(curve[ep][0] - want_val) <= 0.005 * want_val, (ep, curve[ep][0], want_val)
    want_test = om.rmse(2)
    assert abs(curve[-1][1] - want_test) <= 0.005 * want_test
    # learning happened at all
    assert curve[-1][0] < 0.6 * curve[0][0] or curve[-1][0] < 1.2
    eng.close()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rank", [10, 64, 128])
def test_als_epoch_matches_oracle(rank):
    """Per-epoch (teacher-forced) ALS parity: factors within 1e-4 relative (north_star)."""
    splits = small_problem(500, 300, 40000, seed=13)
    om = oracle_model(splits, "mf", rank, maxiter=1, ureg=0.1, ireg=0.1, nthreads=4)
    eng, variant = make_engine(splits, om, rank)
    for ep in range(3):
        U0, V0 = om.factors()
        eng.upload_factors(U0, V0)
        eng.als_half_step(E.USER, 0.1)
        eng.als_half_step(E.ITEM, 0.1)
        om.train("als")
        U, V = eng.download_factors()
        Uo, Vo = om.factors()
        assert rel_err(U, Uo) < 1e-4, (ep, rel_err(U, Uo))
        assert rel_err(V, Vo) < 1e-4, (ep, rel_err(V, Vo))
        assert abs(eng.rmse(E.VAL) - om.rmse(1)) < 1e-4 * om.rmse(1)
    eng.close()


def test_als_long_rows_are_split():
    """A row longer than the per-CTA chunk goes through the workspace path."""
    from matfac_b200 import synth
    rng = np.random.default_rng(2)
    n_users, n_items = 6000, 40
    dense_items = [0, 1]
    users = np.concatenate([np.arange(n_users), np.arange(n_users), rng.integers(0, n_users, 20000)]).astype(np.int64)
    items = np.concatenate([np.zeros(n_users), np.ones(n_users), rng.integers(2, n_items, 20000)]).astype(np.int64)
    key = np.unique(users * n_items + items)
    users, items = (key // n_items).astype(np.int32), (key % n_items).astype(np.int32)
    vals = (np.round(rng.uniform(1, 5, users.shape[0]) * 2) / 2).astype(np.float32)
    vals[::17] = 0.0  # ratings <= 0 are skipped by ALS (modelMF.cpp:819)
    tr = synth.coo_to_csr(users, items, vals, n_users).build_csc()
    va = synth.coo_to_csr(users[::9], items[::9], vals[::9], n_users).build_csc()
    splits = (tr, va, va)
    om = oracle_model(splits, "mf", 16, maxiter=1, ureg=0.1, ireg=0.1, nthreads=4)
    eng, _ = make_engine(splits, om, 16)
    eng.als_half_step(E.USER, 0.1)
    eng.als_half_step(E.ITEM, 0.1)
    om.train("als")
    U, V = eng.download_factors()
    Uo, Vo = om.factors()
    assert rel_err(U, Uo) < 1e-4 and rel_err(V, Vo) < 1e-4
    eng.close()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("freq_adap", [False, True])
@pytest.mark.parametrize("rank", [8, 64])
def test_ccdpp_matches_oracle(rank, freq_adap):
    splits = small_problem(500, 300, 40000, seed=13)
    epochs = 3
    om = oracle_model(splits, "mf", rank, maxiter=epochs, nthreads=4)
    eng, variant = make_engine(splits, om, rank)
    order = np.tile(np.arange(rank, dtype=np.int32), (epochs, 1)) if freq_adap else ol.ccdpp_dim_order(3, rank, epochs)
    om.train("ccd++" if freq_adap else "ccdpp_plain", keep_history=True)
    hist = om.history()
    eng.ccdpp_begin()
    for ep in range(epochs):
        for k in order[ep]:
            eng.ccdpp_rank1(int(k), ep == 0, 5, HP["ureg"], HP["ireg"], 75 if freq_adap else 0)
        U, V = eng.download_factors()
        assert rel_err(U, hist[ep][0]) < 1e-4, (ep, rel_err(U, hist[ep][0]))
        assert rel_err(V, hist[ep][1]) < 1e-4, (ep, rel_err(V, hist[ep][1]))
        obj = eng.objective(HP["ureg"], HP["ireg"])
        assert abs(obj - hist[ep][2]) < 1e-4 * hist[ep][2]
    eng.ccdpp_end()
    eng.close()


def test_snapshot_and_restore_best():
    splits = small_problem()
    om = oracle_model(splits, "mf", 10)
    eng, variant = make_engine(splits, om, 10, with_csc=False)
    eng.sgd_plan(1)
    U0, V0 = eng.download_factors()
    eng.snapshot_best()
    run_sgd(eng, variant, 1, 0.01, 0.05, 0.05)
    Ub, Vb = eng.download_factors(E.BEST)
    assert np.array_equal(U0, Ub) and np.array_equal(V0, Vb)
    U1, _ = eng.download_factors()
    assert not np.array_equal(U0, U1)
    eng.restore_best()
    U2, V2 = eng.download_factors()
    assert np.array_equal(U0, U2) and np.array_equal(V0, V2)
    eng.close()


def test_pack_unpack_rows_roundtrip():
    import torch
    splits = small_problem()
    om = oracle_model(splits, "mf", 10)
    eng, _ = make_engine(splits, om, 10, with_csc=False)
    _, V0 = eng.download_factors()
    ids = np.array([5, 1, 399, 42], np.int32)
    ptr, ld = eng.device_factors(E.ITEM)
    assert ld == 12
    buf = torch.zeros(ids.shape[0] * ld, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    eng.pack_rows(E.ITEM, ids, buf.data_ptr())
    packed = buf.cpu().numpy().reshape(ids.shape[0], ld)
    assert np.array_equal(packed[:, :10], V0[ids]) and not packed[:, 10:].any()
    eng.upload_factors(None, np.zeros_like(V0))
    eng.unpack_rows(E.ITEM, ids, buf.data_ptr())
    _, V1 = eng.download_factors()
    assert np.array_equal(V1[ids], V0[ids])
    mask = np.ones(V0.shape[0], bool); mask[ids] = False
    assert not V1[mask].any()
    eng.close()
