"""The oracle (oracle/mf_oracle.cpp) against the reference's own translation units
(oracle/_ref/mf_ref, built from /root/reference with third-party stand-ins) and against the
golden vectors that run produced (tests/golden/, see tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from matfac_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = [
    # algo, method, threads, extra flags
    ("mf", "sgd", 1, {}),
    ("mf", "hogsgd", 1, {}),
    ("mf", "sgdu", 1, {}),
    ("mf", "sgdpar", 4, {}),
    ("mf", "als", 2, {"ureg": 0.1, "ireg": 0.1}),
    ("mf", "ccdpp_plain", 2, {}),
    ("mf", "ccd++", 2, {}),
    ("mf", "ccd", 1, {}),
    ("IFWMF", "sgd", 1, {"rhorms": 1000.0}),
    ("IFWMF", "sgdpar", 3, {"rhorms": 1000.0}),
    ("TMF", "sgd", 4, {"rhorms": 20.0, "alpha": 0.5}),
    ("TMFDropout", "sgd", 4, {"rhorms": 20.0, "alpha": 0.5}),
]
BASE = dict(facdim=8, maxiter=6, seed=3, ureg=0.05, ireg=0.05, learnrate=0.01)


def golden_problem():
    return synth.make_splits(300, 200, 6000, seed=5)


def make_model(od, algo, threads, fl):
    return ol.OracleModel(od, algo=algo, facdim=fl["facdim"], maxiter=fl["maxiter"], seed=fl["seed"], nthreads=threads,
                          ureg=fl["ureg"], ireg=fl["ireg"], learnrate=fl["learnrate"], rhorms=fl.get("rhorms", 0.0),
                          alpha=fl.get("alpha", 0.0))


@pytest.mark.parametrize("algo,method,threads,extra", CASES)
def test_oracle_bit_exact_against_reference_binary(tmp_path, algo, method, threads, extra):
    if not ol.have_ref():
        pytest.skip("oracle/_ref/mf_ref not built (no /root/reference on this machine)")
    files = synth.write_split_files(str(tmp_path), *golden_problem())
    fl = dict(BASE); fl.update(extra)
    ref = ol.run_ref(files, str(tmp_path / "dump"), algo=algo, method=method, threads=threads, **fl)
    od = ol.OracleData(files=files)
    m = make_model(od, algo, threads, fl)
    U0, V0 = m.factors()
    assert np.array_equal(U0, ref["init_uFac"]) and np.array_equal(V0, ref["init_iFac"])
    m.train(method)
    U, V = m.factors()
    bU, bV = m.factors(best=True)
    assert np.array_equal(U, ref["last_uFac"]) and np.array_equal(V, ref["last_iFac"])
    assert np.array_equal(bU, ref["best_uFac"]) and np.array_equal(bV, ref["best_iFac"])
    bu, bi = m.invalid()
    assert np.array_equal(np.nonzero(bu)[0], ref["invalidUsers"]) and np.array_equal(np.nonzero(bi)[0], ref["invalidItems"])
    assert abs(m.rmse(1, best=True) - ref["best_val_rmse"]) < 1e-12
    assert abs(m.rmse(2, best=True) - ref["best_test_rmse"]) < 1e-12
    assert abs(m.objective() - ref["last_objective"]) < 1e-9 * abs(ref["last_objective"])
    assert abs(m.learn_rate - ref["learn_rate"]) < 1e-10  # dumped with %.9g


def test_csr_reader_and_csc_index_match_reference(tmp_path):
    if not ol.have_ref():
        pytest.skip("oracle/_ref/mf_ref not built")
    tr, va, te = golden_problem()
    files = synth.write_split_files(str(tmp_path), tr, va, te)
    ol.run_ref(files, str(tmp_path / "dump"), facdim=4, maxiter=1)
    od = ol.OracleData(files=files)
    for which, name, mat in ((0, "train", tr), (1, "val", va), (2, "test", te)):
        d = ol.read_csr_dump(str(tmp_path / "dump" / f"{name}.csr.bin"))
        ptr, ind, val = od.csr(which)
        cptr, cind, cval = od.csc(which)
        assert np.array_equal(ptr, d["rowptr"]) and np.array_equal(ind, d["rowind"]) and np.array_equal(val, d["rowval"])
        assert np.array_equal(cptr, d["colptr"]) and np.array_equal(cind, d["colind"]) and np.array_equal(cval, d["colval"])
        # and the numpy generator's own CSC agrees (it feeds the engine in the tests)
        assert np.array_equal(mat.colptr[: len(cptr)], cptr) and np.array_equal(mat.colind, cind)


@pytest.mark.parametrize("algo,method,threads,extra", CASES)
def test_oracle_against_golden_vectors(algo, method, threads, extra):
    """Golden vectors were written by the reference binary in the build container; this test also
    runs on machines where /root/reference does not exist."""
    path = os.path.join(GOLDEN, f"ref_{algo}_{method.replace('+', 'p')}.npz")
    if not os.path.exists(path):
        pytest.skip("golden vector missing: run tests/golden/make_golden.py in the build container")
    g = np.load(path)
    fl = dict(BASE); fl.update(extra)
    od = ol.OracleData(*golden_problem())
    m = make_model(od, algo, threads, fl)
    U0, V0 = m.factors()
    assert np.array_equal(U0, g["init_uFac"]) and np.array_equal(V0, g["init_iFac"])
    m.train(method)
    U, V = m.factors()
    bU, bV = m.factors(best=True)
    assert np.array_equal(U, g["last_uFac"]) and np.array_equal(V, g["last_iFac"])
    assert np.array_equal(bU, g["best_uFac"]) and np.array_equal(bV, g["best_iFac"])
    assert abs(m.rmse(1, best=True) - float(g["best_val_rmse"])) < 1e-12


def test_ldlt_against_float64_solve():
    rng = np.random.default_rng(0)
    n, r = 50, 24
    X = rng.normal(size=(n, 40, r)).astype(np.float32)
    A = np.einsum("nkr,nks->nrs", X, X).astype(np.float32) + 0.1 * np.eye(r, dtype=np.float32)
    b = rng.normal(size=(n, r)).astype(np.float32)
    x = ol.ldlt_solve(A, b)
    want = np.linalg.solve(A.astype(np.float64), b.astype(np.float64)[..., None])[..., 0]
    assert np.abs(x - want).max() / np.abs(want).max() < 1e-4


def test_stratum_schedule_is_a_permutation_matrix():
    od = ol.OracleData(*golden_problem())
    m = make_model(od, "mf", 4, dict(BASE))
    up, ip, sched = m.dsgd_plan(4, 12)
    assert set(np.unique(up)) <= {0, 1, 2, 3} and set(np.unique(ip)) <= {-1, 0, 1, 2, 3}
    for s in range(12):
        assert sorted(sched[s][:, 0]) == [0, 1, 2, 3] and sorted(sched[s][:, 1]) == [0, 1, 2, 3]
    # the boundary quirk of modelMF.cpp:241-249: part 0 holds perPart + 1 ids
    counts = np.bincount(up[up >= 0], minlength=4)
    per = (up >= 0).sum() // 4
    assert counts[0] == per + 1 and counts[1] == per


# ---- ranking metrics (model.cpp:760-1332) -----------------------------------------------------------------------
RANK_CASES = [("mf", "sgd", 1, {}), ("TMF", "sgd", 2, {"rhorms": 20.0, "alpha": 0.5})]
RANK_BASE = dict(facdim=8, maxiter=12, seed=3, ureg=0.05, ireg=0.05, learnrate=0.02)


def ranking_problem():
    return synth.make_ranking_splits(300, 200, 9000, seed=5, user_s=0.3)


def rank_filters(n_users, n_items):
    fu = np.zeros(n_users, np.uint8); fu[::3] = 1   # the sets ref_driver.cpp builds for --rank_metrics 1
    fi = np.zeros(n_items, np.uint8); fi[::2] = 1
    return fu, fi


def oracle_rank_dict(m, od):
    fu, fi = rank_filters(od.n_users, od.n_items)
    out = {}
    for which, name in ((1, "val"), (2, "test")):
        for k, v in m.rank_metrics(which, best=True, filt_users=fu, filt_items=fi).items():
            out[f"{name}_{k}"] = v
    return out


@pytest.mark.parametrize("algo,method,threads,extra", RANK_CASES)
def test_oracle_ranking_metrics_match_reference_binary(tmp_path, algo, method, threads, extra):
    """hitRate / arHR / NDCG and their U / I variants of the best model, computed by the reference's own model.cpp."""
    if not ol.have_ref():
        pytest.skip("oracle/_ref/mf_ref not built (no /root/reference on this machine)")
    files = synth.write_split_files(str(tmp_path), *ranking_problem())
    fl = dict(RANK_BASE); fl.update(extra)
    ref = ol.run_ref(files, str(tmp_path / "dump"), algo=algo, method=method, threads=threads, rank_metrics=1, **fl)
    od = ol.OracleData(files=files)
    m = make_model(od, algo, threads, fl)
    m.train(method)
    got = oracle_rank_dict(m, od)
    assert ref["val_hr"] > 0.02 and ref["val_ndcg"] > 0.5  # a trained model: the metrics are not trivially zero
    for k, v in got.items():
        assert abs(v - ref[k]) <= 1e-12 * max(1.0, abs(ref[k])), (k, v, ref[k])


@pytest.mark.parametrize("algo,method,threads,extra", RANK_CASES)
def test_oracle_ranking_metrics_against_golden_vectors(algo, method, threads, extra):
    path = os.path.join(GOLDEN, f"rank_{algo}_{method}.json")
    if not os.path.exists(path):
        pytest.skip("golden vector missing: run tests/golden/make_golden.py in the build container")
    g = json.load(open(path))
    fl = dict(RANK_BASE); fl.update(extra)
    od = ol.OracleData(*ranking_problem())
    m = make_model(od, algo, threads, fl)
    m.train(method)
    for k, v in oracle_rank_dict(m, od).items():
        assert abs(v - g[k]) <= 1e-12 * max(1.0, abs(g[k])), (k, v, g[k])
