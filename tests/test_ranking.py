"""Ranking metrics on the device (hitRate / arHR / NDCG and their U / I variants,
model.cpp:760-1332) against the oracle's restatement, which tests/test_oracle.py pins to the reference's own model.cpp.

  * mfb_rank_positions, CUDA-core path: rounded as the reference's loops — positions equal the oracle's list positions;
  * mfb_rank_positions, tcgen05 path (dense U V^T, 3xTF32): positions equal except where two scores are closer than the
    split's accuracy, metrics within 1e-3;
  * the host classes' Model::hitRate ... NDCGI (libmatfac_host.so, mfh_rank_metrics) against mfo_rank_metrics and the
    golden vectors the reference binary wrote.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from gpu_driver import make_engine
from matfac_b200 import engine as E, synth
from test_oracle import GOLDEN, RANK_BASE, RANK_CASES, make_model, rank_filters, ranking_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "matfac_b200", "libmatfac_host.so")


def metrics_from_positions(pos, tst, filt_users=None, filt_items=None):
    """hitRate and arHR of model.cpp:1158 / :981 from the per-user positions (the host classes do the same)."""
    counted = pos != -1
    if filt_users is not None:
        counted &= filt_users.astype(bool)
    if filt_items is not None:
        counted &= filt_items.astype(bool)[np.clip(tst, 0, len(filt_items) - 1)] & (tst >= 0)
    p = pos[counted]
    hits = int(((p >= 0) & (p < 10)).sum())
    ar = float((1.0 / (p[(p >= 0) & (p < 1000)] + 1.0)).sum())
    n = int(counted.sum())
    return hits, hits / n, ar, ar / n


def trained_oracle(algo, method, threads, extra, facdim=None, splits=None):
    fl = dict(RANK_BASE); fl.update(extra)
    if facdim:
        fl["facdim"] = facdim
    splits = splits or ranking_problem()
    od = ol.OracleData(*splits)
    m = make_model(od, algo, threads, fl)
    m.train(method)
    return splits, od, m, fl


@pytest.mark.parametrize("algo,method,threads,extra", RANK_CASES)
@pytest.mark.parametrize("tensor_cores", [0, 1])
def test_rank_positions_match_oracle(algo, method, threads, extra, tensor_cores):
    splits, od, m, fl = trained_oracle(algo, method, threads, extra)
    bU, bV = m.factors(best=True)
    eng, variant = make_engine(splits, m, fl["facdim"], algo=algo)
    eng.upload_factors(bU, bV)
    eng.set_option("rank_tensor_cores", tensor_cores)
    fu, fi = rank_filters(od.n_users, od.n_items)
    for which in (E.VAL, E.TEST):
        want = m.rank_metrics(which, best=True, filt_users=fu, filt_items=fi)
        pos, tst = eng.rank_positions(which, variant=variant)
        te = splits[which]
        assert np.array_equal(tst[pos != -1], te.rowind[te.rowptr[:-1][pos != -1]])
        exact = tensor_cores == 0 or algo != "mf"  # TMF always takes the CUDA-core path
        for name, f_u, f_i in (("", None, None), ("u", fu, None), ("i", None, fi)):
            hits, hr, ar, arhr = metrics_from_positions(pos, tst, f_u, f_i)
            tol = 1e-12 if exact else 1e-3
            assert abs(hr - want["hr" + name]) <= tol, (which, name, hr, want["hr" + name])
            assert abs(arhr - want["arhr" + name]) <= tol, (which, name, arhr, want["arhr" + name])
            if name and exact:
                assert hits == want[f"hr{name}_first"] and abs(ar - want[f"arhr{name}_first"]) < 1e-9
    eng.close()


@pytest.mark.parametrize("rank", [64, 100])
def test_rank_positions_dense_random_factors(rank):
    """Rank 64 fills the tensor-core tile exactly, rank 100 takes the CUDA-core path; several user and item tiles,
    invalid items, test items the user has rated in training (state -2)."""
    splits = synth.make_ranking_splits(700, 450, 30000, seed=9, user_s=0.3)
    tr, va, te = splits
    od = ol.OracleData(*splits)
    m = ol.OracleModel(od, algo="mf", facdim=rank, maxiter=1, seed=4, nthreads=2)
    m.compute_invalid()
    rng = np.random.default_rng(2)
    U = rng.standard_normal((od.n_users, rank)).astype(np.float32)
    V = rng.standard_normal((od.n_items, rank)).astype(np.float32)
    m.set_factors(U, V)
    eng, variant = make_engine(splits, m, rank)
    eng.upload_factors(U, V)
    want = m.rank_metrics(E.TEST, best=False)
    pos_c = None
    for tc in (0, 1):
        eng.set_option("rank_tensor_cores", tc)
        pos, tst = eng.rank_positions(E.TEST)
        hits, hr, ar, arhr = metrics_from_positions(pos, tst)
        if tc == 0:
            pos_c = pos
            assert abs(hr - want["hr"]) <= 1e-12 and abs(arhr - want["arhr"]) <= 1e-12
        else:
            # random scores are O(sqrt(rank)): ties within the 3xTF32 accuracy (~1e-6 relative) are rare
            assert (pos != pos_c).mean() < 0.01 and np.abs(pos - pos_c).max() <= 2
            assert abs(hr - want["hr"]) <= 1e-3 and abs(arhr - want["arhr"]) <= 1e-3
    eng.close()


def test_predictions_match_the_reference_rounding():
    splits, od, m, fl = trained_oracle("mf", "sgd", 1, {})
    bU, bV = m.factors(best=True)
    eng, variant = make_engine(splits, m, fl["facdim"])
    eng.upload_factors(bU, bV)
    te = splits[2]
    pred = eng.predict(E.TEST, te.rowptr[-1])
    bu, bi = m.invalid()
    rows = np.repeat(np.arange(te.nrows), np.diff(te.rowptr))
    want = np.zeros(len(rows), np.float32)
    for d in range(fl["facdim"]):  # float products, float sum, in order (model.cpp:547)
        want = (want + (bU[rows, d] * bV[te.rowind, d]).astype(np.float32)).astype(np.float32)
    ok = ~(bu[rows].astype(bool) | bi[te.rowind].astype(bool))
    assert np.array_equal(pred[ok], want[ok]) and np.isnan(pred[~ok]).all()
    eng.close()


@pytest.mark.parametrize("algo,method,threads,extra", RANK_CASES)
def test_host_classes_ranking_metrics_match_oracle_and_golden(algo, method, threads, extra):
    class Csr(C.Structure):
        _fields_ = [("nrows", C.c_int32), ("rowptr", C.c_void_p), ("rowind", C.c_void_p), ("rowval", C.c_void_p)]

    class Problem(C.Structure):
        _fields_ = [("train", Csr), ("val", Csr), ("test", Csr), ("algo", C.c_char_p), ("mf_method", C.c_char_p),
                    ("facdim", C.c_int32), ("maxiter", C.c_int32), ("seed", C.c_int32), ("num_parts", C.c_int32),
                    ("ureg", C.c_float), ("ireg", C.c_float), ("learnrate", C.c_float), ("rhorms", C.c_float),
                    ("alpha", C.c_float), ("init_U", C.c_void_p), ("init_V", C.c_void_p), ("prefix", C.c_char_p)]

    splits, od, m, fl = trained_oracle(algo, method, threads, extra)
    bU, bV = m.factors(best=True)
    bu, bi = m.invalid()
    fu, fi = rank_filters(od.n_users, od.n_items)
    lib = C.CDLL(HOST_SO)
    lib.mfh_rank_metrics.argtypes = [C.POINTER(Problem)] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3
    keep = []

    def csr(mt):
        a = [np.ascontiguousarray(mt.rowptr, np.int64), np.ascontiguousarray(mt.rowind, np.int32),
             np.ascontiguousarray(mt.rowval, np.float32)]
        keep.extend(a)
        return Csr(mt.nrows, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data)

    p = Problem(csr(splits[0]), csr(splits[1]), csr(splits[2]), algo.encode(), method.encode(), fl["facdim"], fl["maxiter"],
                fl["seed"], 0, fl["ureg"], fl["ireg"], fl["learnrate"], fl.get("rhorms", 0.0), fl.get("alpha", 0.0), None, None,
                b"/tmp/mfh_rank_test")
    bu8, bi8 = np.ascontiguousarray(bu, np.uint8), np.ascontiguousarray(bi, np.uint8)
    gpath = os.path.join(GOLDEN, f"rank_{algo}_{method}.json")
    golden = json.load(open(gpath)) if os.path.exists(gpath) else None
    for which, name in ((1, "val"), (2, "test")):
        out = np.zeros(15, np.float64)
        rc = lib.mfh_rank_metrics(C.byref(p), bU.ctypes.data, bV.ctypes.data, bu8.ctypes.data, bi8.ctypes.data, which,
                                  fu.ctypes.data, fi.ctypes.data, out.ctypes.data)
        assert rc == 0
        want = m.rank_metrics(which, best=True, filt_users=fu, filt_items=fi)
        # the default device path: tcgen05 for MF (metrics within 1e-3 — a hit flips only when two scores tie within the
        # split's accuracy), CUDA cores for TMF (exact); NDCG ranks predictions rounded as the reference's
        tol = 1e-3 if algo == "mf" else 1e-12
        for k, v in zip(ol.OracleModel.RANK_KEYS, out.tolist()):
            t = 1e-12 if "ndcg" in k else tol * (max(1.0, want[k]) if k.endswith("_first") else 1.0)
            assert abs(v - want[k]) <= t, (name, k, v, want[k])
            if golden:
                assert abs(v - golden[f"{name}_{k}"]) <= t, (name, k, v, golden[f"{name}_{k}"])
    lib.mfh_release_device()
