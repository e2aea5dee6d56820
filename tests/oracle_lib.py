"""TEST INFRASTRUCTURE — ctypes binding of oracle/libmf_oracle.so (the CPU restatement of the
reference's training path) and helpers to run oracle/_ref/mf_ref (the reference's own
translation units compiled unmodified) and read its binary dumps.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmf_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "mf_ref")

ALGO = {"mf": 0, "IFWMF": 1, "TMF": 2, "TMFDropout": 3}
METHOD = {"sgd": 0, "sgdpar": 1, "als": 2, "ccdpp_plain": 3, "ccd++": 4, "hogsgd": 5, "sgdu": 6, "ccd": 7}


class Params(C.Structure):
    _fields_ = [("facDim", C.c_int), ("maxIter", C.c_int), ("seed", C.c_int), ("nThreads", C.c_int),
                ("uReg", C.c_float), ("iReg", C.c_float), ("learnRate", C.c_float),
                ("rhoRMS", C.c_float), ("alpha", C.c_float)]


def build_oracle() -> str:
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "mf_oracle.cpp")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        L = _lib
        L.mfo_data_read.restype = C.c_void_p
        L.mfo_data_read.argtypes = [C.c_char_p] * 3
        L.mfo_data_from_arrays.restype = C.c_void_p
        L.mfo_data_from_arrays.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p] * 3
        L.mfo_data_free.argtypes = [C.c_void_p]
        L.mfo_data_nusers.argtypes = [C.c_void_p]
        L.mfo_data_nitems.argtypes = [C.c_void_p]
        L.mfo_data_dims.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.mfo_data_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_data_csc.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_model_create.restype = C.c_void_p
        L.mfo_model_create.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int]
        L.mfo_model_free.argtypes = [C.c_void_p]
        L.mfo_train.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.mfo_get_factors.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mfo_set_factors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_history_len.argtypes = [C.c_void_p]
        L.mfo_get_history.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_learn_rate.restype = C.c_float
        L.mfo_learn_rate.argtypes = [C.c_void_p]
        L.mfo_epoch_seconds.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.mfo_get_invalid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_compute_invalid.argtypes = [C.c_void_p, C.c_void_p]
        L.mfo_rmse.restype = C.c_double
        L.mfo_rmse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.mfo_objective.restype = C.c_double
        L.mfo_objective.argtypes = [C.c_void_p, C.c_void_p]
        L.mfo_dsgd_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_rank_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_tmf_ranks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.mfo_ifw_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_ccdpp_dim_order.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.mfo_ccd_dim_orders.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.mfo_ldlt_solve.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleData:
    def __init__(self, train=None, val=None, test=None, files=None):
        L = lib()
        if files is not None:
            self.h = L.mfo_data_read(*[f.encode() for f in files])
            if not self.h:
                raise IOError("oracle could not read %r" % (files,))
        else:
            args = []
            self._keep = []
            for m in (train, val, test):
                ptr = np.ascontiguousarray(m.rowptr, dtype=np.int64)
                ind = np.ascontiguousarray(m.rowind, dtype=np.int32)
                val_ = np.ascontiguousarray(m.rowval, dtype=np.float32)
                self._keep += [ptr, ind, val_]
                args += [m.nrows, _p(ptr), _p(ind), _p(val_)]
            self.h = L.mfo_data_from_arrays(*args)
        self.n_users = L.mfo_data_nusers(self.h)
        self.n_items = L.mfo_data_nitems(self.h)

    def dims(self, which):
        d = np.zeros(3, dtype=np.int64)
        lib().mfo_data_dims(self.h, which, _p(d))
        return tuple(int(x) for x in d)

    def csr(self, which):
        nr, nc, nnz = self.dims(which)
        ptr = np.zeros(nr + 1, np.int64); ind = np.zeros(nnz, np.int32); val = np.zeros(nnz, np.float32)
        lib().mfo_data_csr(self.h, which, _p(ptr), _p(ind), _p(val))
        return ptr, ind, val

    def csc(self, which):
        nr, nc, nnz = self.dims(which)
        ptr = np.zeros(nc + 1, np.int64); ind = np.zeros(nnz, np.int32); val = np.zeros(nnz, np.float32)
        lib().mfo_data_csc(self.h, which, _p(ptr), _p(ind), _p(val))
        return ptr, ind, val

    def __del__(self):
        if getattr(self, "h", None):
            lib().mfo_data_free(self.h)
            self.h = None


class OracleModel:
    def __init__(self, data: OracleData, algo="mf", facdim=8, maxiter=10, seed=1, nthreads=1, ureg=0.01,
                 ireg=0.01, learnrate=0.005, rhorms=0.0, alpha=0.0):
        self.data = data
        self.p = Params(facdim, maxiter, seed, nthreads, ureg, ireg, learnrate, rhorms, alpha)
        self.r = facdim
        self.algo = algo
        self.h = lib().mfo_model_create(data.h, C.byref(self.p), ALGO[algo])

    def train(self, method="sgd", keep_history=False) -> int:
        return lib().mfo_train(self.h, self.data.h, METHOD[method], int(keep_history))

    def factors(self, best=False):
        U = np.zeros((self.data.n_users, self.r), np.float32)
        V = np.zeros((self.data.n_items, self.r), np.float32)
        lib().mfo_get_factors(self.h, int(best), _p(U), _p(V))
        return U, V

    def set_factors(self, U, V):
        U = np.ascontiguousarray(U, np.float32); V = np.ascontiguousarray(V, np.float32)
        lib().mfo_set_factors(self.h, _p(U), _p(V))

    RANK_KEYS = ("hr", "arhr", "ndcg", "hru_first", "hru", "arhru_first", "arhru", "ndcgu_first", "ndcgu",
                 "hri_first", "hri", "arhri_first", "arhri", "ndcgi_first", "ndcgi")

    def rank_metrics(self, which, best=False, filt_users=None, filt_items=None):
        """Ranking metrics of model.cpp:760-1332 as a dict keyed like the lines ref_driver writes (RANK_KEYS)."""
        out = np.zeros(15, np.float64)
        fu = None if filt_users is None else np.ascontiguousarray(filt_users, np.uint8)
        fi = None if filt_items is None else np.ascontiguousarray(filt_items, np.uint8)
        lib().mfo_rank_metrics(self.h, self.data.h, which, int(best), _p(fu), _p(fi), _p(out))
        return dict(zip(self.RANK_KEYS, out.tolist()))

    def history(self):
        out = []
        for e in range(lib().mfo_history_len(self.h)):
            U = np.zeros((self.data.n_users, self.r), np.float32)
            V = np.zeros((self.data.n_items, self.r), np.float32)
            o = C.c_double(); v = C.c_double()
            lib().mfo_get_history(self.h, e, _p(U), _p(V), C.byref(o), C.byref(v))
            out.append((U, V, o.value, v.value))
        return out

    @property
    def learn_rate(self):
        return float(lib().mfo_learn_rate(self.h))

    def epoch_seconds(self):
        out = np.zeros(4096, np.float64)
        n = lib().mfo_epoch_seconds(self.h, _p(out), 4096)
        return out[:n].copy()

    def compute_invalid(self):
        lib().mfo_compute_invalid(self.h, self.data.h)

    def invalid(self):
        u = np.zeros(self.data.n_users, np.uint8); i = np.zeros(self.data.n_items, np.uint8)
        lib().mfo_get_invalid(self.h, _p(u), _p(i))
        return u, i

    def rmse(self, which=1, best=False):
        return float(lib().mfo_rmse(self.h, self.data.h, which, int(best)))

    def objective(self):
        return float(lib().mfo_objective(self.h, self.data.h))

    def dsgd_plan(self, P, n_subepochs):
        up = np.zeros(self.data.n_users, np.int32); ip = np.zeros(self.data.n_items, np.int32)
        sched = np.zeros((n_subepochs, P, 2), np.int32)
        lib().mfo_dsgd_plan(self.h, self.data.h, P, n_subepochs, _p(up), _p(ip), _p(sched))
        return up, ip, sched

    def tmf_ranks(self, for_prediction=False):
        nr, nc, _ = self.data.dims(0)
        ur = np.zeros(nr, np.int32); ir = np.zeros(nc, np.int32)
        lib().mfo_tmf_ranks(self.h, _p(ur), _p(ir), int(for_prediction))
        return ur, ir

    def ifw_weights(self):
        pu = np.zeros(self.data.n_users, np.float64); pi = np.zeros(self.data.n_items, np.float64)
        lib().mfo_ifw_weights(self.h, self.data.h, _p(pu), _p(pi))
        return pu, pi

    def ccd_dim_orders(self, n_epochs):
        """trainCCD's per-row dims orders: list over epochs of (user orders [n_users][r], item orders [n_items][r]) uint8,
        rows of invalid ids left as the identity (they are never visited)."""
        bu, bi = self.invalid()
        vu, vi = np.nonzero(bu == 0)[0], np.nonzero(bi == 0)[0]
        r = self.r
        flat = np.zeros((n_epochs, len(vu) + len(vi), r), np.uint8)
        lib().mfo_ccd_dim_orders(self.h, self.data.h, n_epochs, _p(flat))
        out = []
        for e in range(n_epochs):
            du = np.tile(np.arange(r, dtype=np.uint8), (self.data.n_users, 1)); di = np.tile(np.arange(r, dtype=np.uint8), (self.data.n_items, 1))
            du[vu] = flat[e, :len(vu)]; di[vi] = flat[e, len(vu):]
            out.append((du, di))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().mfo_model_free(self.h)
            self.h = None


def ccdpp_dim_order(seed, r, n_epochs):
    out = np.zeros((n_epochs, r), np.int32)
    lib().mfo_ccdpp_dim_order(seed, r, n_epochs, _p(out))
    return out


def ldlt_solve(A, b):
    A = np.ascontiguousarray(A, np.float32); b = np.ascontiguousarray(b, np.float32)
    n, r = b.shape
    x = np.zeros_like(b)
    lib().mfo_ldlt_solve(n, r, _p(A), _p(b), _p(x))
    return x


# ---------------------------------------------------------------------------------------------
# oracle/_ref/mf_ref: the reference's own code.  Exists only where /root/reference was present
# at build time (the build container); the binary travels to the GPU box with the snapshot.
def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def read_mat(path):
    with open(path, "rb") as f:
        nr, nc = np.frombuffer(f.read(8), np.int32)
        return np.frombuffer(f.read(), np.float32).reshape(nr, nc).copy()


def read_set(path):
    with open(path, "rb") as f:
        n = int(np.frombuffer(f.read(8), np.int64)[0])
        return np.frombuffer(f.read(), np.int32)[:n].copy()


def read_csr_dump(path):
    with open(path, "rb") as f:
        nr, nc, nnz = (int(x) for x in np.frombuffer(f.read(24), np.int64))
        rowptr = np.frombuffer(f.read(8 * (nr + 1)), np.int64).copy()
        rowind = np.frombuffer(f.read(4 * nnz), np.int32).copy()
        rowval = np.frombuffer(f.read(4 * nnz), np.float32).copy()
        colptr = np.frombuffer(f.read(8 * (nc + 1)), np.int64).copy()
        colind = np.frombuffer(f.read(4 * nnz), np.int32).copy()
        colval = np.frombuffer(f.read(4 * nnz), np.float32).copy()
    return dict(nrows=nr, ncols=nc, rowptr=rowptr, rowind=rowind, rowval=rowval, colptr=colptr,
                colind=colind, colval=colval)


def run_ref(files, dump_dir, algo="mf", method="sgd", threads=1, timeout=600, **flags):
    """Run the reference binary; returns dict of dumps. flags: facdim, maxiter, seed, ureg, ..."""
    os.makedirs(dump_dir, exist_ok=True)
    cmd = [REF_BIN, "--trainmat", files[0], "--valmat", files[1], "--testmat", files[2], "--prefix",
           os.path.join(dump_dir, "ref"), "--algo", algo, "--mf_method", method, "--dump", dump_dir]
    for k, v in flags.items():
        cmd += ["--" + k, str(v)]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    out = subprocess.run(cmd, env=env, cwd=dump_dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=timeout, check=True).stdout.decode(errors="replace")
    res = {"stdout": out}
    for name in ("init_uFac", "init_iFac", "last_uFac", "last_iFac", "best_uFac", "best_iFac"):
        res[name] = read_mat(os.path.join(dump_dir, name + ".bin"))
    res["invalidUsers"] = read_set(os.path.join(dump_dir, "invalidUsers.bin"))
    res["invalidItems"] = read_set(os.path.join(dump_dir, "invalidItems.bin"))
    with open(os.path.join(dump_dir, "result.txt")) as f:
        for line in f:
            k, v = line.split(None, 1)
            res[k] = v.strip() if k == "signature" else float(v)
    return res
