"""ctypes binding of the engine's C ABI (include/mfb.h -> matfac_b200/libmfb.so).

This is the same boundary the C++ host classes in matfac_b200/host/ call; the Python side exists
for the parity tests and bench.py.  There is no CPU fallback: if the shared library is missing
or no CUDA device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmfb.so")

TRAIN, VAL, TEST = 0, 1, 2
CURRENT, BEST = 0, 1
USER, ITEM = 0, 1
MF, IFWMF, TMF, TMFDROPOUT = 0, 1, 2, 3
VARIANT = {"mf": MF, "IFWMF": IFWMF, "TMF": TMF, "TMFDropout": TMFDROPOUT}

# every symbol include/mfb.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "mfb_last_error", "mfb_launch_count", "mfb_device_count", "mfb_create", "mfb_destroy", "mfb_sync", "mfb_pin_host",
    "mfb_unpin_host", "mfb_upload_csr", "mfb_set_masks", "mfb_upload_factors", "mfb_download_factors",
    "mfb_set_aux", "mfb_sgd_plan", "mfb_sgd_subepoch", "mfb_sgd_block_nnz", "mfb_debug_sgd_records", "mfb_debug_sgd_hot_batch", "mfb_sgd_epoch_flat", "mfb_set_option", "mfb_als_half_step", "mfb_debug_als_gram",
    "mfb_ccdpp_begin", "mfb_ccdpp_rank1", "mfb_ccdpp_end", "mfb_ccd_half_step", "mfb_debug_chol64", "mfb_eval", "mfb_eval_groups", "mfb_rank_positions", "mfb_predict", "mfb_snapshot_best",
    "mfb_restore_best", "mfb_event_record", "mfb_event_elapsed_ms", "mfb_device_factors", "mfb_stream",
    "mfb_pack_rows", "mfb_unpack_rows", "mfb_set_row_range",
    "mfb_build_csc", "mfb_download_csc", "mfb_comm_init", "mfb_comm_connect", "mfb_comm_barrier", "mfb_comm_error", "mfb_dsgd_push_block",
    "mfb_comm_wait_block", "mfb_comm_allgather_rows", "mfb_comm_connect_local", "mfb_comm_disconnect",
]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_users", C.c_int32), ("n_items", C.c_int32), ("rank", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class EngineError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libmfb.so and declare the prototypes.  Raises if the CUDA extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(matfac_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64
    L.mfb_last_error.restype = C.c_char_p
    L.mfb_launch_count.restype = u64
    L.mfb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.mfb_destroy.argtypes = [vp]
    L.mfb_destroy.restype = None
    L.mfb_sync.argtypes = [vp]
    L.mfb_pin_host.argtypes = [vp, u64]
    L.mfb_unpin_host.argtypes = [vp]
    L.mfb_upload_csr.argtypes = [vp, C.c_int, i32, i32, i64, vp, vp, vp, vp, vp, vp]
    L.mfb_build_csc.argtypes = [vp, C.c_int]
    L.mfb_download_csc.argtypes = [vp, C.c_int, vp, vp, vp]
    L.mfb_set_masks.argtypes = [vp, vp, vp]
    L.mfb_upload_factors.argtypes = [vp, vp, i64, vp, i64]
    L.mfb_download_factors.argtypes = [vp, C.c_int, vp, i64, vp, i64]
    L.mfb_set_aux.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.mfb_sgd_plan.argtypes = [vp, i32, vp, vp]
    L.mfb_sgd_subepoch.argtypes = [vp, vp, i32, C.c_int, f32, f32, f32, u64, u64]
    L.mfb_sgd_epoch_flat.argtypes = [vp, C.c_int, f32, f32, f32, u64, u64]
    L.mfb_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.mfb_sgd_block_nnz.argtypes = [vp, vp, i32, C.POINTER(i64)]
    L.mfb_debug_sgd_records.argtypes = [vp, i32, i32, vp, C.POINTER(i64), vp, C.POINTER(i32)]
    L.mfb_debug_sgd_hot_batch.argtypes = [vp, vp]
    L.mfb_als_half_step.argtypes = [vp, C.c_int, f32]
    L.mfb_debug_als_gram.argtypes = [vp, C.c_int, i32, vp, C.POINTER(i32)]
    L.mfb_ccdpp_begin.argtypes = [vp]
    L.mfb_ccdpp_rank1.argtypes = [vp, i32, C.c_int, i32, f32, f32, i32]
    L.mfb_ccdpp_end.argtypes = [vp]
    L.mfb_ccd_half_step.argtypes = [vp, C.c_int, f32, vp]
    L.mfb_debug_chol64.argtypes = [vp, i32, vp, vp, i32, f32]
    L.mfb_eval.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.mfb_eval_groups.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.mfb_rank_positions.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.mfb_predict.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.mfb_snapshot_best.argtypes = [vp]
    L.mfb_restore_best.argtypes = [vp]
    L.mfb_event_record.argtypes = [vp, i32]
    L.mfb_event_elapsed_ms.argtypes = [vp, i32, i32, C.POINTER(f32)]
    L.mfb_device_factors.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    L.mfb_stream.argtypes = [vp]
    L.mfb_stream.restype = vp
    L.mfb_pack_rows.argtypes = [vp, C.c_int, vp, i32, vp]
    L.mfb_unpack_rows.argtypes = [vp, C.c_int, vp, i32, vp]
    L.mfb_set_row_range.argtypes = [vp, C.c_int, i32, i32]
    L.mfb_comm_init.argtypes = [vp, i32, i32, vp, C.POINTER(i64)]
    L.mfb_comm_connect.argtypes = [vp, vp, i64]
    L.mfb_comm_connect_local.argtypes = [C.POINTER(vp), i32]
    L.mfb_comm_disconnect.argtypes = [vp]
    L.mfb_comm_barrier.argtypes = [vp]
    L.mfb_comm_error.argtypes = [vp, C.POINTER(i32)]
    L.mfb_dsgd_push_block.argtypes = [vp, i32, i32, u64]
    L.mfb_comm_wait_block.argtypes = [vp, i32, u64]
    L.mfb_comm_allgather_rows.argtypes = [vp, C.c_int, vp, i32, i32]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class Engine:
    """One engine = one CUDA device + one stream (mfb_engine)."""

    def __init__(self, n_users: int, n_items: int, rank: int, device: int = 0):
        self.L = load_library()
        self.n_users, self.n_items, self.rank = int(n_users), int(n_items), int(rank)
        cfg = Config(device, n_users, n_items, rank)
        h = C.c_void_p()
        self._check(self.L.mfb_create(C.byref(cfg), C.byref(h)))
        self.h = h

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.L.mfb_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "h", None):
            self.L.mfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data ----
    def upload_csr(self, which, mat, with_csc=True):
        """mat: object with nrows, ncols, rowptr, rowind, rowval[, colptr, colind, colval]."""
        rp, ri, rv = _arr(mat.rowptr, np.int64), _arr(mat.rowind, np.int32), _arr(mat.rowval, np.float32)
        cp = ci = cv = None
        if with_csc and getattr(mat, "colptr", None) is not None:
            cp, ci, cv = _arr(mat.colptr, np.int64), _arr(mat.colind, np.int32), _arr(mat.colval, np.float32)
        self._check(self.L.mfb_upload_csr(self.h, which, mat.nrows, mat.ncols, int(rp[-1]), _p(rp), _p(ri), _p(rv),
                                          _p(cp), _p(ci), _p(cv)))

    def build_csc(self, which=TRAIN):
        """Column index built on the device from the uploaded CSR (gk_csr_CreateIndex)."""
        self._check(self.L.mfb_build_csc(self.h, which))

    def download_csc(self, which, nnz):
        cp = np.zeros(self.n_items + 1, np.int64); ci = np.zeros(nnz, np.int32); cv = np.zeros(nnz, np.float32)
        self._check(self.L.mfb_download_csc(self.h, which, _p(cp), _p(ci), _p(cv)))
        return cp, ci, cv

    def set_masks(self, invalid_users, invalid_items):
        iu, ii = _arr(invalid_users, np.uint8), _arr(invalid_items, np.uint8)
        assert iu.shape[0] == self.n_users and ii.shape[0] == self.n_items
        self._check(self.L.mfb_set_masks(self.h, _p(iu), _p(ii)))

    def upload_factors(self, U=None, V=None):
        U, V = _arr(U, np.float32), _arr(V, np.float32)
        self._check(self.L.mfb_upload_factors(self.h, _p(U), 0 if U is None else U.shape[1], _p(V),
                                              0 if V is None else V.shape[1]))

    def download_factors(self, which=CURRENT, into=None):
        """into = (U, V): caller-owned float32 buffers — with option copy_overlap they are complete after sync()."""
        if into is not None:
            U, V = into
            assert U.dtype == np.float32 and V.dtype == np.float32 and U.flags.c_contiguous and V.flags.c_contiguous
        else:
            U = np.empty((self.n_users, self.rank), np.float32)
            V = np.empty((self.n_items, self.rank), np.float32)
        self._check(self.L.mfb_download_factors(self.h, which, _p(U), U.shape[1], _p(V), V.shape[1]))
        return U, V

    def set_aux(self, variant, user_freq, item_freq, user_train=None, item_train=None, user_pred=None,
                item_pred=None, poisson_cdf=None):
        def pad(a, n, dt):
            if a is None:
                return None
            out = np.zeros(n, dtype=dt)
            a = np.asarray(a)
            out[:a.shape[0]] = a.astype(dt)
            return out
        tdt = np.float32 if variant == IFWMF else np.int32
        args = [pad(user_freq, self.n_users, np.int32), pad(item_freq, self.n_items, np.int32),
                pad(user_train, self.n_users, tdt), pad(item_train, self.n_items, tdt),
                pad(user_pred, self.n_users, np.int32), pad(item_pred, self.n_items, np.int32),
                _arr(poisson_cdf, np.float32)]
        self._check(self.L.mfb_set_aux(self.h, variant, *[_p(a) for a in args]))

    # ---- SGD ----
    def sgd_plan(self, P=1, user_part=None, item_part=None):
        up, ip = _arr(user_part, np.int32), _arr(item_part, np.int32)
        self._check(self.L.mfb_sgd_plan(self.h, P, _p(up), _p(ip)))

    def sgd_subepoch(self, blocks, variant=MF, lr=0.005, ureg=0.01, ireg=0.01, seed=0, counter=0):
        b = _arr(blocks, np.int32).reshape(-1, 2)
        self._check(self.L.mfb_sgd_subepoch(self.h, _p(b), b.shape[0], variant, lr, ureg, ireg, seed, counter))

    def sgd_epoch_flat(self, variant=MF, lr=0.005, ureg=0.01, ireg=0.01, seed=0, counter=0):
        self._check(self.L.mfb_sgd_epoch_flat(self.h, variant, lr, ureg, ireg, seed, counter))

    def set_option(self, name, value):
        self._check(self.L.mfb_set_option(self.h, name.encode(), float(value)))

    def sgd_block_nnz(self, blocks):
        b = _arr(blocks, np.int32).reshape(-1, 2)
        out = C.c_int64()
        self._check(self.L.mfb_sgd_block_nnz(self.h, _p(b), b.shape[0], C.byref(out)))
        return out.value

    def debug_sgd_records(self, user_part=0, item_part=0, with_records=True):
        """(records [n][4] int32, cold record count, lists [h][3] int32) of one stratum block (diagnostics);
        with_records=False skips the download of the records (16 bytes per rating)."""
        n = self.sgd_block_nnz([[user_part, item_part]])
        recs = np.zeros((max(n, 1), 4), np.int32) if with_records else None
        lists = np.zeros((127, 3), np.int32)
        cold, h = C.c_int64(), C.c_int32()
        self._check(self.L.mfb_debug_sgd_records(self.h, user_part, item_part, _p(recs) if with_records else None,
                                                 C.byref(cold), _p(lists), C.byref(h)))
        return (recs[:n] if with_records else None), cold.value, lists[:h.value].copy()

    def debug_sgd_hot_batch(self):
        """[sum degree x |u|^2, sum degree, ratings per round the hot CTAs last used] (diagnostics)."""
        out = np.zeros(3, np.float64)
        self._check(self.L.mfb_debug_sgd_hot_batch(self.h, _p(out)))
        return out

    # ---- ALS / CCD++ ----
    def als_half_step(self, side, reg):
        self._check(self.L.mfb_als_half_step(self.h, side, reg))

    def debug_als_gram(self, side, row):
        rp = C.c_int32()
        out = np.zeros(128 * 128 + 128, np.float32)
        self._check(self.L.mfb_debug_als_gram(self.h, side, row, _p(out), C.byref(rp)))
        R = rp.value
        return out[: R * R].reshape(R, R).copy(), out[R * R: R * R + R].copy()

    def ccdpp_begin(self):
        self._check(self.L.mfb_ccdpp_begin(self.h))

    def ccdpp_rank1(self, k, first_iter, inner=5, ureg=0.01, ireg=0.01, item_freq_thresh=0):
        self._check(self.L.mfb_ccdpp_rank1(self.h, k, int(first_iter), inner, ureg, ireg, item_freq_thresh))

    def ccd_half_step(self, side, reg, dim_order=None):
        """One half of a trainCCD epoch (modelMF.cpp:1527-1606); dim_order = uint8 [n rows of the side][rank] or None."""
        d = None
        if dim_order is not None:
            d = np.ascontiguousarray(dim_order, np.uint8)
            assert d.shape == ((self.n_users if side == USER else self.n_items), self.rank)
        self._check(self.L.mfb_ccd_half_step(self.h, side, reg, None if d is None else d.ctypes.data_as(C.c_void_p)))

    def debug_chol64(self, records, rank, reg):
        """Batched rank-64 solver on host records [n][2440] (see mfb_debug_chol64); returns x [n][64]."""
        rec = np.ascontiguousarray(records, np.float32)
        x = np.zeros((rec.shape[0], 64), np.float32)
        self._check(self.L.mfb_debug_chol64(self.h, rec.shape[0], rec.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), rank, reg))
        return x

    def ccdpp_end(self):
        self._check(self.L.mfb_ccdpp_end(self.h))

    # ---- evaluation ----
    def eval(self, which, factors=CURRENT, variant=MF, weighted=False, want_norms=False):
        out = np.zeros(4, np.float64)
        self._check(self.L.mfb_eval(self.h, which, factors, variant, int(weighted), int(want_norms), _p(out)))
        return out

    def eval_groups(self, which, user_group, item_group, factors=CURRENT, variant=MF):
        """(sse, count) per item group and per user group in one pass: returns array [2 sides][8 groups][2]."""
        ug, ig = _arr(user_group, np.uint8), _arr(item_group, np.uint8)
        assert ug.shape[0] == self.n_users and ig.shape[0] == self.n_items
        out = np.zeros(32, np.float64)
        self._check(self.L.mfb_eval_groups(self.h, which, factors, variant, _p(ug), _p(ig), _p(out)))
        return out.reshape(2, 8, 2)

    def rank_positions(self, which, factors=CURRENT, variant=MF):
        """(pos, test_item) per user: see mfb_rank_positions in include/mfb.h."""
        pos = np.zeros(self.n_users, np.int32)
        tst = np.zeros(self.n_users, np.int32)
        self._check(self.L.mfb_rank_positions(self.h, which, factors, variant, _p(pos), _p(tst)))
        return pos, tst

    def predict(self, which, nnz, factors=CURRENT, variant=MF):
        pred = np.zeros(max(int(nnz), 1), np.float32)
        self._check(self.L.mfb_predict(self.h, which, factors, variant, _p(pred)))
        return pred[: int(nnz)]

    def rmse(self, which, factors=CURRENT, variant=MF):
        o = self.eval(which, factors, variant)
        return float(np.sqrt(o[0] / o[1])) if o[1] > 0 else float("nan")

    def objective(self, ureg, ireg, variant=MF):
        o = self.eval(TRAIN, CURRENT, variant, weighted=(variant == IFWMF), want_norms=True)
        return float(o[0] + ureg * o[2] + ireg * o[3])

    def snapshot_best(self):
        self._check(self.L.mfb_snapshot_best(self.h))

    def restore_best(self):
        self._check(self.L.mfb_restore_best(self.h))

    # ---- timing / plumbing ----
    def sync(self):
        self._check(self.L.mfb_sync(self.h))

    def event_record(self, slot):
        self._check(self.L.mfb_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._check(self.L.mfb_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def device_factors(self, side):
        p, ld = C.c_void_p(), C.c_int64()
        self._check(self.L.mfb_device_factors(self.h, side, C.byref(p), C.byref(ld)))
        return p.value, ld.value

    def stream(self):
        return self.L.mfb_stream(self.h)

    def pack_rows(self, side, ids, dev_ptr):
        ids = _arr(ids, np.int32)
        self._check(self.L.mfb_pack_rows(self.h, side, _p(ids), ids.shape[0], C.c_void_p(dev_ptr)))

    def unpack_rows(self, side, ids, dev_ptr):
        ids = _arr(ids, np.int32)
        self._check(self.L.mfb_unpack_rows(self.h, side, _p(ids), ids.shape[0], C.c_void_p(dev_ptr)))

    def set_row_range(self, side, begin, end):
        self._check(self.L.mfb_set_row_range(self.h, side, begin, end))

    # ---- peer-memory exchange (one engine per rank of one node) ----
    def comm_init(self, rank, world) -> bytes:
        buf = (C.c_uint8 * 512)()
        n = C.c_int64()
        self._check(self.L.mfb_comm_init(self.h, rank, world, buf, C.byref(n)))
        return bytes(buf[: n.value])

    def comm_connect(self, blobs):
        """blobs: the handle blobs of all ranks in rank order (e.g. from all_gather_object)."""
        data = b"".join(blobs)
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        self._check(self.L.mfb_comm_connect(self.h, buf, len(data)))

    def comm_disconnect(self):
        self._check(self.L.mfb_comm_disconnect(self.h))

    def comm_barrier(self):
        self._check(self.L.mfb_comm_barrier(self.h))

    def comm_error(self) -> bool:
        v = C.c_int32()
        self._check(self.L.mfb_comm_error(self.h, C.byref(v)))
        return bool(v.value)

    def dsgd_push_block(self, item_part, dst_rank, seq):
        self._check(self.L.mfb_dsgd_push_block(self.h, item_part, dst_rank, seq))

    def comm_wait_block(self, src_rank, seq):
        self._check(self.L.mfb_comm_wait_block(self.h, src_rank, seq))

    def comm_allgather_rows(self, side, ids=None, first=0, n=0):
        ids = _arr(ids, np.int32)
        if ids is not None:
            n = ids.shape[0]
        self._check(self.L.mfb_comm_allgather_rows(self.h, side, _p(ids), first, n))


def connect_local(engines):
    """Connect engines of this process as ranks 0..n-1 of one exchange group (mfb_comm_connect_local)."""
    L = load_library()
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    if L.mfb_comm_connect_local(arr, len(engines)) != 0:
        raise EngineError(L.mfb_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(load_library().mfb_launch_count())


def pin_host(arr: np.ndarray):
    L = load_library()
    if L.mfb_pin_host(arr.ctypes.data_as(C.c_void_p), arr.nbytes) != 0:
        raise EngineError(L.mfb_last_error().decode(errors="replace"))


def unpin_host(arr: np.ndarray):
    load_library().mfb_unpin_host(arr.ctypes.data_as(C.c_void_p))
