"""Synthetic rating matrices shaped like the reference's datasets.

Modelled on the reference's own fixture recipe ``writeRandMatCSR`` (io.cpp:726-787): ratings
come from known low-rank factors, every user and every item gets at least one rating, items are
sorted ascending inside a row (``checkIfUISorted``, util.cpp:919).  Positions are Zipf-skewed on
both axes because IFWMF / TMF key on frequency skew.  The split into train / val / test colours
each non-zero independently (as io.cpp:410-459 does) and every split keeps exactly ``n_users``
rows, because ``Model::RMSE`` indexes ``rowptr[u]`` for all ``u < nUsers`` (model.cpp:223,231).

Everything here is host-side numpy: it produces inputs, it is not on the training path.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

SHAPES = {
    # name: (n_users, n_items, nnz)     BASELINE.json configs
    "ml1m": (6040, 3706, 1_000_209),
    "netflix": (480_189, 17_770, 100_480_507),
    "ml20m": (138_493, 26_744, 20_000_263),
    "yahoo_r1": (1_000_990, 624_961, 250_000_000),
}


@dataclasses.dataclass
class Csr:
    """Host CSR + CSC of one rating matrix (the reference's ``gk_csr_t`` fields)."""

    nrows: int
    ncols: int
    rowptr: np.ndarray  # int64 [nrows+1]
    rowind: np.ndarray  # int32 [nnz]
    rowval: np.ndarray  # float32 [nnz]
    colptr: np.ndarray | None = None  # int64 [ncols+1]
    colind: np.ndarray | None = None  # int32 [nnz]
    colval: np.ndarray | None = None  # float32 [nnz]

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])

    def build_csc(self) -> "Csr":
        """Stable counting sort by column == ``gk_csr_CreateIndex(mat, GK_CSR_COL)``."""
        nnz = self.nnz
        rows = np.repeat(np.arange(self.nrows, dtype=np.int32), np.diff(self.rowptr))
        order = np.argsort(self.rowind, kind="stable")
        counts = np.bincount(self.rowind, minlength=self.ncols).astype(np.int64)
        self.colptr = np.zeros(self.ncols + 1, dtype=np.int64)
        np.cumsum(counts, out=self.colptr[1:])
        self.colind = rows[order].astype(np.int32)
        self.colval = self.rowval[order].astype(np.float32)
        assert self.colind.shape[0] == nnz
        return self


def _zipf_weights(n: int, s: float, rng: np.random.Generator) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    rng.shuffle(w)  # popularity is not correlated with the id
    return w / w.sum()


def _sample_positions(n_users, n_items, nnz, rng, user_s, item_s):
    """Return unique (user, item) pairs, Zipf-skewed, covering every user and item once."""
    pu = _zipf_weights(n_users, user_s, rng)
    pi = _zipf_weights(n_items, item_s, rng)
    cu = np.cumsum(pu)
    ci = np.cumsum(pi)
    keys = np.empty(0, dtype=np.int64)
    want = nnz
    # cover every user / item at least once (io.cpp:742-752)
    base_u = np.arange(n_users, dtype=np.int64)
    base_i = np.minimum(np.searchsorted(ci, rng.random(n_users)), n_items - 1).astype(np.int64)
    base2_i = np.arange(n_items, dtype=np.int64)
    base2_u = np.minimum(np.searchsorted(cu, rng.random(n_items)), n_users - 1).astype(np.int64)
    keys = np.unique(np.concatenate([base_u * n_items + base_i, base2_u * n_items + base2_i]))
    max_cells = n_users * n_items
    want = min(want, max_cells)
    rounds = 0
    while keys.shape[0] < want and rounds < 64:
        need = want - keys.shape[0]
        draw = int(need * 1.3) + 1024
        u = np.minimum(np.searchsorted(cu, rng.random(draw)), n_users - 1).astype(np.int64)
        i = np.minimum(np.searchsorted(ci, rng.random(draw)), n_items - 1).astype(np.int64)
        keys = np.unique(np.concatenate([keys, u * n_items + i]))
        rounds += 1
    if keys.shape[0] > want:
        # drop a random surplus but never the covering pairs of a user's/item's only rating
        drop = rng.choice(keys.shape[0], size=keys.shape[0] - want, replace=False)
        mask = np.ones(keys.shape[0], dtype=bool)
        mask[drop] = False
        keys = keys[mask]
    users = (keys // n_items).astype(np.int32)
    items = (keys % n_items).astype(np.int32)
    return users, items


def make_ratings(n_users, n_items, nnz, seed=20260101, true_rank=8, user_s=0.9, item_s=1.05,
                 noise=0.3):
    """(users, items, vals) sorted by (user, item); vals in {1.0, 1.5, ..., 5.0}."""
    rng = np.random.default_rng(seed)
    users, items = _sample_positions(n_users, n_items, nnz, rng, user_s, item_s)
    ustar = rng.normal(0.0, 1.0, size=(n_users, true_rank)).astype(np.float32)
    vstar = rng.normal(0.0, 1.0, size=(n_items, true_rank)).astype(np.float32)
    vals = np.empty(users.shape[0], dtype=np.float32)
    step = 1 << 22
    for s in range(0, users.shape[0], step):
        e = min(s + step, users.shape[0])
        d = np.einsum("ij,ij->i", ustar[users[s:e]], vstar[items[s:e]]) / np.sqrt(true_rank)
        vals[s:e] = 3.6 + 1.1 * d + rng.normal(0.0, noise, size=e - s)
    vals = np.clip(np.round(vals * 2.0) / 2.0, 1.0, 5.0).astype(np.float32)
    return users, items, vals


def coo_to_csr(users, items, vals, nrows, ncols=None) -> Csr:
    counts = np.bincount(users, minlength=nrows).astype(np.int64)
    rowptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    if ncols is None:
        ncols = int(items.max()) + 1 if items.shape[0] else 0
    return Csr(nrows, ncols, rowptr, items.astype(np.int32), vals.astype(np.float32))


def make_splits(n_users, n_items, nnz, seed=20260101, fracs=(0.8, 0.1, 0.1), **kw):
    """Return (train, val, test) ``Csr`` objects with CSC built; ncols = max index + 1 per file
    (what ``gk_csr_Read`` derives), exactly ``n_users`` rows each."""
    users, items, vals = make_ratings(n_users, n_items, nnz, seed=seed, **kw)
    rng = np.random.default_rng(seed + 7)
    colour = rng.random(users.shape[0])
    # keep each user's / item's first rating in train so that the training matrix spans the ids
    first_u = np.ones(users.shape[0], dtype=bool)
    first_u[1:] = users[1:] != users[:-1]
    order_i = np.argsort(items, kind="stable")
    first_i = np.zeros(users.shape[0], dtype=bool)
    si = items[order_i]
    fi = np.ones(si.shape[0], dtype=bool)
    fi[1:] = si[1:] != si[:-1]
    first_i[order_i[fi]] = True
    colour[first_u | first_i] = 0.0
    tr = colour < fracs[0]
    va = (~tr) & (colour < fracs[0] + fracs[1])
    te = ~(tr | va)
    out = []
    for m in (tr, va, te):
        out.append(coo_to_csr(users[m], items[m], vals[m], n_users).build_csc())
    return tuple(out)


def make_ranking_splits(n_users, n_items, nnz, seed=20260101, **kw):
    """Leave-one-out style splits for the ranking metrics (model.cpp:981-1332 read the FIRST rating of every user's
    row in the validation / test matrix unconditionally): every user holds at least one validation and one test
    rating, none of them among its training items; the rest of the ratings go to train."""
    users, items, vals = make_ratings(n_users, n_items, nnz, seed=seed, **kw)
    rng = np.random.default_rng(seed + 11)
    key = rng.random(users.shape[0])
    order = np.lexsort((key, users))  # by user, random inside a user
    users, items, vals = users[order], items[order], vals[order]
    first = np.ones(users.shape[0], dtype=bool)
    first[1:] = users[1:] != users[:-1]
    pos = np.arange(users.shape[0]) - np.maximum.accumulate(np.where(first, np.arange(users.shape[0]), 0))
    deg = np.bincount(users, minlength=n_users)
    assert deg.min() >= 3, "every user needs at least three ratings"
    extra = rng.random(users.shape[0])
    te = (pos == 0) | ((pos >= 3) & (extra < 0.05))
    va = (pos == 1) | ((pos >= 3) & (extra >= 0.05) & (extra < 0.10))
    tr = ~(te | va)
    out = []
    for m in (tr, va, te):
        out.append(coo_to_csr(users[m], items[m], vals[m], n_users).build_csc())
    return tuple(out)


def write_text_csr(mat: Csr, path: str) -> None:
    """The reference's input format: one line per user, ``item rating item rating ...``,
    0-indexed (python/convert_scipy_sparse_to_text_csr.py:19-26; datastruct.cpp:16)."""
    with open(path, "w") as f:
        for u in range(mat.nrows):
            s, e = int(mat.rowptr[u]), int(mat.rowptr[u + 1])
            f.write(" ".join(f"{int(i)} {float(v):g}" for i, v in
                             zip(mat.rowind[s:e], mat.rowval[s:e])))
            f.write("\n")


def write_split_files(dirname: str, train: Csr, val: Csr, test: Csr):
    os.makedirs(dirname, exist_ok=True)
    paths = [os.path.join(dirname, n) for n in ("train.csr", "val.csr", "test.csr")]
    for m, p in zip((train, val, test), paths):
        write_text_csr(m, p)
    return paths


# ---------------------------------------------------------------------------------------------------------
# Bit-reproducible large problems (bench.py, the multi-GPU tools and the scale tests).
#
# Everything below is a pure function of (shape, seed): draws come from a counter-based hash (splitmix64 of
# (seed, stream, index)), de-duplication is sort-based, the surplus is trimmed by hash rank, and the rating values
# use only IEEE add / multiply on exactly representable inputs — no atomics, no library RNG stream, no libm call on
# the data path.  The same call therefore returns the same arrays in every process, on every rank and on "cpu" and
# "cuda" alike (tests/test_synth.py; bench.py prints the CRC of the arrays so that two runs can be compared).
_M64 = (1 << 64) - 1


def _i64(x: int) -> int:
    """Python int (mod 2^64) as the two's-complement int64 torch stores."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix64_t(x):
    """splitmix64 finaliser on an int64 tensor (wrap-around arithmetic)."""
    x = x + _i64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _i64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _mix64_py(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def _stream_key(seed: int, stream: int) -> int:
    return _i64(_mix64_py(_mix64_py(seed) ^ (stream * 0xD1B54A32D192ED03 & _M64)))


def _hash_t(idx, seed, stream):
    """64 hashed bits per element of the int64 tensor idx."""
    return _mix64_t((idx * _i64(0xA24BAED4963EE407)) ^ _stream_key(seed, stream))


def _uniform53(h):
    return _lsr(h, 11).double() * (1.0 / 9007199254740992.0)


def _normalish(idx, seed, stream):
    """Sum of twelve 24-bit uniforms minus 6: mean 0, variance 1, every partial sum exact in float64."""
    acc = None
    for k in range(6):
        h = _hash_t(idx, seed, stream * 16 + k)
        a = (_lsr(h, 40)).double() * (1.0 / 16777216.0)
        b = ((h >> 8) & 0xFFFFFF).double() * (1.0 / 16777216.0)
        acc = a + b if acc is None else acc + a + b
    return acc - 6.0


def _perm_weights(n, s, seed, stream, cap=None):
    """Zipf(s) weights over n ids in a hashed (id-uncorrelated) order, optionally with the head capped."""
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    ids = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = np.uint64(_mix64_py(_mix64_py(seed) ^ (stream * 0xD1B54A32D192ED03 & _M64)))
        h = ids * np.uint64(0xA24BAED4963EE407) ^ k
        h = h + np.uint64(0x9E3779B97F4A7C15)
        h = (h ^ (h >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        h = (h ^ (h >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        h = h ^ (h >> np.uint64(31))
    order = np.argsort(h, kind="stable")
    p = np.empty(n, dtype=np.float64)
    p[order] = w  # id order[k] gets the k-th largest weight
    p /= p.sum()
    if cap is not None:
        for _ in range(8):
            p = np.minimum(p, cap)
            p /= p.sum()
    return p


NETFLIX_TOP_ITEM_SHARE = 232_944 / 100_480_507  # the most rated title of the real Netflix Prize data


def array_crc(*arrays) -> str:
    """CRC-32 of the raw bytes of the arrays, chained — the fingerprint bench.py prints for its input matrix."""
    import zlib
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).view(np.uint8).reshape(-1), c)
    return f"{c:08x}"


def skewed_problem(n_users, n_items, nnz, seed, device="cpu", val_frac=0.01, user_s=0.9, item_s=1.05,
                   item_share_cap=NETFLIX_TOP_ITEM_SHARE, true_rank=8, noise=0.3, max_user_share_of_items=0.85):
    """Netflix-/Yahoo-shaped rating matrix: Zipf(user_s) user degrees (capped at 85 % of the catalogue), Zipf(item_s)
    item popularity with the head capped at item_share_cap, every user and item rated at least once (io.cpp:742-752),
    items ascending inside a row (util.cpp:919), ratings from a rank-`true_rank` model + noise rounded to halves in
    [1, 5], a `val_frac` validation split that never takes a user's or an item's first rating.

    Returns {"n_users", "n_items", "train": (rowptr int64, rowind int32, rowval fp32), "val": (...), "crc": str} with
    host numpy arrays.  Deterministic in (arguments) — see the note above."""
    import torch
    dev = torch.device(device)
    pu = _perm_weights(n_users, user_s, seed, 1)
    pi = _perm_weights(n_items, item_s, seed, 2, cap=item_share_cap)
    ci = torch.from_numpy(np.cumsum(pi)).to(dev)
    cu = torch.from_numpy(np.cumsum(pu)).to(dev)
    total = int(nnz * (1.0 + val_frac))
    total = min(total, int(0.5 * n_users * n_items))
    ar_u = torch.arange(n_users, device=dev, dtype=torch.int64)
    ar_i = torch.arange(n_items, device=dev, dtype=torch.int64)
    # every user and every item at least once
    base_i = torch.searchsorted(ci, _uniform53(_hash_t(ar_u, seed, 3))).clamp_(max=n_items - 1)
    base2_u = torch.searchsorted(cu, _uniform53(_hash_t(ar_i, seed, 4))).clamp_(max=n_users - 1)
    cover = torch.unique(torch.cat([ar_u * n_items + base_i, base2_u * n_items + ar_i]))
    keys = cover
    cap = max_user_share_of_items * n_items
    for rnd in range(16):
        need = total - keys.numel()
        if need <= 0:
            break
        draw = int(need * (1.6 if rnd == 0 else 1.3)) + 1024
        deg_np = np.rint(np.minimum(pu * draw, cap)).astype(np.int64)
        deg = torch.from_numpy(deg_np).to(dev)
        u = torch.repeat_interleave(ar_u, deg)
        idx = torch.arange(u.numel(), device=dev, dtype=torch.int64)
        i = torch.searchsorted(ci, _uniform53(_hash_t(idx, seed, 16 + rnd))).clamp_(max=n_items - 1)
        keys = torch.unique(torch.cat([keys, u * n_items + i]))
        del u, i, idx, deg
    if keys.numel() > total:
        # keep the `total` keys of smallest hash; the covering pairs always survive
        h = _lsr(_hash_t(keys, seed, 5), 1)  # non-negative
        h[torch.isin(keys, cover)] = -1
        order = torch.sort(h, stable=True).indices[:total]
        keys = torch.sort(keys[order]).values
        del h, order
    users = keys // n_items
    items = keys - users * n_items
    # ratings: exact-arithmetic "normal" factors and noise, float64 dot with a fixed summation order
    us = torch.stack([_normalish(ar_u, seed, 32 + k) for k in range(true_rank)], 1)
    vs = torch.stack([_normalish(ar_i, seed, 64 + k) for k in range(true_rank)], 1)
    scale = 1.1 / float(np.sqrt(true_rank))
    vals = torch.empty(users.numel(), dtype=torch.float32, device=dev)
    step = 1 << 24
    for s in range(0, users.numel(), step):
        e = min(s + step, users.numel())
        uu, ii = us[users[s:e]], vs[items[s:e]]
        d = uu[:, 0] * ii[:, 0]
        for k in range(1, true_rank):
            d = d + uu[:, k] * ii[:, k]
        v = d * scale + 3.6
        v = v + _normalish(keys[s:e], seed, 7) * noise
        vals[s:e] = (torch.round(v * 2.0) * 0.5).clamp_(1.0, 5.0).float()
        del uu, ii, d, v
    # split: a hashed colour per (user, item); first rating of every user / item stays in train
    colour = _uniform53(_hash_t(keys, seed, 6))
    first_u = torch.ones(users.numel(), dtype=torch.bool, device=dev)
    first_u[1:] = users[1:] != users[:-1]
    order_i = torch.sort(items * n_users + users).indices  # keys are unique: no ties
    si = items[order_i]
    fi = torch.ones(si.numel(), dtype=torch.bool, device=dev)
    fi[1:] = si[1:] != si[:-1]
    first_i = torch.zeros(users.numel(), dtype=torch.bool, device=dev)
    first_i[order_i[fi]] = True
    is_val = (colour < val_frac) & ~first_u & ~first_i
    del order_i, si, fi, colour, keys

    def csr(mask):
        u, i, v = users[mask], items[mask], vals[mask]
        ptr = torch.zeros(n_users + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(torch.bincount(u, minlength=n_users), 0)
        return ptr.cpu().numpy(), i.to(torch.int32).cpu().numpy(), v.cpu().numpy()

    tr, va = csr(~is_val), csr(is_val)
    return dict(n_users=n_users, n_items=n_items, train=tr, val=va, crc=array_crc(*tr, *va))
