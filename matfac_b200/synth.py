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
