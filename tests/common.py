"""Shared fixtures for the parity tests: small seeded rating matrices and helpers that drive the
engine (through the C ABI) the way the host trainers do."""
from __future__ import annotations

import numpy as np

from matfac_b200 import synth

_cache = {}


def small_problem(n_users=600, n_items=400, nnz=30000, seed=11):
    key = (n_users, n_items, nnz, seed)
    if key not in _cache:
        _cache[key] = synth.make_splits(n_users, n_items, nnz, seed=seed)
    return _cache[key]


def disjoint_problem(n_users=257, per_user=9, seed=3):
    """Every item is rated by exactly one user: no two runs of the SGD kernel share an item row,
    so the device result does not depend on scheduling and can be compared value-for-value."""
    rng = np.random.default_rng(seed)
    n_items = n_users * per_user
    users = np.repeat(np.arange(n_users, dtype=np.int32), per_user)
    items = np.arange(n_items, dtype=np.int32)
    vals = (np.round(rng.uniform(1, 5, size=n_items) * 2) / 2).astype(np.float32)
    tr = synth.coo_to_csr(users, items, vals, n_users).build_csc()
    # val/test: one rating per user on one of its own items
    vi = (np.arange(n_users) * per_user).astype(np.int32)
    vv = (np.round(rng.uniform(1, 5, size=n_users) * 2) / 2).astype(np.float32)
    va = synth.coo_to_csr(np.arange(n_users, dtype=np.int32), vi, vv, n_users).build_csc()
    te = synth.coo_to_csr(np.arange(n_users, dtype=np.int32), vi + 1, vv, n_users).build_csc()
    return tr, va, te


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
