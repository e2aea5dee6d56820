#!/usr/bin/env python
"""Diagnostic: row-sharded ALS (several engines, peer stores) against one engine, half-step by half-step, bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_multi_rank as T  # noqa: E402
from matfac_b200 import engine as E  # noqa: E402

r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
fracs = (0.7, 0.1, 0.1, 0.1)
tr, va, n_users, n_items, bad_u, bad_i = T._problem()
one = T._sharded_group(E, tr, va, n_users, n_items, r, (1.0,), bad_u, bad_i)
sh = T._sharded_group(E, tr, va, n_users, n_items, r, fracs, bad_u, bad_i)
lens_u = np.diff(tr.rowptr)
lens_i = np.bincount(tr.rowind, minlength=n_items)
for ep in range(2):
    for side, name, lens in ((E.USER, "U", lens_u), (E.ITEM, "V", lens_i)):
        for eng in one + sh:
            eng.als_half_step(side, 0.1)
        F1 = one[0].download_factors()[0 if side == E.USER else 1]
        for k, eng in enumerate(sh):
            F2 = eng.download_factors()[0 if side == E.USER else 1]
            bad = np.where((F1 != F2).any(axis=1))[0]
            rel = np.abs(F1 - F2).max() / max(np.abs(F1).max(), 1e-30)
            print(f"epoch {ep} {name} rank {k}: {len(bad)} rows differ (max abs diff / max {rel:.2e}); lengths of the first few: "
                  f"{[(int(b), int(lens[b])) for b in bad[:8]]}", flush=True)
        # continue from the single engine's factors so that every half-step is compared from identical inputs
        U1, V1 = one[0].download_factors()
        for eng in sh:
            eng.upload_factors(U1, V1)
T._close(one); T._close(sh)
