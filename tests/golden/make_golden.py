"""Writes tests/golden/ref_*.npz from the reference's own translation units (oracle/_ref/mf_ref).
Run in the build container, where /root/reference exists:  python tests/golden/make_golden.py
The problem is synth.make_splits(300, 200, 6000, seed=5); flags are test_oracle.BASE + CASES."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol  # noqa: E402
from matfac_b200 import synth  # noqa: E402
import json  # noqa: E402

from test_oracle import BASE, CASES, RANK_BASE, RANK_CASES, golden_problem, ranking_problem  # noqa: E402

assert ol.have_ref(), "build oracle/_ref first: make -C oracle ref"
d = tempfile.mkdtemp()
files = synth.write_split_files(d, *golden_problem())
for algo, method, threads, extra in CASES:
    fl = dict(BASE); fl.update(extra)
    ref = ol.run_ref(files, os.path.join(d, f"dump_{algo}_{method}"), algo=algo, method=method, threads=threads, **fl)
    out = os.path.join(HERE, f"ref_{algo}_{method.replace('+', 'p')}.npz")
    np.savez_compressed(out, **{k: v for k, v in ref.items() if k != "stdout" and k != "signature"})
    print("wrote", out, os.path.getsize(out))

# ranking metrics of the reference's own model.cpp:760-1332 (ref_driver --rank_metrics 1) on the leave-one-out problem
d2 = tempfile.mkdtemp()
files = synth.write_split_files(d2, *ranking_problem())
for algo, method, threads, extra in RANK_CASES:
    fl = dict(RANK_BASE); fl.update(extra)
    ref = ol.run_ref(files, os.path.join(d2, f"dump_{algo}_{method}"), algo=algo, method=method, threads=threads, rank_metrics=1, **fl)
    keys = [k for k in ref if k.startswith(("val_", "test_"))]
    out = os.path.join(HERE, f"rank_{algo}_{method}.json")
    json.dump({k: ref[k] for k in sorted(keys)}, open(out, "w"), indent=1)
    print("wrote", out, {k: round(ref[k], 4) for k in ("val_hr", "val_arhr", "val_ndcg", "test_hr")})
