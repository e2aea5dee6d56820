#!/usr/bin/env python
"""Bitwise repeatability of an ALS half-step (diagnostic): the warp-specialised kernel sums every row in a fixed order, so
repeated launches from the same inputs must give identical bits.  usage: python tools/als_determinism.py [rank ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import rel_err  # noqa: E402
from gpu_driver import make_engine  # noqa: E402
from matfac_b200 import engine as E, synth  # noqa: E402
import oracle_lib as ol  # noqa: E402

for rank in [int(x) for x in sys.argv[1:]] or [64, 128]:
    for shape in ((3000, 1500, 300000), (900, 600, 380000)):
        splits = synth.make_splits(*shape, seed=21)
        od = ol.OracleData(*splits)
        om = ol.OracleModel(od, algo="mf", facdim=rank, maxiter=1, seed=3, nthreads=4, ureg=0.1, ireg=0.1)
        eng, _ = make_engine(splits, om, rank)
        rng = np.random.default_rng(1)
        U0, V0 = om.factors()
        U0 = rng.standard_normal(U0.shape).astype(np.float32)
        V0 = rng.standard_normal(V0.shape).astype(np.float32)
        for split in (1, 2):
            eng.set_option("als_ws_split", split)
            outs = []
            for rep in range(6):
                eng.upload_factors(U0, V0)
                eng.als_half_step(E.USER, 0.1)
                eng.als_half_step(E.ITEM, 0.1)
                outs.append(eng.download_factors())
            nb = [(int((outs[i][0] != outs[0][0]).sum()), int((outs[i][1] != outs[0][1]).sum())) for i in range(1, 6)]
            eng.set_option("als_tensor_cores", 0)
            eng.upload_factors(U0, V0)
            eng.als_half_step(E.USER, 0.1)
            eng.als_half_step(E.ITEM, 0.1)
            Ur, Vr = eng.download_factors()
            eng.set_option("als_tensor_cores", 1)
            print(f"rank {rank} shape {shape} split {split}: differing elements vs run 0 {nb}; vs CUDA-core Gram rel err U {rel_err(outs[0][0], Ur):.2e} V {rel_err(outs[0][1], Vr):.2e}", flush=True)
        eng.close()
