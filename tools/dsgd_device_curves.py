#!/usr/bin/env python
"""Validation-RMSE curves of the device DSGD path for 1, 2, 4, 8 ranks on the 1/20-scale bench matrix — all ranks as
engines of ONE process on one GPU (mfb_comm_connect_local: the same kernels, flags and pushes as one process per GPU) —
next to the oracle curves of tools/dsgd_oracle_curves.py (profiles/r2_dsgd_oracle_curves.json).

usage: python tools/dsgd_device_curves.py [--epochs 25] [--out gpurun_out/dsgd_device_curves.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matfac_b200 import dsgd, synth  # noqa: E402
from matfac_b200 import engine as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=25)
    ap.add_argument("--worlds", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--shape", default="netflix", choices=["netflix", "yahoo"])
    ap.add_argument("--seeds", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--relax_after", type=int, default=-1, help="epochs after which the in-flight budget becomes --inflight_late")
    ap.add_argument("--inflight_late", type=float, default=8e-4)
    ap.add_argument("--lr", type=float, default=0.002)
    ap.add_argument("--inflight", type=float, default=2e-4)
    ap.add_argument("--configs", nargs="+", default=["reference:1", "reference:0", "balanced:1"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "dsgd_device_curves.json"))
    a = ap.parse_args()
    import torch
    ndev = torch.cuda.device_count()
    import bench
    if a.shape == "netflix":
        nu, ni, nnz = synth.SHAPES["netflix"]
        nu, ni, nnz = int(nu * a.scale), int(ni * max(a.scale, 0.05)), int(nnz * a.scale)
        prob = synth.skewed_problem(nu, ni, nnz, 20260102, device="cuda:0")
    else:  # the bench's Yahoo-R1-shaped block (BASELINE.json configs[4])
        prob = bench.make_problem(bench.YAHOO_SHAPE, a.scale, "cuda:0")
        nu, ni = prob["n_users"], prob["n_items"]
    ptr, ind, _ = prob["train"]
    bad_u = (np.diff(ptr) == 0).astype(np.uint8)
    bad_i = (np.bincount(ind, minlength=ni) == 0).astype(np.uint8)
    rng = np.random.default_rng(1)
    U0 = rng.uniform(-0.01, 0.01, (nu, 64)).astype(np.float32)
    V0 = rng.uniform(-0.01, 0.01, (ni, 64)).astype(np.float32)
    out = {"matrix": {"n_users": nu, "n_items": ni, "train_nnz": int(ptr[-1]), "crc": prob["crc"]}, "lr": a.lr,
           "inflight_frac": a.inflight, "curves": {}, "ms_per_epoch": {}}
    for cfg in a.configs:
        parts = cfg.split(":")  # plan:block_order[:option=value[,option=value...]]
        plan, order = parts[0], parts[1]
        extra = dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in parts[2].split(",")) if len(parts) > 2 else {}
        for world in a.worlds:
            for seed in a.seeds:
                t0 = time.time()
                d = dsgd.Dsgd(nu, ni, 64, world, {r: r % ndev for r in range(world)}, prob["train"], prob["val"], U0, V0, bad_u,
                              bad_i, a.epochs * world, plan=plan, seed=seed, block_order=int(order),
                              options=dict({"sgd_flat_inflight_frac": a.inflight}, **extra))
                curve, ms = [], []
                e0 = d.engines[0]
                for ep in range(a.epochs):
                    if ep == a.relax_after:
                        for e in d.engines.values():
                            e.set_option("sgd_flat_inflight_frac", a.inflight_late)
                    e0.event_record(0)
                    d.run(ep * world, (ep + 1) * world, a.lr, 0.05, 0.05, seed)
                    e0.event_record(1)
                    d.publish()
                    s = d.eval_sums(E.VAL)
                    d.barrier()
                    curve.append(float(np.sqrt(s[0] / s[1])))
                    ms.append(e0.event_elapsed_ms(0, 1))
                err = any(e.comm_error() for e in d.engines.values()) if world > 1 else False
                d.close()
                key = f"{cfg}_N{world}_seed{seed}"
                out["curves"][key] = curve
                out["ms_per_epoch"][key] = float(np.median(ms))
                print(key, f"{time.time()-t0:.0f}s", f"{np.median(ms):.2f}ms", "COMM_ERROR" if err else "",
                      " ".join(f"{v:.4f}" for v in curve), flush=True)
                json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
