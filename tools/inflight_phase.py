"""Diagnostic (ML-20M shape, MF and IFWMF, lr 0.005): does the in-flight budget of the shuffled kernel only matter while the
factors grow?  Strict budget (2e-4) throughout / relaxed (1e-3) throughout / strict for the first K epochs, relaxed after."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E, synth
n_users, n_items, nnz = synth.SHAPES["ml20m"]
R = 64
prob = bench.gen_problem(n_users, n_items, nnz, 20260104, "cuda:0")
ptr, ind, val = prob["train"]
rng = np.random.default_rng(1)
U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
eng = E.Engine(n_users, n_items, R)
eng.upload_csr(E.TRAIN, bench.Mat(n_users, n_items, prob["train"]), with_csc=False)
eng.upload_csr(E.VAL, bench.Mat(n_users, n_items, prob["val"]), with_csc=False)
ufreq = np.diff(ptr).astype(np.int32); ifreq = np.bincount(ind, minlength=n_items).astype(np.int32)
eng.set_masks((ufreq == 0).astype(np.uint8), (ifreq == 0).astype(np.uint8))
rho = 1e5
pu = ufreq / n_items; pu = pu / pu.sum(); pi = ifreq / n_users; pi = pi / pi.sum()
wu = (1.0 / (1.0 + rho * pu)).astype(np.float32); wi = (1.0 / (1.0 + rho * pi)).astype(np.float32)
epochs = 12
for name, variant in (("MF", E.MF), ("IFWMF", E.IFWMF)):
    if variant == E.MF: eng.set_aux(E.MF, ufreq, ifreq)
    else: eng.set_aux(variant, ufreq, ifreq, wu, wi)
    base = None
    for label, switch in (("strict", 99), ("relaxed", 0), ("strict 1 epoch", 1), ("strict 2 epochs", 2), ("strict 3 epochs", 3)):
        eng.set_option("sgd_flat_inflight_frac", 2e-4 if switch > 0 else 1e-3)
        eng.upload_factors(U0, V0)
        eng.sgd_plan(1)
        curve, ms = [], []
        for ep in range(epochs):
            if ep == switch: eng.set_option("sgd_flat_inflight_frac", 1e-3)
            eng.event_record(0); eng.sgd_epoch_flat(variant, 0.005, 0.05, 0.05, 1, ep); eng.event_record(1)
            ms.append(eng.event_elapsed_ms(0, 1)); curve.append(eng.rmse(E.VAL, E.CURRENT, variant))
        curve = np.array(curve)
        if base is None: base = curve
        print(f"{name:6s} {label:16s} ms/epoch first {ms[0]:.2f} last {ms[-1]:.2f}  val " + " ".join(f"{x:.4f}" for x in curve) +
              "   max |rel dev| after epoch 2: " + f"{np.max(np.abs(curve[3:] / base[3:] - 1)) * 100:.2f} %", flush=True)
