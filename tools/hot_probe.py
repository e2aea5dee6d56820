"""Hot-row CTAs (sgd_hot_kernel) on the bench matrix (diagnostic): (1) one-launch epoch, hot on / off, several list
thresholds and CTA sizes, with the validation curve; (2) one-GPU emulation of the N = 8 DSGD bench (every
(g, sigma_t(g)) block of the balanced partition as its own launch; emulated parallel epoch = sum over sub-epochs
of the slowest block)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from matfac_b200 import engine as E, dsgd

n_users, n_items, nnz = bench.SHAPE
prob = bench.gen_problem(n_users, n_items, nnz, 20260102, "cuda:0")
ptr, ind, val = prob["train"]
train_nnz = int(ptr[-1])
rng = np.random.default_rng(1)
R = bench.RANK
U0 = rng.uniform(-0.01, 0.01, size=(n_users, R)).astype(np.float32)
V0 = rng.uniform(-0.01, 0.01, size=(n_items, R)).astype(np.float32)
eng = E.Engine(n_users, n_items, R)
eng.upload_csr(E.TRAIN, bench.Mat(n_users, n_items, prob["train"]), with_csc=False)
eng.upload_csr(E.VAL, bench.Mat(n_users, n_items, prob["val"]), with_csc=False)
eng.set_masks((np.diff(ptr) == 0).astype(np.uint8), (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8))
LR = 0.002
epochs = int(os.environ.get("EPOCHS", "8"))
which = os.environ.get("PARTS", "2")

if "1" in which:
    for hot, min_count, lists_max, batch, dbg, pace in ((0, 4096, 64, 0, 0, 1), (1, 4096, 64, 0, 0, 1), (1, 4096, 127, 0, 0, 1), (1, 4096, 127, 0, 0, 0), (1, 4096, 127, 16, 0, 1)):
        eng.set_option("sgd_hot", hot)
        eng.set_option("sgd_flat_debug", dbg)
        eng.set_option("sgd_hot_pace", pace)
        eng.set_option("sgd_hot_min_count", min_count)
        eng.set_option("sgd_hot_max_lists", lists_max)
        eng.set_option("sgd_hot_batch", batch)
        eng.sync(); t0 = time.perf_counter()
        eng.sgd_plan(1)
        eng.sync(); plan_ms = (time.perf_counter() - t0) * 1e3
        _, cold, lists = eng.debug_sgd_records(0, 0, with_records=False) if hot else (None, train_nnz, [])
        eng.upload_factors(U0, V0)
        ms, curve = [], []
        for ep in range(epochs):
            eng.event_record(0)
            eng.sgd_epoch_flat(E.MF, LR, 0.05, 0.05, 1, ep)
            eng.event_record(1)
            ms.append(eng.event_elapsed_ms(0, 1))
            curve.append(eng.rmse(E.VAL))
        eng.event_record(2)
        for ep in range(4):
            eng.sgd_epoch_flat(E.MF, LR, 0.05, 0.05, 1, epochs + ep)
        eng.event_record(3)
        b2b = eng.event_elapsed_ms(2, 3) / 4
        st = eng.debug_sgd_hot_batch()
        print(f"   mean|u|^2 {st[0] / max(st[1], 1):.3f} batch used {st[2]:.0f}")
        print(f"N=1 hot {hot} min_count {min_count:6d} max_lists {lists_max:3d} batch {batch:2d} debug {dbg:2d} pace {pace}: lists {len(lists):3d} hot share {1 - cold / train_nnz:.3f} "
              f"plan {plan_ms:6.1f} ms epoch {np.median(ms[1:]):7.3f} ms (back to back {b2b:7.3f})  val " + " ".join(f"{x:.4f}" for x in curve), flush=True)
    eng.set_option("sgd_hot_batch", 0); eng.set_option("sgd_hot_max_lists", 127); eng.set_option("sgd_hot_min_count", 4096)
    eng.set_option("sgd_flat_debug", 0); eng.set_option("sgd_hot_pace", 1)

if "2" in which:
    P = int(os.environ.get("P", "8"))
    user_part = dsgd.balanced_partition(np.diff(ptr), P)
    item_part = dsgd.balanced_partition(np.bincount(ind, minlength=n_items), P)
    sched = dsgd.rotation_schedule(P, epochs * P)
    for hot, min_count, warps, dbg, stages in ((0, 1024, 0, 0, 8), (1, 1024, 0, 0, 8), (1, 1024, 0, 0, 8), (1, 1024, 0, 32, 8), (1, 1024, 0, 16, 8), (1, 1024, 32, 0, 8)):
        eng.set_option("sgd_hot_stages", stages)
        eng.set_option("sgd_hot", hot)
        eng.set_option("sgd_flat_debug", dbg)
        eng.set_option("sgd_hot_min_count", min_count)
        eng.set_option("sgd_hot_batch", warps)
        eng.sync(); t0 = time.perf_counter()
        eng.sgd_plan(P, user_part, item_part)
        eng.sync(); plan_ms = (time.perf_counter() - t0) * 1e3
        eng.set_option("sgd_block_order", 1)
        nl = sum(len(eng.debug_sgd_records(a, b, with_records=False)[2]) for a in range(2) for b in range(P)) if hot else 0
        eng.upload_factors(U0, V0)
        curve, par_ms, tot_ms = [], [], []
        for ep in range(epochs):
            par = tot = 0.0
            for t in range(ep * P, (ep + 1) * P):
                worst = 0.0
                for g in range(P):
                    eng.event_record(0)
                    eng.sgd_subepoch(np.array([[g, sched[t, g]]], np.int32), E.MF, LR, 0.05, 0.05, 1, t)
                    eng.event_record(1)
                    m = eng.event_elapsed_ms(0, 1)
                    worst = max(worst, m); tot += m
                par += worst
            par_ms.append(par); tot_ms.append(tot)
            curve.append(eng.rmse(E.VAL))
        st = eng.debug_sgd_hot_batch()
        print(f"   mean|u|^2 {st[0] / max(st[1], 1):.3f} batch used {st[2]:.0f}")
        print(f"P={P} emulation hot {hot} min_count {min_count} batch {warps} debug {dbg} stages {stages}: lists in user parts 0-1: {nl}, plan {plan_ms:.1f} ms, emulated parallel "
              f"ms/epoch {np.median(par_ms[1:]):7.3f} (sum of all blocks {np.median(tot_ms[1:]):7.3f}) val " + " ".join(f"{x:.4f}" for x in curve), flush=True)
eng.close()
