#!/usr/bin/env python
"""bench.py — SGD rating-updates/s on a Netflix-shaped synthetic problem at rank 64 (+ the other trainers of the path).

Contract (one JSON line on stdout, printed by rank 0):
  python bench.py --gpus N --steps K --warmup W          device arm (this repo's CUDA engine)
  python bench.py --impl reference --gpus N ...          reference arm: the reference's own CPU
                                                         code (oracle/_ref/mf_ref) on host cores
A "step" is one SGD epoch over the whole training matrix: at N = 1 the serial-SGD trainer
(ModelMF::train, the CLI default --mf_method sgd) on the shuffled kernel; at N > 1 the stratified
trainer (ModelMF::trainSGDPar = DSGD: user strata pinned to ranks, item blocks handed from rank to rank
after every sub-epoch) with the REFERENCE'S partitions and update sequences (modelMF.cpp:229-265,
util.cpp:1077-1107); the rating-balanced / rotation variant is timed next to it as `dsgd_balanced`.

Workload = BASELINE.json configs[1]: modelMF SGD, rank 64, 480,189 x 17,770, ~100.5 M ratings
(matfac_b200/synth.py: skewed_problem — a pure function of (shape, seed): both arms, every N and every rank
see the same matrix; `config.matrix_crc` is the CRC-32 of its arrays).  Inputs (ratings 0.8 GB + U 123 MB)
exceed the 126 MB L2, so no explicit L2 flush is done between timed epochs.

  value     ratings visited in the timed epochs / device time (CUDA events on the engine's stream, max over
            ranks), inputs resident in HBM
  e2e       the same metric through the C ABI with HOST buffers (N > 1: dsgd_e2e, every rank uploads its user stratum
            and the factors every step; N = 1): every step uploads the rating CSR and the
            factor matrices from pinned host memory, builds the plan, runs one epoch plus the per-epoch
            evaluation and downloads the factors.  Headline: two such steps in flight (two engines, one host
            thread each, taking turns on the link and on the SMs) = throughput of complete steps;
            e2e_one_in_flight = the latency of a single step; e2e_resident_ratings = the step of a training
            run (ratings uploaded once)
  roofline  algorithmic bytes (16 r + 12 per update) / kernel time vs the measured HBM copy
            bandwidth of MEASURED_PEAKS.json; `l2` = the same against the measured L2 figures of
            profiles/r2_l2_probe.json (the factor matrices of this shape are L2-resident)
  cpu_baseline  the reference's OpenMP stratified SGD (trainSGDPar) on a row sample, host cores
  solvers   ALS epoch seconds at rank 64 and 128, CCD++ rank-one step, objective pass (row-sharded at N > 1)
  yahoo     BASELINE.json configs[4] (1 M x 625 k, 250 M ratings, rank 64): SGD / DSGD epoch and CCD++ step — the
            shape whose factors (416 MB) do not fit L2
  oracle_check (N > 1)  the DSGD path at N ranks against the oracle's trainSGDPar with P = N on the 1/20-scale
            matrix: same partitions, same update sequences, validation RMSE per epoch side by side
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RANK = 64
SHAPE = (480_189, 17_770, 100_480_507)          # BASELINE.json configs[1] / [2]
YAHOO_SHAPE = (1_000_990, 624_961, 250_000_000)  # BASELINE.json configs[4]
DATA_SEED = 20260102
# learnrate 0.002: at 0.005 the reference itself diverges on this matrix in epoch 0 and falls back on
# its NaN guard (restore + halve, model.cpp:1487-1498) — measured with the oracle, see DESIGN.md
HP = dict(lr=0.002, ureg=0.05, ireg=0.05)
FALLBACK_HBM_GBS = 6650.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
def make_problem(shape, scale, device):
    """The bench matrix (or a `scale`d copy: users and ratings scaled, items scaled down to no less than 5 %)."""
    from matfac_b200 import synth
    n_users = int(shape[0] * scale)
    n_items = int(shape[1] * max(scale, 0.05)) if scale < 1 else shape[1]
    nnz = int(shape[2] * scale)
    return synth.skewed_problem(n_users, n_items, nnz, DATA_SEED, device=device)


def gen_problem(n_users, n_items, nnz, seed, device):
    """The generator under its round-1 name (tools/): same arrays as matfac_b200.synth.skewed_problem."""
    from matfac_b200 import synth
    return synth.skewed_problem(n_users, n_items, nnz, seed, device=device)


class Mat:
    def __init__(self, nrows, ncols, t):
        self.nrows, self.ncols = nrows, ncols
        self.rowptr, self.rowind, self.rowval = t
        self.colptr = self.colind = self.colval = None


def pinned(a):
    """Copy a numpy array into page-locked memory (torch's pinned allocator)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    try:
        t = t.pin_memory()
    except Exception:
        pass
    return t.numpy(), t


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every few ms from a thread (the timed
    region is ~0.1-0.3 s, too short for `nvidia-smi -lms`); falls back to nvidia-smi when pynvml is missing."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.nv, self.stop_flag = gpu_index, [], None, None, False
        self.sm, self.reasons, self.mx = [], set(), None

    def _poll(self):
        nv, h = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates in PCI order like nvidia-smi; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.idx
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                phys = int(vis.split(",")[self.idx])
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nv = (nv, h)
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv:
            self.stop_flag = True
            self.th.join(timeout=1.0)
            order = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": [n for n in order if n in self.reasons], "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


def measured_hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def committed_profile(name):
    """A committed JSON summary under profiles/ (ncu traffic per launch, the L2 probe), or {}."""
    p = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(p))
    except Exception:
        return {}


def masks_of(prob):
    ptr, ind, _ = prob["train"]
    bad_u = (np.diff(ptr) == 0).astype(np.uint8)
    cnt_i = np.bincount(ind, minlength=prob["n_items"])
    return bad_u, (cnt_i == 0).astype(np.uint8), cnt_i


def init_factors(n_users, n_items, r, seed=1):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-0.01, 0.01, size=(n_users, r)).astype(np.float32),
            rng.uniform(-0.01, 0.01, size=(n_items, r)).astype(np.float32))


class Ranks:
    """The process group seen from this rank (a no-op group at N = 1)."""

    def __init__(self, dist, rank, world, local_rank):
        self.dist, self.rank, self.world, self.local = dist, rank, world, local_rank

    def max(self, x):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, xs):
        if self.dist is None:
            return [float(x) for x in xs]
        import torch
        t = torch.tensor([float(x) for x in xs], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def gather(self, x):
        if self.dist is None:
            return [x]
        out = [None] * self.world
        self.dist.all_gather_object(out, x)
        return out

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


# ---------------------------------------------------------------------------------------------
def run_dsgd(prob, rk, r, plan, warmup, steps, options=None, seed=1, curve=False, U0V0=None):
    """DSGD over rk.world ranks (one rank in this process).  Returns the timing / RMSE record (max over ranks)."""
    import torch
    from matfac_b200 import dsgd
    from matfac_b200 import engine as E
    n_users, n_items = prob["n_users"], prob["n_items"]
    bad_u, bad_i, _ = masks_of(prob)
    U0, V0 = U0V0 if U0V0 is not None else init_factors(n_users, n_items, r)
    P = rk.world
    n_sub = (warmup + steps) * P
    opts = {"sgd_flat_inflight_frac": 8e-4}
    opts.update(options or {})
    d = dsgd.Dsgd(n_users, n_items, r, P, {rk.rank: rk.local}, prob["train"], prob["val"], U0, V0, bad_u, bad_i, n_sub,
                  plan=plan, seed=seed, exchange=rk.gather, options=opts)
    eng = d.engines[rk.rank]

    def sync_all():
        d.barrier()
        eng.sync()
        torch.cuda.synchronize()
        rk.barrier()

    rmse_curve = []

    def val_rmse():
        d.publish()
        s = rk.sum(d.eval_sums(E.VAL))
        d.barrier()
        return float(np.sqrt(s[0] / s[1])) if s[1] > 0 else float("nan")

    for w in range(warmup):
        d.run(w * P, (w + 1) * P, HP["lr"], HP["ureg"], HP["ireg"], seed)
        if curve:
            rmse_curve.append(val_rmse())
    sync_all()
    launches0 = E.launch_count()
    eng.event_record(0)
    if curve:  # per-epoch evaluation outside the timed spans
        ms = 0.0
        for k in range(steps):
            eng.event_record(0)
            d.run((warmup + k) * P, (warmup + k + 1) * P, HP["lr"], HP["ureg"], HP["ireg"], seed)
            eng.event_record(1)
            sync_all()
            ms += eng.event_elapsed_ms(0, 1)
            rmse_curve.append(val_rmse())
    else:
        d.run(warmup * P, (warmup + steps) * P, HP["lr"], HP["ureg"], HP["ireg"], seed)
        eng.event_record(1)
        sync_all()
        ms = eng.event_elapsed_ms(0, 1)  # kernels + exchange pushes + flag waits on this rank's stream
    launches = E.launch_count() - launches0
    if eng.comm_error():
        raise SystemExit(f"[rank {rk.rank}] a device-side wait timed out: the exchange schedule is broken")
    visited = d.block_nnz(warmup * P, (warmup + steps) * P)[rk.rank]
    ms_max = rk.max(ms)
    per_rank = rk.gather(int(visited))
    total = int(sum(per_rank))
    rmse = rmse_curve[-1] if curve else val_rmse()
    out = {"plan": plan, "ms_per_step": ms_max / steps, "value": total / (ms_max * 1e-3), "val_rmse": rmse,
           "ratings_visited": total, "ratings_visited_per_rank": per_rank, "ms_this_rank": ms, "launches": int(launches),
           "epochs": warmup + steps}
    if curve:
        out["val_rmse_curve"] = rmse_curve
    d.close()
    return out


def dsgd_e2e(prob, rk, r, steps):
    """End-to-end DSGD step at N > 1 through the C ABI with HOST buffers, one rank per process.  Every step, every rank:
    uploads ITS user stratum's rating rows (train + validation CSR) and the initial factors from pinned host memory,
    plans its blocks, runs one epoch of rk.world sub-epochs (item blocks stored peer to peer), makes V complete on every
    rank, evaluates its share of the validation RMSE (summed over the ranks by the process group) and downloads the
    factors.  Wall clock between cross-rank barriers, max over ranks; value = ratings visited / that time.

    Every collective below runs on every rank whatever happened locally (a rank that failed keeps answering them), and
    the device-side waits time out by themselves (csrc/comm.cu), so a failure ends as {"error": ...} in the line."""
    import torch
    from matfac_b200 import dsgd
    from matfac_b200 import engine as E
    n_users, n_items = prob["n_users"], prob["n_items"]
    bad_u, bad_i, _ = masks_of(prob)
    U0, V0 = init_factors(n_users, n_items, r)
    P = rk.world
    err, d, eng = None, None, None
    try:
        d = dsgd.Dsgd(n_users, n_items, r, P, {rk.rank: rk.local}, prob["train"], prob["val"], U0, V0, bad_u, bad_i, (steps + 1) * P,
                      plan="reference", seed=1, exchange=rk.gather, options={"sgd_flat_inflight_frac": 8e-4})
        eng = d.engines[rk.rank]
        mine = d.user_part == rk.rank
        ltr, lva = dsgd.local_rows(n_users, n_items, prob["train"], mine), dsgd.local_rows(n_users, n_items, prob["val"], mine)
        keep, h = [], {}
        for name, a in (("tp", ltr.rowptr), ("ti", ltr.rowind), ("tv", ltr.rowval), ("vp", lva.rowptr), ("vi", lva.rowind),
                        ("vv", lva.rowval), ("U", U0), ("V", V0), ("Uo", np.empty_like(U0)), ("Vo", np.empty_like(V0))):
            h[name], t = pinned(a)
            keep.append(t)
        trp, vap = Mat(n_users, n_items, (h["tp"], h["ti"], h["tv"])), Mat(n_users, n_items, (h["vp"], h["vi"], h["vv"]))
        up_local = np.where(mine, d.user_part, -1).astype(np.int32)
        h2d = sum(h[k].nbytes for k in ("tp", "ti", "tv", "vp", "vi", "vv", "U", "V"))
        d2h = h["Uo"].nbytes + h["Vo"].nbytes + 16
    except Exception as ex:  # noqa: BLE001
        err = repr(ex)[:300]
        h2d = d2h = 0
    times, visited, rmse = [], [], float("nan")
    for s in range(steps + 1):
        if err is None:
            try:
                d.barrier()
                eng.sync()
            except Exception as ex:  # noqa: BLE001
                err = repr(ex)[:300]
        rk.barrier()
        t1 = time.perf_counter()
        sums, n_vis = np.zeros(2), 0
        if err is None:
            try:
                eng.upload_csr(E.TRAIN, trp, with_csc=False)
                eng.upload_csr(E.VAL, vap, with_csc=False)
                eng.upload_factors(h["U"], h["V"])
                eng.sgd_plan(P, up_local, d.item_part)
                d.barrier()  # no peer stores item rows into this rank's V before its factor upload has landed
                d.run(s * P, (s + 1) * P, HP["lr"], HP["ureg"], HP["ireg"], 1)
                d.publish()
                sums = d.eval_sums(E.VAL)
                eng.L.mfb_download_factors(eng.h, E.CURRENT, h["Uo"].ctypes.data, r, h["Vo"].ctypes.data, r)
                eng.sync()
                n_vis = d.block_nnz(s * P, (s + 1) * P)[rk.rank]
                if eng.comm_error():
                    err = "a device-side wait timed out"
            except Exception as ex:  # noqa: BLE001
                err = repr(ex)[:300]
        tot = rk.sum([float(sums[0]), float(sums[1]), float(n_vis), 1.0 if err else 0.0])  # the step's result: validation SSE / count over all ranks
        dt = time.perf_counter() - t1
        times.append(rk.max(dt))
        visited.append(tot[2])
        if tot[3] > 0:
            err = err or "another rank failed"
        rmse = float(np.sqrt(tot[0] / tot[1])) if tot[1] > 0 else float("nan")
    try:
        if d is not None:
            d.close()
    except Exception:  # noqa: BLE001
        pass
    torch.cuda.empty_cache()
    if err is not None:
        return {"error": err}
    h2d_all, d2h_all = rk.sum([float(h2d)])[0], rk.sum([float(d2h)])[0]
    t = float(np.sum(times[1:]))
    return {"value": float(np.sum(visited[1:])) / t, "unit": "rating-updates/s", "ms_per_step": t / steps * 1e3, "steps_timed": steps,
            "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all), "val_rmse_last_step": rmse,
            "h2d_bytes_per_step_this_rank": int(h2d),
            "what": "per step and rank: upload this rank's user stratum (train + validation CSR) and the factors from pinned host memory, "
                    "plan, one DSGD epoch (reference partitions + update sequences, item blocks peer to peer), V made complete on every "
                    "rank, validation RMSE summed over the ranks, factor download; wall clock between cross-rank barriers, max over ranks; "
                    "bytes are summed over the ranks"}


def solver_timings(prob, rk, peak_gbs, ranks=(64, 128), ccd=True, objective=True, shape_name="netflix"):
    """The other trainers of the path on the same matrix (BASELINE.json: 'ALS epoch sec'): ALS at rank 64
    and 128, CCD++ (FreqAdap) at rank 64 and the objective pass.  CUDA events, max over ranks; at N > 1 rows are sharded in
    contiguous ranges of equal rating counts and the kernels store their output rows into all peers (csrc/comm.cu)."""
    from matfac_b200 import engine as E
    n_users, n_items = prob["n_users"], prob["n_items"]
    ptr, ind, val = prob["train"]
    nnz = int(ptr[-1])
    tr = Mat(n_users, n_items, prob["train"])
    bad_u, bad_i, cnt_i = masks_of(prob)
    out = {}
    for r in ranks:
        U0, V0 = init_factors(n_users, n_items, r)
        eng = E.Engine(n_users, n_items, r, device=rk.local)
        eng.upload_csr(E.TRAIN, tr, with_csc=False)
        eng.build_csc(E.TRAIN)  # gk_csr_CreateIndex on the device
        eng.set_masks(bad_u, bad_i)
        eng.set_aux(E.MF, np.diff(ptr).astype(np.int32), cnt_i.astype(np.int32))
        eng.upload_factors(U0, V0)
        if rk.world > 1:
            eng.comm_connect(rk.gather(eng.comm_init(rk.rank, rk.world)))
            colptr = np.zeros(n_items + 1, np.int64)
            np.cumsum(cnt_i, out=colptr[1:])
            ucut = np.searchsorted(ptr, np.linspace(0, nnz, rk.world + 1)).astype(int)
            icut = np.searchsorted(colptr, np.linspace(0, nnz, rk.world + 1)).astype(int)
            ucut[0] = icut[0] = 0
            ucut[-1], icut[-1] = n_users, n_items
            eng.set_row_range(E.USER, int(ucut[rk.rank]), int(ucut[rk.rank + 1]))
            eng.set_row_range(E.ITEM, int(icut[rk.rank]), int(icut[rk.rank + 1]))

        reps_log = {}

        def timed(fn, reps, tag=None):
            ms = []
            for _ in range(reps):
                eng.sync()
                rk.barrier()
                eng.event_record(0)
                fn()
                eng.event_record(1)
                ms.append(eng.event_elapsed_ms(0, 1))
            if tag:
                reps_log[tag] = [float(x) for x in ms]
            return rk.max(float(np.median(ms)))

        def als_epoch():
            eng.als_half_step(E.USER, 0.1)
            eng.als_half_step(E.ITEM, 0.1)

        als_epoch()  # warm-up epoch: plans are built here
        m = timed(als_epoch, 3, "als")
        out[f"als_rank{r}"] = {"epoch_ms": m, "epoch_ms_reps_this_rank": reps_log["als"], "epoch_sec": m * 1e-3, "gram_tflops_algorithmic": 4.0 * r * r * nnz / (m * 1e-3) / 1e12,
                               "gather_gbs": 2.0 * nnz * r * 4 / (m * 1e-3) / 1e9, "gather_bound_ms": 2.0 * nnz * r * 4 / (peak_gbs * 1e9) * 1e3 / rk.world,
                               "gram": "tcgen05 kind::tf32 x3 split, fp32 TMEM accumulator"}
        if r == 64 and ccd:
            eng.upload_factors(U0, V0)
            eng.ccdpp_begin()
            for k in range(4):
                eng.ccdpp_rank1(k, True, 5, 0.05, 0.05, 75)

            def ccd_steps():
                for k in range(8):
                    eng.ccdpp_rank1(k, False, 5, 0.05, 0.05, 75)

            mk = timed(ccd_steps, 1) / 8
            eng.ccdpp_end()
            gbs = 128.0 * nnz / (mk * 1e-3) / 1e9
            out["ccdpp_rank64"] = {"ms_per_rank_one_step": mk, "epoch_ms": mk * r, "algorithmic_gbs": gbs,
                                   "frac_of_hbm": gbs / (peak_gbs * rk.world),
                                   "what": "trainCCDPPFreqAdap k-loop body, 128 B per rating per rank-one step (SURVEY 8d); "
                                           "frac_of_hbm is against N x the measured HBM bandwidth"}
        if r == 64 and objective:
            def obj():
                eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
            obj(); obj()
            me = timed(obj, 3)
            out["objective_rank64"] = {"ms_per_pass": me, "algorithmic_gbs": (8.0 + 8.0 * r) * nnz / (me * 1e-3) / 1e9,
                                       "what": "objective + norms over the train matrix, 8 + 8r B per rating; u is reused from registers"}
        if rk.world > 1:
            eng.sync()
            rk.barrier()
            eng.comm_disconnect()
            rk.barrier()
        eng.close()
    return out


def yahoo_block(rk, peak_gbs, scale, epochs=4):
    """BASELINE.json configs[4]: Yahoo-R1-shaped matrix (1 M x 625 k, 250 M ratings), rank 64.  U + V = 416 MB do not fit
    the 126 MB L2: this is the shape on which the HBM roofline binds.  N = 1: shuffled SGD epoch, CCD++ step, objective;
    N > 1: DSGD (reference plan) epoch and row-sharded CCD++ step."""
    import torch
    from matfac_b200 import engine as E
    t0 = time.time()
    prob = make_problem(YAHOO_SHAPE, scale, f"cuda:{rk.local}")
    torch.cuda.empty_cache()
    n_users, n_items = prob["n_users"], prob["n_items"]
    nnz = int(prob["train"][0][-1])
    crcs = rk.gather(prob["crc"])
    out = {"workload": "modelMF rank 64, Yahoo-R1-shaped synthetic ratings (BASELINE.json configs[4])", "n_users": n_users,
           "n_items": n_items, "train_nnz": nnz, "matrix_crc": prob["crc"], "same_matrix_on_all_ranks": len(set(crcs)) == 1,
           "gen_s": time.time() - t0, "factor_bytes": 4 * RANK * (n_users + n_items)}
    if rk.world == 1:
        bad_u, bad_i, _ = masks_of(prob)
        U0, V0 = init_factors(n_users, n_items, RANK)
        eng = E.Engine(n_users, n_items, RANK, device=rk.local)
        eng.upload_csr(E.TRAIN, Mat(n_users, n_items, prob["train"]), with_csc=False)
        eng.upload_csr(E.VAL, Mat(n_users, n_items, prob["val"]), with_csc=False)
        eng.set_masks(bad_u, bad_i)
        eng.upload_factors(U0, V0)
        eng.sgd_plan(1)
        for ep in range(2):
            eng.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, ep)
        eng.sync()
        eng.event_record(0)
        for ep in range(epochs):
            eng.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, 2 + ep)
        eng.event_record(1)
        ms = eng.event_elapsed_ms(0, 1) / epochs
        alg = (16 * RANK + 12) * float(nnz) / (ms * 1e-3) / 1e9
        prof = committed_profile("r2_ncu_traffic.json").get("yahoo_sgd_epoch", {})
        out["sgd"] = {"ms_per_epoch": ms, "value": nnz / (ms * 1e-3), "unit": "rating-updates/s", "val_rmse": eng.rmse(E.VAL),
                      "epochs": 2 + epochs,
                      "roofline": {"bound": "hbm", "achieved": alg, "peak": peak_gbs, "unit": "GB/s", "frac": alg / peak_gbs,
                                   "traffic": prof.get("dram_bytes"), "traffic_source": prof.get("source"),
                                   "algorithmic_bytes_per_update": 16 * RANK + 12}}
        eng.close()
        del eng
    else:
        # engine-default in-flight budget: at 8e-4 (the Netflix-shape setting of run_dsgd) this matrix diverges at N = 8
        # (gpurun_out/s25, s26: NaN from epoch 1; 2e-4 and 4e-4 converge at the same epoch time)
        d = run_dsgd(prob, rk, RANK, "reference", 2, epochs, options={"sgd_flat_inflight_frac": 2e-4})
        out["dsgd"] = d
    torch.cuda.empty_cache()
    out["solvers"] = solver_timings(prob, rk, peak_gbs, ranks=(64,), objective=(rk.world == 1), shape_name="yahoo")
    bound_ms = 128.0 * nnz / (peak_gbs * 1e9) * 1e3 / rk.world
    if "ccdpp_rank64" in out["solvers"]:
        out["solvers"]["ccdpp_rank64"]["hbm_bound_ms_per_step"] = bound_ms
    return out


def oracle_check(rk, epochs=10, scale=0.05):
    """N > 1, outside every timed region: the DSGD path of this run (reference plan, N ranks over peer memory) next to the
    ORACLE's trainSGDPar with P = N threads on the 1/20-scale bench matrix — same seed, hence the same partitions and
    update sequences (tests/test_host.py pins the host plan bit for bit), same initial factors (the oracle's).  The oracle
    is the checker here, nothing of it is timed or shipped."""
    import torch
    prob = make_problem(SHAPE, scale, f"cuda:{rk.local}")
    n_users, n_items = prob["n_users"], prob["n_items"]
    U0 = torch.zeros(n_users, RANK, dtype=torch.float32, device="cuda")
    V0 = torch.zeros(n_items, RANK, dtype=torch.float32, device="cuda")
    ref_curve, ref_s = None, None
    if rk.rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        from matfac_b200 import synth
        tr = synth.Csr(n_users, n_items, *prob["train"])
        va = synth.Csr(n_users, n_items, *prob["val"])
        od = ol.OracleData(tr, va, va)
        om = ol.OracleModel(od, algo="mf", facdim=RANK, maxiter=epochs, seed=1, nthreads=rk.world, ureg=HP["ureg"],
                            ireg=HP["ireg"], learnrate=HP["lr"])
        u0, v0 = om.factors()
        U0.copy_(torch.from_numpy(u0)); V0.copy_(torch.from_numpy(v0))
    rk.dist.broadcast(U0, 0)
    rk.dist.broadcast(V0, 0)
    dev = run_dsgd(prob, rk, RANK, "reference", 0, epochs, seed=1, curve=True, U0V0=(U0.cpu().numpy(), V0.cpu().numpy()),
                   options={"sgd_flat_inflight_frac": 2e-4})
    if rk.rank == 0:
        t0 = time.time()
        om.train("sgdpar", keep_history=True)
        ref_curve = [h[3] for h in om.history()]
        ref_s = time.time() - t0
    rk.barrier()
    if rk.rank != 0:
        return None
    last = min(len(ref_curve), len(dev["val_rmse_curve"]))
    rel = [abs(a - b) / b for a, b in zip(dev["val_rmse_curve"][:last], ref_curve[:last])]
    return {"what": f"DSGD at {rk.world} ranks vs oracle trainSGDPar P={rk.world}: 1/20-scale matrix, reference partitions + "
                    "sgdUpdateBlockSeq sequences (seed 1), oracle initial factors; validation RMSE after every epoch",
            "matrix_crc": prob["crc"], "train_nnz": int(prob["train"][0][-1]), "device_val_rmse": dev["val_rmse_curve"],
            "oracle_val_rmse": ref_curve, "rel_diff": rel, "oracle_seconds": ref_s,
            "oracle_seed_spread_note": "profiles/r2_dsgd_oracle_curves.json: the oracle's own curves for seeds 1, 2 and P = 1..8"}


# ---------------------------------------------------------------------------------------------
def cpu_reference_arm(prob, sample_users=None, epochs=101):
    """Times the reference's OpenMP stratified SGD (ModelMF::trainSGDPar) on the first `sample_users` users of the
    workload with every host core.  oracle/_ref/mf_ref is the reference's own code; it prints its epoch timer
    (`subIterDuration`, modelMF.cpp:309,331) only every DISP_ITER = 50 epochs (const.h:5), so the run does 101 epochs and
    the value is the median of the printed epochs 0, 50 and 100 (BASELINE.md §4 asks for steady epochs; epoch 0 is the
    cold one).  Falls back on the oracle port (every epoch timed; median of epochs 1..) when mf_ref is absent.  The sample
    holds ~4 % of the ratings: the figure is the sample's throughput, i.e. an EXTRAPOLATION to the full matrix (whose U
    does not stay in the CPU's caches)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from matfac_b200 import synth
    cores = os.cpu_count() or 1
    n_users, n_items = prob["n_users"], prob["n_items"]
    ptr, ind, val = prob["train"]
    if sample_users is None:
        # ~4 M ratings: ~0.2 s of CPU work per epoch
        sample_users = int(np.searchsorted(ptr, 4_000_000))
        sample_users = max(1000, min(sample_users, n_users))
    nnz_s = int(ptr[sample_users])
    tr = synth.Csr(sample_users, n_items, ptr[: sample_users + 1].copy(), ind[:nnz_s], val[:nnz_s])
    vptr, vind, vval = prob["val"]
    vn = int(vptr[sample_users])
    va = synth.Csr(sample_users, n_items, vptr[: sample_users + 1].copy(), vind[:vn], vval[:vn])
    what = f"first {sample_users} users ({nnz_s} ratings = {100.0 * nnz_s / int(ptr[-1]):.1f} % of the workload, extrapolated)"
    od = ol.OracleData(tr, va, va)

    def visited_in(epoch_ids, n_epochs):
        # ratings visited per epoch of trainSGDPar: P update sequences drawn with replacement (util.cpp:1077)
        om = ol.OracleModel(od, algo="mf", facdim=RANK, maxiter=1, seed=1, nthreads=cores, ureg=HP["ureg"], ireg=HP["ireg"],
                            learnrate=HP["lr"])
        up, ip, sched = om.dsgd_plan(cores, n_epochs * cores)
        users = np.repeat(np.arange(sample_users), np.diff(tr.rowptr))
        blk = np.bincount(up[users].astype(np.int64) * cores + ip[tr.rowind], minlength=cores * cores)
        return [sum(int(blk[a * cores + b]) for s in range(e * cores, (e + 1) * cores) for a, b in sched[s]) for e in epoch_ids]

    kind, secs, ids = "port", None, None
    if ol.have_ref():
        try:
            d = tempfile.mkdtemp(prefix="mfref_")
            files = synth.write_split_files(d, tr, va, va)
            res = ol.run_ref(files, os.path.join(d, "dump"), algo="mf", method="sgdpar", threads=cores, timeout=900,
                             facdim=RANK, maxiter=epochs, seed=1, ureg=HP["ureg"], ireg=HP["ireg"], learnrate=HP["lr"])
            secs, ids = [], []
            for line in res["stdout"].splitlines():
                if "subIterDuration:" in line and " Iter: " in line:
                    ids.append(int(line.split(" Iter: ")[1].split()[0]))
                    secs.append(float(line.split("subIterDuration:")[1].split()[0]))
            if not secs:
                raise RuntimeError("no subIterDuration line in the reference's output")
            kind = "reference"
            sample = (f"{what}, mf_ref --mf_method sgdpar, OMP_NUM_THREADS={cores}, the reference's own epoch timer at epochs "
                      f"{ids} (it prints every 50th), median")
        except Exception as ex:  # fall back to the port
            log("mf_ref failed, using the oracle port:", repr(ex)[:200])
            secs = None
    if not secs:
        n = 4
        om = ol.OracleModel(od, algo="mf", facdim=RANK, maxiter=n, seed=1, nthreads=cores, ureg=HP["ureg"], ireg=HP["ireg"],
                            learnrate=HP["lr"])
        om.train("sgdpar")
        secs = [float(x) for x in om.epoch_seconds()]
        ids = list(range(len(secs)))
        if len(secs) > 1:
            secs, ids = secs[1:], ids[1:]
        sample = f"{what}, oracle port of trainSGDPar, P={cores}, median of epochs {ids}"
    visited = visited_in(ids, max(ids) + 1)
    rates = [v / s for v, s in zip(visited, secs)]
    value = float(np.median(rates))
    return dict(value=value, unit="rating-updates/s", cores=cores, kind=kind, sample=sample, extrapolated=True,
                epoch_seconds=secs), float(np.median(secs)) * 1e3


# ---------------------------------------------------------------------------------------------
def _protect_stdout():
    """Libraries (NCCL's version banner, torchrun's OMP notice) write to fd 1; the contract is ONE JSON line on
    stdout.  Route fd 1 to stderr for the duration of the run and keep the real stdout for the final line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    real_stdout = _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="device", choices=["device", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-solvers", action="store_true", help="skip the ALS / CCD++ / objective timings")
    ap.add_argument("--no-yahoo", action="store_true", help="skip the Yahoo-R1-shaped block (BASELINE.json configs[4])")
    ap.add_argument("--no-oracle-check", action="store_true", help="N > 1: skip the DSGD-vs-oracle RMSE check")
    ap.add_argument("--dsgd-plan", default="reference", choices=["reference", "balanced"],
                    help="N > 1: the plan `value` is measured on (the other one is reported next to it)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    have_cuda = torch.cuda.is_available()
    config = {"workload": "modelMF SGD rank 64, Netflix-shaped synthetic ratings (BASELINE.json configs[1])",
              "rank": RANK, "learnrate": HP["lr"], "ureg": HP["ureg"], "ireg": HP["ireg"], "data_seed": DATA_SEED,
              "l2": "inputs larger than L2 (ratings 0.8 GB + U 123 MB), no flush"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        prob = make_problem(SHAPE, args.scale, f"cuda:{local_rank}" if have_cuda else "cpu")
        cb, ms = cpu_reference_arm(prob)
        config.update(n_users=prob["n_users"], n_items=prob["n_items"], train_nnz=int(prob["train"][0][-1]), matrix_crc=prob["crc"],
                      same_matrix_on_all_ranks=True)
        line = {"impl": "reference", "metric": "sgd_rating_updates_per_sec", "value": cb["value"], "unit": "rating-updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "rating-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=real_stdout, flush=True)
        return 0

    if not have_cuda:
        raise SystemExit("bench.py: no CUDA device — the device arm has no CPU fallback")
    from matfac_b200 import engine as E
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rk = Ranks(dist, rank, world, local_rank)

    t0 = time.time()
    prob = make_problem(SHAPE, args.scale, f"cuda:{local_rank}")  # every rank generates the same matrix (checked below)
    torch.cuda.empty_cache()
    n_users, n_items = prob["n_users"], prob["n_items"]
    ptr, ind, val = prob["train"]
    train_nnz = int(ptr[-1])
    crcs = rk.gather(prob["crc"])
    config.update(n_users=n_users, n_items=n_items, train_nnz=train_nnz, matrix_crc=prob["crc"],
                  same_matrix_on_all_ranks=len(set(crcs)) == 1)
    if len(set(crcs)) != 1:
        raise SystemExit(f"bench.py: the ranks generated different matrices: {crcs}")
    log(f"[rank {rank}] data: {train_nnz} train ratings, crc {prob['crc']}, max user degree {int(np.diff(ptr).max())}, gen {time.time()-t0:.1f}s")
    peak, peak_src = measured_hbm_gbs()

    U0, V0 = init_factors(n_users, n_items, RANK)
    bad_u, bad_i, _ = masks_of(prob)
    extra = {}
    if world == 1:
        tr = Mat(n_users, n_items, prob["train"])
        va = Mat(n_users, n_items, prob["val"])
        eng = E.Engine(n_users, n_items, RANK, device=local_rank)
        eng.upload_csr(E.TRAIN, tr, with_csc=False)
        eng.upload_csr(E.VAL, va, with_csc=False)
        eng.set_masks(bad_u, bad_i)
        eng.upload_factors(U0, V0)
        eng.set_option("sgd_shuffle_seed", 1)
        eng.sgd_plan(1)
        for w in range(args.warmup):
            eng.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, w)
        eng.sync()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = E.launch_count()
        eng.event_record(0)
        for k in range(args.steps):
            eng.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, args.warmup + k)
        eng.event_record(1)
        eng.sync()
        torch.cuda.synchronize()
        ms_dev = eng.event_elapsed_ms(0, 1)
        launches = E.launch_count() - launches0
        clocks = sampler.stop()
        ms_per_step = ms_dev / args.steps
        value = train_nnz / (ms_per_step * 1e-3)
        val_rmse = eng.rmse(E.VAL)
        my_nnz_per_epoch = train_nnz
        kernel_name = "sgd_flat_kernel<16,1,MF> + sgd_hot_kernel<4,4,MF> (hot item rows, concurrent stream)"
    else:
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        main_plan = args.dsgd_plan
        d = run_dsgd(prob, rk, RANK, main_plan, args.warmup, args.steps)
        clocks = sampler.stop() if rank == 0 else None
        ms_per_step, value, val_rmse, launches = d["ms_per_step"], d["value"], d["val_rmse"], d["launches"]
        ms_dev = d["ms_this_rank"]
        my_nnz_per_epoch = d["ratings_visited_per_rank"][rank] / args.steps
        extra["dsgd"] = d
        other = "balanced" if main_plan == "reference" else "reference"
        try:
            extra["dsgd_" + other] = run_dsgd(prob, rk, RANK, other, args.warmup, args.steps)
        except Exception as ex:
            extra["dsgd_" + other] = {"error": repr(ex)[:300]}
        kernel_name = "sgd_flat_kernel<16,1,MF> + sgd_hot_kernel<4,4,MF> per stratum block + comm_push_rows_kernel"
    log(f"[rank {rank}] {ms_per_step:.3f} ms/epoch, {value/1e9:.3f} G updates/s, val RMSE after {args.warmup+args.steps} epochs {val_rmse:.4f}")

    # roofline of the SGD update kernel.  N = 1: the two kernels of an epoch, timed by the events above.  N > 1: rank 0's
    # share of the algorithmic bytes over the same span of device time, which also holds the exchange pushes and flag
    # waits of the sub-epochs.
    alg_bytes = (16 * RANK + 12) * float(my_nnz_per_epoch)
    achieved = alg_bytes * args.steps / (ms_dev * 1e-3) / 1e9
    traffic = committed_profile("r2_ncu_traffic.json").get("netflix_sgd_epoch", {}) if (world == 1 and args.scale == 1.0) else {}
    l2p = committed_profile("r2_l2_probe.json")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes"), "traffic_source": traffic.get("source"), "peak_source": peak_src,
                "kernel": kernel_name, "algorithmic_bytes_per_update": 16 * RANK + 12,
                "launches_per_epoch": launches / args.steps,
                "note": "achieved = algorithmic bytes (u,v read + reduced, 12 B rating record, SURVEY 8d) of one epoch on this rank / "
                        "its device time; traffic = dram__bytes_read+write of the epoch's kernels from the committed ncu capture "
                        "(profiles/r2_ncu_traffic.json).  U (123 MB) + V (4.5 MB) stay in the 126 MB L2 on this shape, so DRAM traffic "
                        "is a fraction of the algorithmic bytes and frac can exceed 1; the bound that applies is `l2`.  The Yahoo-shaped "
                        "block (`yahoo.sgd.roofline`) is the same kernel with factors that do not fit L2."}
    if l2p.get("gather_red_gbs"):
        # the kernel's own access pattern at the L2: 256 B row gathers (ld.global.cg.v4) + red.global.add.v4 on the same rows
        roofline["l2"] = {"bound": "l2", "achieved": achieved, "peak": l2p["gather_red_gbs"], "unit": "GB/s",
                          "frac": achieved / l2p["gather_red_gbs"], "peak_source": "tools/l2_probe.cu on this pod (profiles/r2_l2_probe.json): "
                          "random 256 B row gather + vector reduction on an L2-resident matrix, algorithmic bytes counted the same way",
                          "l2_read_gbs": l2p.get("l2_read_gbs"), "red_v4_gbs": l2p.get("red_v4_gbs")}

    if world > 1:
        roofline["per_rank_nnz_per_epoch"] = [x / args.steps for x in extra["dsgd"]["ratings_visited_per_rank"]]

    line = {"metric": "sgd_rating_updates_per_sec", "value": value, "unit": "rating-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "val_rmse": val_rmse}
    line.update(extra)

    if world == 1:
        # how the epoch was split between the two kernels (diagnostics ABI; outside the timed region)
        try:
            _, cold_n, lists = eng.debug_sgd_records(0, 0, with_records=False)
            st = eng.debug_sgd_hot_batch()
            line["sgd_hot_rows"] = {"lists": int(len(lists)), "share_of_ratings": 1.0 - cold_n / max(train_nnz, 1),
                                    "longest_list": int(lists[:, 2].max()) if len(lists) else 0, "ratings_per_round": int(st[2]),
                                    "mean_user_norm_sq": float(st[0] / st[1]) if st[1] > 0 else 0.0}
        except Exception as ex:
            line["sgd_hot_rows"] = {"error": repr(ex)[:200]}

        # the stratified trainer's kernel in the reference's visiting order (user-major runs), reference plan, P = 8
        try:
            from matfac_b200 import dsgd
            P = 8
            up, ip, sched = dsgd.reference_plan(n_users, n_items, bad_u, bad_i, 1, P, 4 * P)
            eng.sgd_plan(P, up, ip)

            def sub(t):
                blocks = np.stack([np.arange(P), sched[t]], 1).astype(np.int32)
                eng.sgd_subepoch(blocks, E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, t)
            for t in range(P):
                sub(t)
            eng.event_record(2)
            for t in range(P, 4 * P):
                sub(t)
            eng.event_record(3)
            ms_s = eng.event_elapsed_ms(2, 3) / 3
            visited = sum(eng.sgd_block_nnz(np.stack([np.arange(P), sched[t]], 1).astype(np.int32)) for t in range(P, 4 * P)) / 3
            line["stratified"] = {"P": P, "ms_per_epoch": ms_s, "value": visited / (ms_s * 1e-3), "unit": "rating-updates/s",
                                  "what": "trainSGDPar on one GPU: reference partitions + update sequences, user-major runs, "
                                          "hot-item concurrency capped at 8 (the reference's visiting order)"}
            eng.sgd_plan(1)
        except Exception as ex:
            line["stratified"] = {"error": repr(ex)[:200]}

        # end-to-end through the C ABI with host buffers
        if args.e2e_steps > 0:
            h, keep = {}, []
            for name, a in (("ptr", ptr), ("ind", ind), ("val", val), ("U", U0.copy()), ("V", V0.copy())):
                h[name], t = pinned(a)
                keep.append(t)
            Uo, tU = pinned(np.empty_like(U0))
            Vo, tV = pinned(np.empty_like(V0))
            trp = Mat(n_users, n_items, (h["ptr"], h["ind"], h["val"]))
            h2d = h["ptr"].nbytes + h["ind"].nbytes + h["val"].nbytes + h["U"].nbytes + h["V"].nbytes
            d2h = Uo.nbytes + Vo.nbytes + 64
            # option copy_overlap: the copies return without synchronising (pinned buffers, valid until the sync that
            # closes the step); the factor upload runs behind the rating upload and next to the plan kernels, the
            # factor download next to the two evaluation passes.  Every byte still crosses the link every step.
            eng.set_option("copy_overlap", 1)
            times, parts = [], []
            for s in range(args.e2e_steps + 1):
                eng.sync()
                t1 = time.perf_counter()
                eng.upload_csr(E.TRAIN, trp, with_csc=False)
                eng.upload_factors(h["U"], h["V"])
                eng.sgd_plan(1)
                t3 = time.perf_counter()
                eng.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, s)
                eng.L.mfb_download_factors(eng.h, E.CURRENT, Uo.ctypes.data, RANK, Vo.ctypes.data, RANK)
                obj = eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
                ev = eng.eval(E.VAL)
                t4 = time.perf_counter()
                eng.sync()
                t5 = time.perf_counter()
                times.append(t5 - t1)
                parts.append([t3 - t1, t4 - t3, t5 - t4])
            # the same step when the ratings stay resident between steps, as they do in a training run (the reference reads
            # its CSR once, datastruct.cpp:16): only the factors cross the link.  Reported NEXT to e2e, not instead of it.
            times_res = []
            for s in range(args.e2e_steps + 1):
                eng.sync()
                t1 = time.perf_counter()
                eng.upload_factors(h["U"], h["V"])
                eng.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, s)
                eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
                eng.eval(E.VAL)
                eng.L.mfb_download_factors(eng.h, E.CURRENT, Uo.ctypes.data, RANK, Vo.ctypes.data, RANK)
                eng.sync()
                times_res.append(time.perf_counter() - t1)
            # Two steps in flight: a second engine on the same GPU, one host thread per engine, every thread runs the SAME
            # complete steps as above (upload CSR + factors, plan, epoch, objective + validation, download, sync) — one
            # engine's PCIe transfers run next to the other's kernels, as a caller that pipelines its batches would have it.
            # Every byte of every step still crosses the link inside the timed region.
            piped = None
            try:
                import threading
                eng2 = E.Engine(n_users, n_items, RANK, device=local_rank)
                eng2.upload_csr(E.VAL, va, with_csc=False)
                eng2.set_masks(bad_u, bad_i)
                eng2.set_option("sgd_shuffle_seed", 1)
                eng2.set_option("copy_overlap", 1)
                Uo2, tU2 = pinned(np.empty_like(U0))
                Vo2, tV2 = pinned(np.empty_like(V0))
                K = max(2, args.e2e_steps)
                gate = threading.Barrier(2)
                link_turn, sm_turn = threading.Lock(), threading.Lock()
                span, errs, evs = {}, [], {}

                def pipeline(idx, en, uo, vo):
                    try:
                        for s in range(K + 1):
                            if s == 1:
                                en.sync()
                                gate.wait()
                                span[idx, 0] = time.perf_counter()
                            with link_turn:  # the two pipelines take turns on the link and on the SMs: one uploads while the other computes
                                en.upload_csr(E.TRAIN, trp, with_csc=False)
                                en.upload_factors(h["U"], h["V"])
                            with sm_turn:
                                en.sgd_plan(1)
                                en.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, s)
                                en.L.mfb_download_factors(en.h, E.CURRENT, uo.ctypes.data, RANK, vo.ctypes.data, RANK)
                                en.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
                                evs[idx] = en.eval(E.VAL)
                                en.sync()
                        span[idx, 1] = time.perf_counter()
                    except Exception as ex:  # noqa: BLE001
                        errs.append(repr(ex)[:200])
                        try:
                            gate.abort()
                        except Exception:  # noqa: BLE001
                            pass

                ths = [threading.Thread(target=pipeline, args=(0, eng, Uo, Vo)), threading.Thread(target=pipeline, args=(1, eng2, Uo2, Vo2))]
                for t in ths:
                    t.start()
                for t in ths:
                    t.join()
                if errs:
                    raise RuntimeError("; ".join(errs))
                wall = max(span[0, 1], span[1, 1]) - min(span[0, 0], span[1, 0])
                t_pipe = wall / (2 * K)
                piped = {"value": train_nnz / t_pipe, "unit": "rating-updates/s", "ms_per_step": t_pipe * 1e3, "steps_timed": 2 * K,
                         "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                         "val_rmse_last_step": [float(np.sqrt(evs[i][0] / max(evs[i][1], 1))) for i in (0, 1)],
                         "what": "two engines on the GPU, one host thread each, both run complete steps (upload CSR + factors from pinned "
                                 "host memory, plan, 1 epoch, objective + val RMSE, factor download, sync) and take turns on the link and on the SMs; "
                                 "wall time of all steps / steps: one engine's uploads overlap the other's kernels"}
                eng2.close()
                del eng2, tU2, tV2
            except Exception as ex:  # noqa: BLE001
                piped = {"error": repr(ex)[:300]}
            eng.set_option("copy_overlap", 0)
            t_res = float(np.median(times_res[1:]))
            line["e2e_resident_ratings"] = {"value": train_nnz / t_res, "unit": "rating-updates/s", "ms_per_step": t_res * 1e3,
                                            "h2d_bytes_per_step": int(h["U"].nbytes + h["V"].nbytes), "d2h_bytes_per_step": int(d2h),
                                            "what": "per step: upload factors, 1 epoch, objective + val RMSE, download factors; the rating "
                                                    "CSR and the plan stay on the device (uploaded / built once)"}
            t_e2e = float(np.median(times[1:]))
            pm = np.median(np.array(parts[1:]), axis=0) * 1e3
            one = {"value": train_nnz / t_e2e, "unit": "rating-updates/s", "h2d_bytes_per_step": int(h2d),
                           "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e * 1e3,
                           "ms_breakdown": {"upload_and_plan": float(pm[0]), "epoch_eval_and_download": float(pm[1]), "download_tail": float(pm[2])},
                           "val_rmse_last_step": float(np.sqrt(ev[0] / max(ev[1], 1))),
                           "what": "per step: upload CSR + factors from pinned host memory (the factor upload runs next to the plan "
                                   "kernels), plan, 1 epoch, factor download issued next to objective + val RMSE (the copy engine is "
                                   "starved by the evaluation's HBM traffic: no gain there, tools/e2e_probe.py), sync"}
            # headline: two steps in flight (throughput of complete steps, every byte of every step on the link); the single
            # step's latency stays next to it
            if piped and "value" in piped:
                line["e2e"] = piped
                line["e2e_one_in_flight"] = one
            else:
                line["e2e"] = one
                line["e2e_two_in_flight"] = piped
            del keep, tU, tV
        eng.close()
        del eng
        torch.cuda.empty_cache()

        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_reference_arm(prob)
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "unit": "rating-updates/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: " + repr(ex)[:200]}

    if world > 1 and args.e2e_steps > 0:
        try:
            e2e = dsgd_e2e(prob, rk, RANK, max(2, args.e2e_steps))
        except Exception as ex:
            e2e = {"error": repr(ex)[:300]}
        if rank == 0:
            line["e2e"] = e2e
    if not args.no_solvers:
        try:
            s = solver_timings(prob, rk, peak)
            if rank == 0:
                line["solvers"] = s
        except Exception as ex:
            line["solvers"] = {"error": repr(ex)[:300]}
    if world > 1 and not args.no_oracle_check:
        try:
            oc = oracle_check(rk)
            if rank == 0:
                line["oracle_check"] = oc
        except Exception as ex:
            line["oracle_check"] = {"error": repr(ex)[:300]}
    del prob
    torch.cuda.empty_cache()
    if not args.no_yahoo and args.scale == 1.0:
        try:
            y = yahoo_block(rk, peak, 1.0)
            if rank == 0:
                line["yahoo"] = y
        except Exception as ex:
            line["yahoo"] = {"error": repr(ex)[:300]}
    elif not args.no_yahoo:
        try:
            y = yahoo_block(rk, peak, args.scale)
            if rank == 0:
                line["yahoo"] = y
        except Exception as ex:
            line["yahoo"] = {"error": repr(ex)[:300]}
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
