#!/usr/bin/env python
"""bench.py — SGD rating-updates/s on a Netflix-shaped synthetic problem at rank 64.

Contract (one JSON line on stdout, printed by rank 0):
  python bench.py --gpus N --steps K --warmup W          device arm (this repo's CUDA engine)
  python bench.py --impl reference --gpus N ...          reference arm: the reference's own CPU
                                                         code (oracle/_ref/mf_ref) on host cores
A "step" is one SGD epoch over the whole training matrix: at N = 1 the serial-SGD trainer
(ModelMF::train, the CLI default --mf_method sgd) on the shuffled kernel; at N > 1 the stratified
trainer (DSGD: user strata pinned to ranks, item blocks exchanged after every sub-epoch).

Workload = BASELINE.json configs[1]: modelMF SGD, rank 64, 480,189 x 17,770, ~100.5 M ratings
(synthetic, Zipf-skewed positions, ratings from a rank-8 model + noise; random-init factors
U(-0.01, 0.01) as model.cpp:2331-2350).  Inputs (ratings 0.8 GB + U 123 MB) exceed the 126 MB L2,
so no explicit L2 flush is done between timed epochs.

  value     epochs * valid-ratings / device time (CUDA events on the engine's stream, max over
            ranks), inputs resident in HBM
  e2e       the same metric through the C ABI with HOST buffers: every step uploads the rating
            CSR and the factor matrices from pinned host memory, builds the stratum plan, runs
            one epoch plus the per-epoch evaluation and downloads the factors
  roofline  algorithmic bytes (16 r + 12 per update) / kernel time vs the measured
            HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the reference's OpenMP stratified SGD (trainSGDPar) on a row sample, host cores
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
SHAPE = (480_189, 17_770, 100_480_507)
# learnrate 0.002: at 0.005 the reference itself diverges on this matrix in epoch 0 and falls back on
# its NaN guard (restore + halve, model.cpp:1487-1498) — measured with the oracle, see DESIGN.md
HP = dict(lr=0.002, ureg=0.05, ireg=0.05)
FALLBACK_HBM_GBS = 6650.0
ITEM_SHARE_CAP = 232_944 / 100_480_507
# dram__bytes_read.sum + dram__bytes_write.sum of the SGD kernel launches of one epoch (profiles/r1_sgd_flat.md)
DRAM_TRAFFIC_BYTES_PER_EPOCH = 10.44e9


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
def gen_problem(n_users, n_items, nnz, seed, device):
    """Netflix-shaped training CSR (+ a 1 % validation CSR) generated with torch on `device`.
    Returns dict of numpy arrays (host)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dev = torch.device(device)

    def zipf(n, s):
        w = 1.0 / torch.arange(1, n + 1, dtype=torch.float64, device=dev).pow(s)
        w = w[torch.randperm(n, generator=g, device=dev)]
        return w / w.sum()

    pu, pi = zipf(n_users, 0.9), zipf(n_items, 1.05)
    # head of the item distribution as in the real Netflix Prize data: the most-rated title holds
    # 232,944 of 100,480,507 ratings (0.232 %); an uncapped Zipf(1.05) head would hold twice that
    for _ in range(8):
        pi = pi.clamp(max=ITEM_SHARE_CAP)
        pi = pi / pi.sum()
    ci = torch.cumsum(pi, 0)
    total = int(nnz * 1.01)
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    # every user and every item at least once (io.cpp:742-752)
    base_u = torch.arange(n_users, device=dev, dtype=torch.int64)
    base_i = torch.searchsorted(ci, torch.rand(n_users, generator=g, device=dev, dtype=torch.float64)).clamp_(max=n_items - 1)
    base2_i = torch.arange(n_items, device=dev, dtype=torch.int64)
    base2_u = torch.multinomial(pu.float(), n_items, replacement=True, generator=g)
    keys = torch.unique(torch.cat([base_u * n_items + base_i, base2_u * n_items + base2_i]))
    cap = 0.85 * n_items
    for rnd in range(12):
        need = total - keys.numel()
        if need <= 0:
            break
        draw = int(need * (1.6 if rnd == 0 else 1.3)) + 1024
        # user degrees ~ Zipf, capped so that no user exceeds ~85 % of the catalogue
        deg = torch.clamp(pu * draw, max=cap).round().to(torch.int64)
        u = torch.repeat_interleave(torch.arange(n_users, device=dev, dtype=torch.int64), deg)
        i = torch.searchsorted(ci, torch.rand(u.numel(), generator=g, device=dev, dtype=torch.float64)).clamp_(max=n_items - 1)
        keys = torch.unique(torch.cat([keys, u * n_items + i]))
        del u, i, deg
    if keys.numel() > total:
        drop = torch.randperm(keys.numel(), generator=g, device=dev)[: keys.numel() - total]
        mask = torch.ones(keys.numel(), dtype=torch.bool, device=dev)
        mask[drop] = False
        # never drop a user's or an item's covering pair: re-add them
        keys = torch.unique(torch.cat([keys[mask], base_u * n_items + base_i, base2_u * n_items + base2_i]))
    users = (keys // n_items).to(torch.int32)
    items = (keys % n_items).to(torch.int32)
    del keys
    tr_rank = 8
    us = torch.randn(n_users, tr_rank, generator=g, device=dev)
    vs = torch.randn(n_items, tr_rank, generator=g, device=dev)
    vals = torch.empty(users.numel(), dtype=torch.float32, device=dev)
    step = 1 << 24
    for s in range(0, users.numel(), step):
        e = min(s + step, users.numel())
        d = (us[users[s:e].long()] * vs[items[s:e].long()]).sum(1) / tr_rank ** 0.5
        vals[s:e] = 3.6 + 1.1 * d + 0.3 * torch.randn(e - s, generator=g, device=dev)
    vals = (torch.round(vals * 2) / 2).clamp_(1.0, 5.0)
    # split: 1 % validation, never a user's/item's first rating
    colour = torch.rand(users.numel(), generator=g, device=dev)
    first_u = torch.ones(users.numel(), dtype=torch.bool, device=dev)
    first_u[1:] = users[1:] != users[:-1]
    order_i = torch.argsort(items.long() * n_users + users.long())
    si = items[order_i]
    fi = torch.ones(si.numel(), dtype=torch.bool, device=dev)
    fi[1:] = si[1:] != si[:-1]
    first_i = torch.zeros(users.numel(), dtype=torch.bool, device=dev)
    first_i[order_i[fi]] = True
    is_val = (colour < 0.01) & ~first_u & ~first_i
    del order_i, si, fi, colour

    def csr(mask):
        u, i, v = users[mask], items[mask], vals[mask]
        ptr = torch.zeros(n_users + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(torch.bincount(u.long(), minlength=n_users), 0)
        return ptr.cpu().numpy(), i.cpu().numpy(), v.cpu().numpy()

    tr, va = csr(~is_val), csr(is_val)
    return dict(n_users=n_users, n_items=n_items, train=tr, val=va)


def shared_problem(n_users, n_items, nnz, seed, rank, dist, device):
    """N > 1: rank 0 generates the matrix and broadcasts it — the torch generator is not bit-reproducible across
    processes (atomics in unique / scatter), and every rank must partition the SAME matrix."""
    import torch
    prob = gen_problem(n_users, n_items, nnz, seed, device) if rank == 0 else None
    head = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == 0:
        head[0], head[1] = int(prob["train"][0][-1]), int(prob["val"][0][-1])
    dist.broadcast(head, 0)
    out = {"n_users": n_users, "n_items": n_items}
    for name, n in (("train", int(head[0])), ("val", int(head[1]))):
        arrs = []
        for k, (dt, ln) in enumerate(((torch.int64, n_users + 1), (torch.int32, n), (torch.float32, n))):
            t = torch.from_numpy(prob[name][k]).to(device) if rank == 0 else torch.empty(ln, dtype=dt, device=device)
            dist.broadcast(t, 0)
            arrs.append(t.cpu().numpy())
            del t
        out[name] = tuple(arrs)
    torch.cuda.empty_cache()
    return out


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



def csc_on_device(n_users, n_items, ptr, ind, val, device):
    """gk_csr_CreateIndex(mat, GK_CSR_COL) as a stable sort by column (torch, set-up only)."""
    import torch
    dev = torch.device(device)
    ind_d = torch.from_numpy(ind).to(dev)
    deg = torch.from_numpy(np.diff(ptr)).to(dev)
    rows = torch.repeat_interleave(torch.arange(n_users, device=dev, dtype=torch.int32), deg)
    order = torch.sort(ind_d.long() * n_users + rows.long()).indices  # (col, row) ascending == stable by col
    colind = rows[order].cpu().numpy()
    colval = torch.from_numpy(val).to(dev)[order].cpu().numpy()
    cnt = torch.bincount(ind_d.long(), minlength=n_items)
    colptr = np.zeros(n_items + 1, np.int64)
    colptr[1:] = torch.cumsum(cnt, 0).cpu().numpy()
    del ind_d, deg, rows, order, cnt
    torch.cuda.empty_cache()
    return colptr, colind, colval


def solver_timings(prob, device_index, peak_gbs):
    """The other trainers of the path on the same matrix (BASELINE.json: 'ALS epoch sec'):
    ALS at rank 64 and 128, CCD++ (FreqAdap) at rank 64 and the objective pass, CUDA events, one GPU."""
    from matfac_b200 import engine as E
    n_users, n_items = prob["n_users"], prob["n_items"]
    ptr, ind, val = prob["train"]
    nnz = int(ptr[-1])
    tr = Mat(n_users, n_items, prob["train"])
    bad_u = (np.diff(ptr) == 0).astype(np.uint8)
    cnt_i = np.bincount(ind, minlength=n_items)
    out = {}
    for r in (64, 128):
        rng = np.random.default_rng(1)
        eng = E.Engine(n_users, n_items, r, device=device_index)
        eng.upload_csr(E.TRAIN, tr, with_csc=False)
        eng.build_csc(E.TRAIN)  # gk_csr_CreateIndex on the device
        eng.set_masks(bad_u, (cnt_i == 0).astype(np.uint8))
        eng.set_aux(E.MF, np.diff(ptr).astype(np.int32), cnt_i.astype(np.int32))
        eng.upload_factors(rng.uniform(-0.01, 0.01, (n_users, r)).astype(np.float32),
                           rng.uniform(-0.01, 0.01, (n_items, r)).astype(np.float32))
        eng.als_half_step(E.USER, 0.1)
        eng.als_half_step(E.ITEM, 0.1)  # warm-up epoch: plans are built here
        ms = []
        for _ in range(2):
            eng.event_record(0)
            eng.als_half_step(E.USER, 0.1)
            eng.als_half_step(E.ITEM, 0.1)
            eng.event_record(1)
            ms.append(eng.event_elapsed_ms(0, 1))
        m = float(np.median(ms))
        out[f"als_rank{r}"] = {"epoch_ms": m, "epoch_sec": m * 1e-3, "gram_tflops_algorithmic": 4.0 * r * r * nnz / (m * 1e-3) / 1e12,
                               "gather_gbs": 2.0 * nnz * r * 4 / (m * 1e-3) / 1e9,
                               "gram": "tcgen05 kind::tf32 x3 split, fp32 TMEM accumulator" if r > 64 else "fp32 FMA register tiles"}
        if r == 64:
            eng.upload_factors(rng.uniform(-0.01, 0.01, (n_users, r)).astype(np.float32),
                               rng.uniform(-0.01, 0.01, (n_items, r)).astype(np.float32))
            eng.ccdpp_begin()
            for k in range(4):
                eng.ccdpp_rank1(k, True, 5, 0.05, 0.05, 75)
            eng.event_record(0)
            for k in range(8):
                eng.ccdpp_rank1(k, False, 5, 0.05, 0.05, 75)
            eng.event_record(1)
            mk = eng.event_elapsed_ms(0, 1) / 8
            eng.ccdpp_end()
            gbs = 128.0 * nnz / (mk * 1e-3) / 1e9
            out["ccdpp_rank64"] = {"ms_per_rank_one_step": mk, "epoch_ms": mk * r, "algorithmic_gbs": gbs, "frac_of_hbm": gbs / peak_gbs,
                                   "what": "trainCCDPPFreqAdap k-loop body, 128 B per rating per rank-one step (SURVEY 8d)"}
            for _ in range(2):
                eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
            eng.event_record(0)
            for _ in range(3):
                eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
            eng.event_record(1)
            me = eng.event_elapsed_ms(0, 1) / 3
            out["objective_rank64"] = {"ms_per_pass": me, "algorithmic_gbs": (8.0 + 8.0 * r) * nnz / (me * 1e-3) / 1e9,
                                       "what": "objective + norms over the train matrix, 8 + 8r B per rating; u is reused from registers"}
        eng.close()
    return out


# ---------------------------------------------------------------------------------------------
def cpu_reference_arm(prob, steps, warmup, sample_users=None):
    """Times the reference's OpenMP stratified SGD (ModelMF::trainSGDPar) on the first
    `sample_users` users of the workload with every host core.  Uses oracle/_ref/mf_ref (the
    reference's own code) when present, else the oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from matfac_b200 import synth
    cores = os.cpu_count() or 1
    n_users, n_items = prob["n_users"], prob["n_items"]
    ptr, ind, val = prob["train"]
    if sample_users is None:
        # ~4 M ratings: a few seconds of CPU work per epoch
        sample_users = int(np.searchsorted(ptr, 4_000_000))
        sample_users = max(1000, min(sample_users, n_users))
    nnz_s = int(ptr[sample_users])
    tr = synth.Csr(sample_users, n_items, ptr[: sample_users + 1].copy(), ind[:nnz_s], val[:nnz_s])
    vptr, vind, vval = prob["val"]
    vn = int(vptr[sample_users])
    va = synth.Csr(sample_users, n_items, vptr[: sample_users + 1].copy(), vind[:vn], vval[:vn])
    epochs = max(1, min(steps, 2))
    sample = f"first {sample_users} users ({nnz_s} ratings) of the workload, trainSGDPar P={cores}, {epochs} epoch(s)"
    kind = "port"
    secs = None
    if ol.have_ref():
        try:
            d = tempfile.mkdtemp(prefix="mfref_")
            files = synth.write_split_files(d, tr, va, va)
            res = ol.run_ref(files, os.path.join(d, "dump"), algo="mf", method="sgdpar", threads=cores, timeout=900,
                             facdim=RANK, maxiter=1, seed=1, ureg=HP["ureg"], ireg=HP["ireg"], learnrate=HP["lr"])
            for line in res["stdout"].splitlines():
                if "subIterDuration:" in line:
                    secs = float(line.split("subIterDuration:")[1].split()[0])
            kind = "reference"
            epochs = 1
            sample = f"first {sample_users} users ({nnz_s} ratings) of the workload, mf_ref --mf_method sgdpar, OMP_NUM_THREADS={cores}, epoch 0"
        except Exception as ex:  # fall back to the port
            log("mf_ref failed, using the oracle port:", repr(ex)[:200])
            secs = None
    od = ol.OracleData(tr, va, va)
    om = ol.OracleModel(od, algo="mf", facdim=RANK, maxiter=epochs, seed=1, nthreads=cores, ureg=HP["ureg"],
                        ireg=HP["ireg"], learnrate=HP["lr"])
    # ratings visited in an epoch of trainSGDPar: P schedules drawn with replacement (util.cpp:1077)
    up, ip, sched = om.dsgd_plan(cores, epochs * cores)
    users = np.repeat(np.arange(sample_users), np.diff(tr.rowptr))
    bid = up[users].astype(np.int64) * cores + ip[tr.rowind]
    blk = np.bincount(bid, minlength=cores * cores)
    visited = sum(int(blk[a * cores + b]) for s in range(epochs * cores) for a, b in sched[s])
    if secs is None:
        om.train("sgdpar")
        secs = float(np.sum(om.epoch_seconds()))
    value = visited / secs
    return dict(value=value, unit="rating-updates/s", cores=cores, kind=kind, sample=sample), secs / epochs * 1e3


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
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_users, n_items, nnz = (int(SHAPE[0] * args.scale), int(SHAPE[1] * max(args.scale, 0.05) if args.scale < 1 else SHAPE[1]),
                             int(SHAPE[2] * args.scale))
    config = {"workload": "modelMF SGD rank 64, Netflix-shaped synthetic ratings (BASELINE.json configs[1])",
              "n_users": n_users, "n_items": n_items, "rank": RANK, "learnrate": HP["lr"], "ureg": HP["ureg"],
              "ireg": HP["ireg"], "l2": "inputs larger than L2 (ratings 0.8 GB + U 123 MB), no flush"}

    import torch
    have_cuda = torch.cuda.is_available()

    if args.impl == "reference":
        if rank != 0:
            return 0
        prob = gen_problem(n_users, n_items, nnz, 20260102, "cuda" if have_cuda else "cpu")
        cb, ms = cpu_reference_arm(prob, args.steps, args.warmup)
        config["train_nnz"] = int(prob["train"][0][-1])
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

    t0 = time.time()
    if world > 1:
        prob = shared_problem(n_users, n_items, nnz, 20260102, rank, dist, f"cuda:{local_rank}")
    else:
        prob = gen_problem(n_users, n_items, nnz, 20260102, f"cuda:{local_rank}")
    ptr, ind, val = prob["train"]
    train_nnz = int(ptr[-1])
    config["train_nnz"] = train_nnz
    log(f"[rank {rank}] data: {train_nnz} train ratings, max user degree {int(np.diff(ptr).max())}, gen {time.time()-t0:.1f}s")
    torch.cuda.empty_cache()

    rng = np.random.default_rng(1)
    U0 = rng.uniform(-0.01, 0.01, size=(n_users, RANK)).astype(np.float32)
    V0 = rng.uniform(-0.01, 0.01, size=(n_items, RANK)).astype(np.float32)
    tr = Mat(n_users, n_items, prob["train"])
    va = Mat(n_users, n_items, prob["val"])
    bad_u = (np.diff(ptr) == 0).astype(np.uint8)
    bad_i = (np.bincount(ind, minlength=n_items) == 0).astype(np.uint8)

    eng = E.Engine(n_users, n_items, RANK, device=local_rank)
    P = world
    if world == 1:
        eng.upload_csr(E.TRAIN, tr, with_csc=False)
        eng.upload_csr(E.VAL, va, with_csc=False)
        eng.set_masks(bad_u, bad_i)
        eng.upload_factors(U0, V0)
        eng.sgd_plan(1)
        sched_blocks = [np.array([[0, 0]], np.int32)]
        my_nnz_per_epoch = train_nnz
    else:
        # DSGD: user stratum g pinned to rank g (only its CSR rows are uploaded), P = N item
        # blocks; after every sub-epoch the updated item block is stored straight into the next owner's V
        # over NVLink by a kernel and ordered by device-side sequence flags (matfac_b200/csrc/comm.cu)
        from matfac_b200 import dsgd
        user_part = dsgd.balanced_partition(np.diff(ptr), P)
        item_part = dsgd.balanced_partition(np.bincount(ind, minlength=n_items), P)
        mine = user_part == rank

        def local_rows(t):
            p_, i_, v_ = t
            rows = np.repeat(mine, np.diff(p_))
            lptr = np.zeros(n_users + 1, np.int64)
            np.cumsum(np.where(mine, np.diff(p_), 0), out=lptr[1:])
            return Mat(n_users, n_items, (lptr, i_[rows], v_[rows]))

        ltr, lva = local_rows(prob["train"]), local_rows(prob["val"])
        eng.upload_csr(E.TRAIN, ltr, with_csc=False)
        eng.upload_csr(E.VAL, lva, with_csc=False)
        eng.set_masks(bad_u, bad_i)
        eng.upload_factors(U0, V0)
        # in-flight budget of the shuffled kernel per rank: 8e-4 of the rank's ratings (default 2e-4, calibrated for the
        # first epochs of small matrices).  tools/dsgd_sweep.py on this matrix, 8 x 8 strata: validation RMSE after 10
        # epochs 0.5348 at 8e-4 against 0.5321 at 2e-4, epoch 3.5 ms against 6.9 ms (profiles/r1_dsgd_scaling.md)
        eng.set_option("sgd_flat_inflight_frac", 8e-4)
        eng.sgd_plan(P, np.where(mine, user_part, -1).astype(np.int32), item_part)
        eng.set_option("sgd_block_order", 1)  # shuffled inside the blocks: full concurrency
        blobs = [None] * world
        dist.all_gather_object(blobs, eng.comm_init(rank, world))
        eng.comm_connect(blobs)
        my_nnz_per_epoch = int(ltr.rowptr[-1])
        sched = dsgd.rotation_schedule(P, (args.warmup + args.steps) * P + 1)
        transport = dsgd.EngineTransport(eng, rank)

    def epoch(ep):
        if world == 1:
            eng.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, ep)
            return
        dsgd.run_steps(sched, ep * P, (ep + 1) * P, rank, transport,
                       lambda block, t: eng.sgd_subepoch(np.array([[rank, block]], np.int32), E.MF, HP["lr"], HP["ureg"],
                                                         HP["ireg"], 1, t))

    def barrier():
        if dist is not None:
            eng.comm_barrier()
        eng.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for w in range(args.warmup):
        epoch(w)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = E.launch_count()
    eng.event_record(0)
    for k in range(args.steps):
        epoch(args.warmup + k)
    eng.event_record(1)
    barrier()
    ms_dev = eng.event_elapsed_ms(0, 1)  # CUDA events on the engine's stream: kernels + exchange + waits
    launches = E.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ms_dev
    total_nnz = my_nnz_per_epoch
    nnz_per_rank = [my_nnz_per_epoch]
    if dist is not None:
        if eng.comm_error():
            raise SystemExit(f"[rank {rank}] a device-side wait timed out: the exchange schedule is broken")
        t = torch.tensor([ms_total, float(my_nnz_per_epoch)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total = float(tmax[0])
        total_nnz = int(tsum[1])
        allnnz = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(allnnz, t[1:2].clone())
        nnz_per_rank = [int(x.item()) for x in allnnz]
        # every rank publishes the item block it holds, then evaluates its own users' validation rows
        dsgd.publish(sched, (args.warmup + args.steps) * P - 1, rank, transport)
    ms_per_step = ms_total / args.steps
    value = total_nnz / (ms_per_step * 1e-3)
    ev = eng.eval(E.VAL)
    if dist is not None:
        t = torch.tensor([ev[0], ev[1]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ev = [float(t[0]), float(t[1])]
        eng.comm_barrier()
    val_rmse = float(np.sqrt(ev[0] / ev[1])) if ev[1] > 0 else float("nan")
    log(f"[rank {rank}] {ms_per_step:.3f} ms/epoch, {value/1e9:.3f} G updates/s, val RMSE after {args.warmup+args.steps} epochs {val_rmse:.4f}")

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # roofline of the SGD update kernel.  N = 1: one launch (or one per user band) per epoch, timed by the
    # events above.  N > 1: rank 0's share of the algorithmic bytes over the same wall of device time,
    # which also holds the exchange pushes and flag waits of the sub-epochs.
    peak, peak_src = measured_hbm_gbs()
    alg_bytes = (16 * RANK + 12) * float(my_nnz_per_epoch)
    achieved = alg_bytes * args.steps / (ms_dev * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": DRAM_TRAFFIC_BYTES_PER_EPOCH if world == 1 and args.scale == 1.0 else None, "peak_source": peak_src,
                "kernel": "sgd_flat_kernel<16,1,MF> + sgd_hot_kernel<4,4,MF> (hot item rows, concurrent stream)",
                "algorithmic_bytes_per_update": 16 * RANK + 12,
                "launches_per_epoch": launches / args.steps,
                "note": "achieved = algorithmic bytes (u,v read + reduced, 12 B rating record, SURVEY 8d) of one epoch on this rank / "
                        "its device time (both kernels of the epoch run concurrently inside the timed region); traffic = "
                        "dram__bytes_read+write of the two kernels per epoch from the ncu --set full captures in profiles/. "
                        "U (123 MB) and V (4.5 MB) stay in the 126 MB L2, so DRAM traffic is a tenth of the algorithmic bytes and "
                        "frac can exceed 1: the epoch is bound by L2 reductions and instruction issue, not by HBM (profiles/r1_sgd_flat.md)"}
    if world > 1:
        roofline["per_rank_nnz"] = nnz_per_rank

    # how the epoch was split between the two kernels (diagnostics ABI; outside the timed region)
    hot_info = None
    try:
        if world == 1:
            _, cold_n, lists = eng.debug_sgd_records(0, 0, with_records=False)
            st = eng.debug_sgd_hot_batch()
            hot_info = {"lists": int(len(lists)), "share_of_ratings": 1.0 - cold_n / max(train_nnz, 1),
                        "longest_list": int(lists[:, 2].max()) if len(lists) else 0, "ratings_per_round": int(st[2]),
                        "mean_user_norm_sq": float(st[0] / st[1]) if st[1] > 0 else 0.0}
    except Exception as ex:
        hot_info = {"error": repr(ex)[:200]}

    # the stratified trainer's kernel (user-major runs, concurrency capped for parity) for the record
    stratified = None
    if world == 1:
        try:
            P = 8
            prng = np.random.default_rng(7)
            eng.sgd_plan(P, prng.integers(0, P, n_users).astype(np.int32), prng.integers(0, P, n_items).astype(np.int32))
            blocks = [np.stack([np.arange(P), (np.arange(P) + s) % P], 1).astype(np.int32) for s in range(P)]
            for b in blocks:
                eng.sgd_subepoch(b, E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, 0)
            eng.event_record(2)
            for ep in range(3):
                for b in blocks:
                    eng.sgd_subepoch(b, E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, ep + 1)
            eng.event_record(3)
            ms_s = eng.event_elapsed_ms(2, 3) / 3
            stratified = {"P": P, "ms_per_epoch": ms_s, "value": train_nnz / (ms_s * 1e-3), "unit": "rating-updates/s",
                          "what": "trainSGDPar kernel: user-major runs, hot-item concurrency capped at 8 (parity with the reference's order)"}
            eng.sgd_plan(1)
        except Exception as ex:
            stratified = {"error": repr(ex)[:200]}

    # end-to-end through the C ABI with host buffers
    e2e = None
    if world == 1 and args.e2e_steps > 0:
        h = {}
        keep = []
        for name, a in (("ptr", ptr), ("ind", ind), ("val", val), ("U", U0.copy()), ("V", V0.copy())):
            h[name], t = pinned(a)
            keep.append(t)
        Uo = np.empty_like(U0); Vo = np.empty_like(V0)
        Uo, tU = pinned(Uo); Vo, tV = pinned(Vo)
        trp = Mat(n_users, n_items, (h["ptr"], h["ind"], h["val"]))
        h2d = h["ptr"].nbytes + h["ind"].nbytes + h["val"].nbytes + h["U"].nbytes + h["V"].nbytes
        d2h = Uo.nbytes + Vo.nbytes + 64
        times, parts = [], []
        for s in range(args.e2e_steps + 1):
            eng.sync()
            t1 = time.perf_counter()
            eng.upload_csr(E.TRAIN, trp, with_csc=False)
            eng.upload_factors(h["U"], h["V"])
            t2 = time.perf_counter()
            eng.sgd_plan(1)
            t3 = time.perf_counter()
            eng.sgd_epoch_flat(E.MF, HP["lr"], HP["ureg"], HP["ireg"], 1, s)
            obj = eng.eval(E.TRAIN, E.CURRENT, E.MF, False, True)
            vr = eng.eval(E.VAL)
            t4 = time.perf_counter()
            eng.L.mfb_download_factors(eng.h, E.CURRENT, Uo.ctypes.data, RANK, Vo.ctypes.data, RANK)
            eng.sync()
            t5 = time.perf_counter()
            times.append(t5 - t1)
            parts.append([t2 - t1, t3 - t2, t4 - t3, t5 - t4])
        t_e2e = float(np.median(times[1:]))
        pm = np.median(np.array(parts[1:]), axis=0) * 1e3
        e2e = {"value": train_nnz / t_e2e, "unit": "rating-updates/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e * 1e3,
               "ms_breakdown": {"upload": float(pm[0]), "plan": float(pm[1]), "epoch_and_eval": float(pm[2]), "download": float(pm[3])},
               "what": "per step: upload CSR + factors from pinned host memory, plan, 1 epoch, objective + val RMSE, download factors"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline, _ = cpu_reference_arm(prob, 1, 0)
        except Exception as ex:
            cpu_baseline = {"value": None, "unit": "rating-updates/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": "failed: " + repr(ex)[:200]}

    line = {"metric": "sgd_rating_updates_per_sec", "value": value, "unit": "rating-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "val_rmse": val_rmse}
    if e2e:
        line["e2e"] = e2e
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if hot_info:
        line["sgd_hot_rows"] = hot_info
    if stratified:
        line["stratified"] = stratified
    if world == 1 and not args.no_solvers:
        try:
            eng.close()
            torch.cuda.empty_cache()
            line["solvers"] = solver_timings(prob, local_rank, peak)
        except Exception as ex:
            line["solvers"] = {"error": repr(ex)[:300]}
    print(json.dumps(line), file=real_stdout, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
