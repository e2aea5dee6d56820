"""Multi-GPU stratified SGD (DSGD) driver: one process per GPU, user strata pinned to ranks, item
blocks handed from rank to rank after every sub-epoch.

This is the reference's own algorithm with P = number of ranks (ModelMF::trainSGDPar,
modelMF.cpp:229-304): P user parts x P item parts, a sub-epoch runs P blocks that share no user and
no item part (`sgdUpdateBlockSeq`, util.cpp:1077-1107, draws them as a random permutation).  Across
GPUs the user part of a block decides the rank, so the block a rank works on in sub-epoch t is
`sigma_t(rank)`, and after the sub-epoch the updated item rows go to the rank that owns that item
part next: `dst = sigma_{t+1}^-1(sigma_t(rank))`.  Every rank sends one block and receives one block
per sub-epoch — the routing below — and nothing else crosses the links.

The transport is pluggable: `EngineTransport` stores the rows straight into the destination's V over
NVLink from a kernel (matfac_b200/csrc/comm.cu) and orders with sequence flags on the device;
the CPU tests plug a gloo send/recv transport into the same routing.
"""
from __future__ import annotations

import numpy as np


def rotation_schedule(world: int, n_steps: int, start: int = 0) -> np.ndarray:
    """sigma_t(g) = (g + t) mod world: the Latin-square rotation (every block once per epoch)."""
    t = np.arange(start, start + n_steps)[:, None]
    return ((np.arange(world)[None, :] + t) % world).astype(np.int32)


def random_schedule(world: int, n_steps: int, seed: int) -> np.ndarray:
    """Independent random permutations per sub-epoch, as util.cpp:1077-1107 draws them."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.permutation(world) for _ in range(n_steps)]).astype(np.int32)


def balanced_partition(weights, P: int) -> np.ndarray:
    """Longest-processing-time greedy split of ids into P parts of (nearly) equal total weight; ids
    with weight 0 get part -1 (never trained).  The reference cuts a shuffled id list into equal
    COUNTS (modelMF.cpp:233-265); with one stratum per GPU the sub-epoch lasts as long as its heaviest
    block, so the multi-GPU driver balances ratings instead and spreads the heaviest rows over the
    parts (the hottest item row of a block bounds that block's useful concurrency)."""
    import heapq
    w = np.asarray(weights, dtype=np.int64)
    part = np.full(w.shape[0], -1, np.int32)
    order = np.argsort(-w, kind="stable")
    heap = [(0, p) for p in range(P)]
    heapq.heapify(heap)
    for i in order:
        if w[i] <= 0:
            break
        load, p = heapq.heappop(heap)
        part[i] = p
        heapq.heappush(heap, (load + int(w[i]), p))
    return part


def route(schedule: np.ndarray, t: int, rank: int):
    """(block, dst, src) of `rank` in sub-epoch t: the item part it updates, the rank that needs that
    part in sub-epoch t+1 (-1 at the end of the schedule) and the rank whose sub-epoch t-1 output it
    must have received before starting (-1 for t = 0).  dst / src equal to `rank` mean no transfer."""
    block = int(schedule[t, rank])
    dst = -1
    if t + 1 < schedule.shape[0]:
        dst = int(np.nonzero(schedule[t + 1] == block)[0][0])
    src = -1
    if t > 0:
        src = int(np.nonzero(schedule[t - 1] == block)[0][0])
    return block, dst, src


class EngineTransport:
    """Peer-memory transport over the engine's C ABI."""

    def __init__(self, eng, rank):
        self.eng, self.rank = eng, rank

    def wait(self, src, seq, block):
        self.eng.comm_wait_block(src, seq)

    def push(self, block, dst, seq):
        self.eng.dsgd_push_block(block, dst, seq)

    def publish_all(self, block):
        self.eng.dsgd_push_block(block, -1, 0)

    def barrier(self):
        self.eng.comm_barrier()


def run_steps(schedule, t0, t1, rank, transport, update, seq_base=0):
    """Run sub-epochs [t0, t1) of `schedule` on this rank.  update(block, t) performs the SGD
    sub-epoch on block (rank, block).  Sequence numbers are seq_base + t + 1 for the push that ends
    sub-epoch t, so a schedule may be continued across calls."""
    for t in range(t0, t1):
        block, dst, src = route(schedule, t, rank)
        if src >= 0 and src != rank:
            transport.wait(src, seq_base + t, block)
        update(block, t)
        if dst >= 0 and dst != rank:
            transport.push(block, dst, seq_base + t + 1)


def publish(schedule, t_last, rank, transport):
    """After sub-epoch t_last: every rank stores the block it holds into all peers, then barrier —
    all ranks see the complete item matrix (needed by the per-epoch evaluation)."""
    transport.publish_all(int(schedule[t_last, rank]))
    transport.barrier()


# ---------------------------------------------------------------------------------------------------------
# The reference's own plan (Model::trainSGDPar) from the product's host library, and a driver that runs the
# ranks living in this process — one rank per process under torchrun (bench.py), or all ranks in one
# process (the host classes' way; also how a 1-GPU box exercises the whole multi-rank protocol).
def reference_plan(n_users, n_items, invalid_users, invalid_items, seed, P, n_subepochs):
    """(user_part, item_part, schedule[n_subepochs][P]) exactly as ModelMF::trainSGDPar draws them: ids shuffled by
    mt19937(seed) and cut into P parts of equal COUNT (modelMF.cpp:229-265), one `sgdUpdateBlockSeq` permutation per
    sub-epoch from the same engine (util.cpp:1077-1107).  schedule[t][g] = item part of user part g in sub-epoch t.
    Computed by libmatfac_host.so (matfac_b200/host/capi.cpp: mfh_sgd_plan) — the same code `mf` runs."""
    import ctypes as C
    import os
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmatfac_host.so"))
    vp = C.c_void_p
    lib.mfh_sgd_plan.argtypes = [C.c_int32, C.c_int32, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
    bu = np.ascontiguousarray(invalid_users, np.uint8)
    bi = np.ascontiguousarray(invalid_items, np.uint8)
    up = np.zeros(n_users, np.int32)
    ip = np.zeros(n_items, np.int32)
    pairs = np.zeros((max(n_subepochs, 1), P, 2), np.int32)
    rc = lib.mfh_sgd_plan(n_users, n_items, bu.ctypes.data, bi.ctypes.data, seed, P, n_subepochs, up.ctypes.data,
                          ip.ctypes.data, pairs.ctypes.data)
    if rc != 0:
        raise RuntimeError("mfh_sgd_plan failed")
    sched = np.zeros((n_subepochs, P), np.int32)
    for t in range(n_subepochs):
        sched[t, pairs[t, :, 0]] = pairs[t, :, 1]
    return up, ip, sched


class _Mat:
    def __init__(self, nrows, ncols, ptr, ind, val):
        self.nrows, self.ncols, self.rowptr, self.rowind, self.rowval = nrows, ncols, ptr, ind, val
        self.colptr = self.colind = self.colval = None


def local_rows(n_users, n_items, csr, mine):
    """The CSR restricted to the users flagged in `mine` (other rows empty) — what a rank uploads."""
    ptr, ind, val = csr
    deg = np.diff(ptr)
    rows = np.repeat(mine, deg)
    lptr = np.zeros(n_users + 1, np.int64)
    np.cumsum(np.where(mine, deg, 0), out=lptr[1:])
    return _Mat(n_users, n_items, lptr, ind[rows], val[rows])


class Dsgd:
    """Stratified SGD over `world` ranks: user part g pinned to rank g, item parts handed from rank to rank.

    plan = "reference": the reference's partitions and per-sub-epoch random permutations (reference_plan) — the
           configuration that is checked against the oracle's trainSGDPar;
    plan = "balanced":  rating-balanced parts + Latin-square rotation (every block exactly once per epoch) — an
           option for speed, not the reference's schedule.
    local_ranks: the ranks this process drives ({rank: device ordinal}).  With all ranks local the engines are connected
    in-process (mfb_comm_connect_local); otherwise `exchange(blob) -> [blobs of all ranks]` swaps the CUDA IPC handles
    (e.g. torch.distributed.all_gather_object) and there must be exactly one local rank.
    """

    def __init__(self, n_users, n_items, dim, world, local_ranks, train, val, U0, V0, bad_u, bad_i, n_subepochs,
                 plan="reference", seed=1, exchange=None, options=None, variant=0, aux=None, block_order=1):
        from . import engine as E
        self.E, self.world, self.P, self.variant = E, world, world, variant
        self.n_users, self.n_items = n_users, n_items
        if plan == "reference":
            self.user_part, self.item_part, self.sched = reference_plan(n_users, n_items, bad_u, bad_i, seed, world, n_subepochs)
        elif plan == "balanced":
            self.user_part = balanced_partition(np.where(bad_u != 0, 0, np.diff(train[0])), world)
            self.item_part = balanced_partition(np.where(bad_i != 0, 0, np.bincount(train[1], minlength=n_items)), world)
            self.sched = rotation_schedule(world, n_subepochs)
        else:
            raise ValueError(plan)
        self.plan = plan
        self.engines, self.transports, self.local_nnz = {}, {}, {}
        for r, dev in sorted(local_ranks.items()):
            mine = self.user_part == r
            eng = E.Engine(n_users, n_items, dim, device=dev)
            ltr = local_rows(n_users, n_items, train, mine)
            eng.upload_csr(E.TRAIN, ltr, with_csc=False)
            if val is not None:
                eng.upload_csr(E.VAL, local_rows(n_users, n_items, val, mine), with_csc=False)
            eng.set_masks(bad_u, bad_i)
            if aux is not None:
                eng.set_aux(variant, *aux)
            eng.upload_factors(U0, V0)
            for k, v in (options or {}).items():
                eng.set_option(k, v)
            eng.set_option("sgd_shuffle_seed", seed)
            eng.sgd_plan(world, np.where(mine, self.user_part, -1).astype(np.int32), self.item_part)
            # 1 = shuffled inside the blocks (DSGD's usual formulation), 0 = user-major runs in CSR order (the
            # reference's visiting order inside a block, modelMF.cpp:279-281)
            eng.set_option("sgd_block_order", block_order)
            self.engines[r] = eng
            self.transports[r] = EngineTransport(eng, r)
            self.local_nnz[r] = int(ltr.rowptr[-1])
        if world > 1:
            if len(self.engines) == world:
                E.connect_local([self.engines[r] for r in range(world)])
            else:
                assert len(self.engines) == 1 and exchange is not None
                (r, eng), = self.engines.items()
                eng.comm_connect(exchange(eng.comm_init(r, world)))
        self.t = 0

    def block_nnz(self, t0, t1):
        """ratings the local ranks visit in sub-epochs [t0, t1) (a block may be drawn twice or not at all per epoch
        under the reference's schedule)"""
        return {r: sum(e.sgd_block_nnz([[r, int(self.sched[t, r])]]) for t in range(t0, t1)) for r, e in self.engines.items()}

    def run(self, t0, t1, lr, ureg, ireg, seed=1):
        """Sub-epochs [t0, t1): for every sub-epoch, every local rank waits for its item block, updates block
        (rank, sched[t][rank]) and pushes the item rows to their next owner.  Asynchronous (device-side flags)."""
        E = self.E
        for t in range(t0, t1):
            for r, eng in self.engines.items():
                run_steps(self.sched, t, t + 1, r, self.transports[r] if self.world > 1 else _NoTransport(),
                          lambda block, tt, eng=eng, r=r: eng.sgd_subepoch(np.array([[r, block]], np.int32), self.variant, lr, ureg,
                                                                           ireg, seed, tt))
        self.t = t1

    def publish(self):
        """every rank stores the item block it holds into all peers, then barrier: all ranks see all of V"""
        if self.world == 1 or self.t == 0:
            return
        for r in self.engines:
            self.transports[r].publish_all(int(self.sched[self.t - 1, r]))
        for r in self.engines:
            self.transports[r].barrier()

    def barrier(self):
        if self.world > 1:
            for r in self.engines:
                self.transports[r].barrier()

    def eval_sums(self, which=1):
        """(sse, count) of the local ranks' rows (call publish() first)"""
        s = np.zeros(2)
        for eng in self.engines.values():
            o = eng.eval(which, self.E.CURRENT, self.variant)
            s += o[:2]
        return s

    def close(self):
        if self.world > 1:
            for eng in self.engines.values():
                eng.sync()
            for eng in self.engines.values():
                eng.comm_disconnect()
        for eng in self.engines.values():
            eng.close()
        self.engines = {}


class _NoTransport:
    def wait(self, *a): pass
    def push(self, *a): pass
    def publish_all(self, *a): pass
    def barrier(self): pass
