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
