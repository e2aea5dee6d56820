"""Host-side logic of the multi-GPU stratified SGD driver (matfac_b200/dsgd.py) on CPU: the routing
of item blocks between ranks, run by two gloo processes with a send/recv transport in place of the
peer-memory one, must reproduce a single-process execution of the same schedule."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matfac_b200 import dsgd  # noqa: E402


def test_route_is_consistent_for_random_schedules():
    for world in (2, 3, 8):
        sched = dsgd.random_schedule(world, 40, seed=world)
        for t in range(40):
            assert sorted(sched[t]) == list(range(world))  # conflict-free: a permutation (util.cpp:1077-1107)
            for g in range(world):
                block, dst, src = dsgd.route(sched, t, g)
                if t + 1 < 40:
                    assert sched[t + 1, dst] == block
                    assert dsgd.route(sched, t + 1, dst)[2] == g  # the receiver expects exactly this sender
                else:
                    assert dst == -1
                if t == 0:
                    assert src == -1


def test_rotation_visits_every_block_once_per_epoch():
    s = dsgd.rotation_schedule(4, 8)
    for g in range(4):
        assert sorted(s[:4, g]) == [0, 1, 2, 3] and sorted(s[4:, g]) == [0, 1, 2, 3]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _serial(sched, world, n_items, item_part, epochs):
    V = np.zeros((n_items, 4))
    for t in range(sched.shape[0]):
        for g in range(world):
            rows = np.nonzero(item_part == sched[t, g])[0]
            V[rows] = V[rows] * 1.01 + (g + 1) * (t + 1)
    return V


def _worker(rank, world, port, sched, n_items, item_part, steps_per_epoch, out):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    V = np.zeros((n_items, 4))
    pending = []

    class GlooTransport:
        def wait(self, src, seq, block):
            rows = np.nonzero(item_part == block)[0]
            buf = torch.zeros(len(rows), 4, dtype=torch.float64)
            dist.recv(buf, src=src, tag=seq)
            V[rows] = buf.numpy()

        def push(self, block, dst, seq):
            rows = np.nonzero(item_part == block)[0]
            pending.append(dist.isend(torch.from_numpy(V[rows].copy()), dst=dst, tag=seq))

        def publish_all(self, block):
            rows = np.nonzero(item_part == block)[0]
            mine = torch.zeros(n_items, 4, dtype=torch.float64)
            mine[rows] = torch.from_numpy(V[rows])
            mask = torch.zeros(n_items, 1, dtype=torch.float64)
            mask[rows] = 1
            dist.all_reduce(mine)
            dist.all_reduce(mask)
            assert float(mask.min()) == 1.0 and float(mask.max()) == 1.0  # every block held by exactly one rank
            V[:] = mine.numpy()

        def barrier(self):
            dist.barrier()

    def update(block, t):
        rows = np.nonzero(item_part == block)[0]
        V[rows] = V[rows] * 1.01 + (rank + 1) * (t + 1)

    tr = GlooTransport()
    n = sched.shape[0]
    for t0 in range(0, n, steps_per_epoch):
        dsgd.run_steps(sched, t0, min(t0 + steps_per_epoch, n), rank, tr, update)
        dsgd.publish(sched, min(t0 + steps_per_epoch, n) - 1, rank, tr)
    for p in pending:
        p.wait()
    if rank == 0:
        np.save(out, V)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["rotation", "random"])
def test_two_rank_gloo_run_matches_serial_execution(tmp_path, kind):
    import torch.multiprocessing as mp
    world, n_items, epochs = 2, 37, 3
    item_part = (np.arange(n_items) * 7 % world).astype(np.int32)
    sched = dsgd.rotation_schedule(world, world * epochs) if kind == "rotation" else dsgd.random_schedule(world, world * epochs, 5)
    out = str(tmp_path / "v.npy")
    mp.spawn(_worker, args=(world, _free_port(), sched, n_items, item_part, world, out), nprocs=world, join=True)
    np.testing.assert_allclose(np.load(out), _serial(sched, world, n_items, item_part, epochs), rtol=0, atol=0)


def test_balanced_partition_equalises_ratings_and_spreads_the_head():
    rng = np.random.default_rng(0)
    w = rng.zipf(1.3, 5000).clip(max=4000)
    w[::50] = 0  # ids without ratings are never trained
    for P in (2, 8):
        part = dsgd.balanced_partition(w, P)
        assert np.all(part[w == 0] == -1) and np.all(part[w > 0] >= 0) and part.max() == P - 1
        loads = np.bincount(part[part >= 0], weights=w[part >= 0], minlength=P)
        assert loads.max() - loads.min() <= w.max()
        heavy = np.argsort(-w, kind="stable")[:P]
        assert sorted(part[heavy]) == list(range(P))  # the P heaviest rows land in P different parts
