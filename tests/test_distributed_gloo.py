"""The N>1 path on CPU: world_size 2 over gloo, each rank driving the host-emulation build of the library.

* root sharding (BASELINE configs 2-4): the union of the ranks' results equals the single-process result
  bit-for-bit, because Philox streams are keyed by GLOBAL root index; no data-path collective.
* root-parallel search of one root (config 5): per-rank partial sums merged with one allreduce equal the
  single-process sums, and the decision is the first argmax of sum(NQ)/sum(N)."""
import os
import socket

import numpy as np
import pytest

import mcts_jl_b200 as m
from mcts_jl_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libmctsb200_emu.so")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = m.Context(0, lib_path=EMU)
        gw = m.SimpleGridWorld()
        states = [(1 + i % 10, 1 + (i // 10) % 10) for i in range(37)]  # ragged split: 19 + 18
        uct = m.MCTSSolver(n_iterations=60, depth=10, exploration_constant=5.0, rng=4)
        act, bq, _ = D.plan_sharded(ctx, uct, gw, states)
        dpw = m.DPWSolver(n_iterations=60, depth=10, rng=4)
        act2, bq2, _ = D.plan_sharded(ctx, dpw, gw, states)
        a, best, qs, sum_n, _ = D.plan_root_parallel(ctx, m.MCTSSolver(n_iterations=40, depth=10, exploration_constant=5.0, rng=6), gw, (1, 1), 24)
        q.put((rank, act, bq, act2, bq2, a, best, qs, sum_n))
    finally:
        dist.destroy_process_group()


def test_world_size_2_matches_single_process(emu_lib_path):
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # single-process reference
    ctx = m.Context(0, lib_path=emu_lib_path)
    gw = m.SimpleGridWorld()
    states = [(1 + i % 10, 1 + (i // 10) % 10) for i in range(37)]
    act, bq, _ = D.plan_sharded(ctx, m.MCTSSolver(n_iterations=60, depth=10, exploration_constant=5.0, rng=4), gw, states, rank=0, world=1)
    act2, bq2, _ = D.plan_sharded(ctx, m.DPWSolver(n_iterations=60, depth=10, rng=4), gw, states, rank=0, world=1)
    a, best, qs, sum_n, _ = D.plan_root_parallel(ctx, m.MCTSSolver(n_iterations=40, depth=10, exploration_constant=5.0, rng=6), gw, (1, 1), 24,
                                                 rank=0, world=1)
    for r in results:
        assert np.array_equal(r[1], act) and np.array_equal(r[2].view(np.int64), bq.view(np.int64))
        assert np.array_equal(r[3], act2) and np.array_equal(r[4].view(np.int64), bq2.view(np.int64))
        assert r[5] == a and np.array_equal(r[8], sum_n)  # visit counts are integers: exact under any summation order
        assert np.allclose(r[7], qs, rtol=1e-12) and abs(r[6] - best) <= 1e-12 * max(1.0, abs(best))
    assert int(sum_n.sum()) >= 24 * 40


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 65536):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
