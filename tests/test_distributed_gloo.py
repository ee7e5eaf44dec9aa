"""The N>1 path on CPU: world_size 2 over gloo, each rank driving the host-emulation build of the library.

* root sharding (BASELINE configs 2-4): the union of the ranks' results equals the single-process result
  bit-for-bit, because Philox streams are keyed by GLOBAL root index; no data-path collective.
* root-parallel search of one root (config 5): per-rank partial sums (integer words) merged with one allreduce equal the
  single-process sums bit for bit, and the decision is the first argmax of sum(NQ)/sum(N).
* the same two paths through ONE multi-device ctx (mctsb200_ctx_create_multi: fan-out inside the library), here with several
  ctxs on the emulator's single device."""
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
        # the partial sums travel as integer words (include/mctsb200.h: mctsb200_plan_root_parallel): the merged sums, hence the
        # ensemble's Q values, do not depend on how the trees were split over the ranks -- bit for bit
        assert r[5] == a and np.array_equal(r[8], sum_n)
        assert np.array_equal(r[7].view(np.int64), qs.view(np.int64)) and r[6] == best
    assert int(sum_n.sum()) >= 24 * 40


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 65536):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_multi_device_ctx_matches_single_ctx(emu_lib_path, oracle):
    """mctsb200_ctx_create_multi: the same entry points, sharded inside the library. (The emulator has one device, so the three
    children share it and the host adds the root-parallel words instead of NCCL; on B200 boxes test_gpu_parity.py runs the same
    comparison over real devices.)"""
    from helpers import assert_trees_equal, bits

    gw = m.SimpleGridWorld()
    one = m.Context(0, lib_path=emu_lib_path)
    multi = m.Context([0, 0, 0], lib_path=emu_lib_path)
    states = [(1 + i % 10, 1 + (i // 10) % 10) for i in range(41)]  # ragged: 14 + 14 + 13
    roots = gw.encode_states(states)
    for solver in (m.MCTSSolver(n_iterations=60, depth=10, exploration_constant=5.0, rng=4), m.DPWSolver(n_iterations=60, depth=10, rng=4)):
        cfg, keep = m.build_cfg(solver, gw)
        for c in (one, multi):
            c.configure(gw.mdp_id, gw.params())
        a1, q1, s1, st1 = one.plan(cfg, gw.mdp_id, roots, 500, want_status=True)
        a2, q2, s2, st2 = multi.plan(cfg, gw.mdp_id, roots, 500, want_status=True)
        assert np.array_equal(a1, a2) and np.array_equal(bits(q1), bits(q2)) and np.array_equal(s1, s2)
        for k in ("n_simulations", "n_tree_levels", "n_rollout_steps", "n_state_inserts", "algorithmic_bytes", "n_trees"):
            assert getattr(st1, k) == getattr(st2, k), k
        for i in (0, 13, 14, 27, 28, 40):  # tree queries address tree i of the whole call, whichever device holds it
            assert_trees_equal(one.tree_export(i), multi.tree_export(i), f"tree {i}")
            r1, r2 = one.root_stats(i), multi.root_stats(i)
            assert np.array_equal(r1["n"], r2["n"]) and np.array_equal(bits(r1["q"]), bits(r2["q"]))
        with pytest.raises(m.MCTSB200Error):
            multi.root_stats(41)
        # closed-loop episodes shard the same way
        e1 = one.run_episodes(cfg, gw.mdp_id, roots[:10], 5, env_seed=3)
        e2 = multi.run_episodes(cfg, gw.mdp_id, roots[:10], 5, env_seed=3)
        assert np.array_equal(bits(e1[0]), bits(e2[0])) and np.array_equal(e1[1], e2[1]) and np.array_equal(e1[2], e2[2])
        # root-parallel ensemble: 25 trees split 9 + 8 + 8, merged sums equal the single ctx's and the oracle's bit for bit
        root = gw.encode_states([(1, 1)])
        n1, nq1, _ = one.plan_root_parallel(cfg, gw.mdp_id, root, 25, tree_index_base=7)
        n2, nq2, _ = multi.plan_root_parallel(cfg, gw.mdp_id, root, 25, tree_index_base=7)
        oracle.configure(gw.mdp_id, gw.params())
        n3, nq3, _ = oracle.plan_root_parallel(cfg, gw.mdp_id, root, 25, tree_index_base=7)
        assert np.array_equal(n1, n2) and np.array_equal(bits(nq1), bits(nq2))
        assert np.array_equal(n1, n3) and np.array_equal(bits(nq1), bits(nq3))
        # and the words are what a host with its own transport would add up
        w = one.root_sums_words()
        sn, snq = one.root_sums_merge(w, 4)
        assert np.array_equal(sn, n1) and np.array_equal(bits(snq), bits(nq1))
        del keep
    # what cannot work across devices says so
    with pytest.raises(m.MCTSB200Error) as ei:
        multi.set_progress_callback(lambda d, t: 0, every=1)
    assert ei.value.status == m.abi.ERR_UNSUPPORTED
    one.close()
    multi.close()


def test_root_sums_fixed_point_is_order_independent(emu_lib_path):
    """The fixed-point words: adding partial sums in any grouping gives the same doubles (the property the allreduce relies on)."""
    ctx = m.Context(0, lib_path=emu_lib_path)
    gw = m.SimpleGridWorld()
    ctx.configure(gw.mdp_id, gw.params())
    cfg, keep = m.build_cfg(m.MCTSSolver(n_iterations=80, depth=12, exploration_constant=5.0, rng=11), gw)
    root = gw.encode_states([(2, 3)])
    whole_n, whole_nq, _ = ctx.plan_root_parallel(cfg, gw.mdp_id, root, 32, tree_index_base=0)
    words = np.zeros(21, dtype=np.int64)
    for lo, hi in ((20, 32), (0, 3), (3, 20)):  # any split, any order
        ctx.plan_root_parallel(cfg, gw.mdp_id, root, hi - lo, tree_index_base=lo)
        words += ctx.root_sums_words()
    sn, snq = ctx.root_sums_merge(words, 4)
    assert np.array_equal(sn, whole_n) and np.array_equal(snq.view(np.int64), whole_nq.view(np.int64))
    assert np.all(np.isfinite(whole_nq)) and abs(whole_nq).max() > 0
    del keep
    ctx.close()
