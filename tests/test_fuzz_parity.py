"""Randomised differential test: the library's kernels (host emulation here, the CUDA build under -m gpu) against the oracle over
random models, solver options and call sequences -- every field of every compared tree bit for bit. The cases are drawn from fixed
seeds, so a failure names a reproducible case. The hand-written matrices of test_host_emu_parity.py / test_gpu_parity.py cover the
options one at a time; this covers their combinations (and combinations are where the dispatch between the three kernel families,
the kept-tree sizing and the table-driven hooks meet)."""
import os

import numpy as np
import pytest

import mcts_jl_b200 as m
from helpers import run_both, run_episodes_both


SCALE = int(os.environ.get("MCTSB200_FUZZ_SCALE", "1"))  # MCTSB200_FUZZ_SCALE=10: ten times as many cases (stress runs)


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def ctx(request):
    return request.getfixturevalue("emu_ctx" if request.param == "emu" else "gpu_ctx")


def _random_gridworld(rng):
    nx, ny = int(rng.integers(2, 14)), int(rng.integers(2, 14))  # (above 127 cells the lane-per-tree kernels hand over to the tile kernel)
    cells = [(x, y) for x in range(1, nx + 1) for y in range(1, ny + 1)]
    k = int(rng.integers(1, min(5, len(cells)) + 1))
    picks = [cells[i] for i in rng.choice(len(cells), size=k, replace=False)]
    rewards = {c: float(rng.choice([-10.0, -5.0, 3.0, 10.0, 0.5])) for c in picks}
    term = set(c for c in picks if rng.random() < 0.7)
    tprob = float(rng.choice([1.0, 0.7, 0.4, 0.25]))
    return m.SimpleGridWorld(size=(nx, ny), rewards=rewards, terminate_from=term, tprob=tprob, discount=float(rng.choice([0.95, 0.9, 1.0])))


def _random_estimate(rng, mdp, dense=True):
    kind = rng.integers(0, 6 if dense else 4)
    if kind == 0:
        return m.RolloutEstimator(m.RandomSolver(), max_depth=int(rng.choice([-1, 3, 10, 50])), eps=float(rng.choice([0.0, 0.0, 0.2])))
    if kind == 1:
        return float(rng.choice([0.0, -1.5, 2.25]))
    if kind == 2 and dense:
        return m.RolloutEstimator(m.ConstantPolicy(mdp.actions[int(rng.integers(0, len(mdp.actions)))]), max_depth=int(rng.choice([5, 50])))
    if kind == 3 and dense:
        return m.RolloutEstimator(m.SeekTarget((int(rng.integers(1, mdp.size[0] + 1)), int(rng.integers(1, mdp.size[1] + 1)))))
    if kind == 4:
        acts = list(mdp.actions)
        return m.RolloutEstimator(lambda s, acts=acts: acts[(s[0] * 7 + s[1] * 3) % len(acts)], max_depth=12)  # FunctionPolicy -> table
    if kind == 5:
        return lambda mdp_, s, d: 0.25 * s[0] - 0.5 * s[1]  # Function-valued estimate -> value table
    return m.RolloutEstimator(m.RandomSolver())


def _random_solver(rng, mdp, dense=True):
    common = dict(n_iterations=int(rng.integers(1, 120)), depth=int(rng.integers(1, 16)), rng=int(rng.integers(0, 2**31)),
                  exploration_constant=float(rng.choice([0.0, 0.5, 1.0, 5.0, 30.0])), estimate_value=_random_estimate(rng, mdp, dense))
    if rng.random() < 0.3:
        common["init_Q"] = float(rng.choice([-2.0, 0.0, 4.0]))
        common["init_N"] = int(rng.integers(0, 4))
    elif dense and rng.random() < 0.2:
        common["init_Q"] = lambda mdp_, s, a: 0.125 * (s[0] - s[1]) + (1.0 if a == "up" else 0.0)  # Function-valued -> tables
        common["init_N"] = lambda mdp_, s, a: (s[0] + s[1]) % 3
    if rng.random() < 0.15 and isinstance(common["estimate_value"], m.RolloutEstimator):
        common["rollouts_per_leaf"] = int(rng.choice([2, 3, 8]))  # leaf-parallel rollouts, averaged in stream order
    if rng.random() < 0.25:
        common["lanes_per_tree"] = int(rng.choice([1, 2, 4, 8, 16, 32]))  # tile width (results do not depend on it)
    if rng.random() < 0.5:
        return m.MCTSSolver(reuse_tree=bool(rng.random() < 0.4), enable_tree_vis=bool(rng.random() < 0.2), **common)
    kw = dict(k_action=float(rng.choice([1.0, 4.0, 10.0])), alpha_action=float(rng.choice([0.1, 0.5, 0.8])),
              k_state=float(rng.choice([1.0, 3.0, 10.0])), alpha_state=float(rng.choice([0.1, 0.5])),
              enable_action_pw=bool(rng.random() < 0.8), enable_state_pw=bool(rng.random() < 0.8),
              check_repeat_state=bool(rng.random() < 0.75), keep_tree=bool(rng.random() < 0.4), keep_tree_calls=6)
    if dense and rng.random() < 0.15:
        acts = list(mdp.actions)
        kw["next_action"] = lambda mdp_, s, snode, acts=acts: acts[(s[0] + 2 * s[1]) % len(acts)]
    return m.DPWSolver(**kw, **common)


def _random_roots(rng, mdp, n):
    cells = [(x, y) for x in range(1, mdp.size[0] + 1) for y in range(1, mdp.size[1] + 1)]
    return [cells[int(i)] for i in rng.integers(0, len(cells), size=n)]


@pytest.mark.parametrize("seed", range(200 * SCALE))
def test_fuzz_gridworld(ctx, oracle, seed, monkeypatch):
    rng = np.random.default_rng(1000 + seed)
    if rng.random() < 0.3:
        monkeypatch.setenv("MCTSB200_FAST", "0")  # the tile kernel on the dense-state model
    mdp = _random_gridworld(rng)
    solver = _random_solver(rng, mdp)
    roots = _random_roots(rng, mdp, int(rng.integers(1, 40)))
    if rng.random() < 0.04:  # visit counts beyond the 2 048 UCB table entries the fast kernels keep in shared memory
        solver.n_iterations = int(rng.integers(2100, 2700))
        roots = roots[:3]
    ctx.clear_trees()
    oracle.clear_trees()
    kept = getattr(solver, "keep_tree", False) or getattr(solver, "reuse_tree", False)
    for call in range(3 if kept else 1):
        run_both(ctx, oracle, solver, mdp, roots, base=int(rng.integers(0, 1000)) if not kept else 5,
                 compare_trees=range(0, len(roots), max(1, len(roots) // 4)))
        if getattr(solver, "enable_tree_vis", False):  # tree._vis_stats (src/vanilla.jl:328-334)
            for i in range(0, len(roots), max(1, len(roots) // 3)):
                assert ctx.tree_vis_stats(i) == oracle.tree_vis_stats(i)
        # an episode-like move: every root steps to a neighbouring cell (kept trees continue, or start a new root node)
        roots = [(min(max(x + int(rng.integers(-1, 2)), 1), mdp.size[0]), min(max(y + int(rng.integers(-1, 2)), 1), mdp.size[1])) for x, y in roots]
    ctx.clear_trees()
    oracle.clear_trees()


@pytest.mark.parametrize("seed", range(60 * SCALE))
def test_fuzz_mountaincar(ctx, oracle, seed):
    rng = np.random.default_rng(5000 + seed)
    mdp = m.MountainCar()
    solver = _random_solver(rng, mdp, dense=False)
    if isinstance(solver, m.DPWSolver) and not solver.check_repeat_state:
        solver.keep_tree = False  # (unmerged kept trees of a hashed model: covered by the hand-written tests; sizes differ per call here)
    roots = [(float(rng.uniform(-1.1, 0.45)), float(rng.uniform(-0.06, 0.06))) for _ in range(int(rng.integers(1, 12)))]
    ctx.clear_trees()
    oracle.clear_trees()
    kept = getattr(solver, "keep_tree", False) or getattr(solver, "reuse_tree", False)
    lanes = int(rng.choice([0, 1, 2, 8]))
    for call in range(2 if kept else 1):
        run_both(ctx, oracle, solver, mdp, roots, base=3, lanes=lanes or None)
    ctx.clear_trees()
    oracle.clear_trees()


@pytest.mark.parametrize("seed", range(40 * SCALE))
def test_fuzz_box(ctx, oracle, seed):
    from test_host_emu_parity import _run_box_both

    rng = np.random.default_rng(9000 + seed)
    mdp = m.GaussianBox(drift=float(rng.choice([0.0, 0.2, 0.5])), cost=float(rng.choice([0.0, 1.0])),
                        var=tuple(float(v) for v in rng.choice([0.01, 0.1, 0.3], size=3)))
    est = rng.integers(0, 3)
    solver = m.DPWSolver(n_iterations=int(rng.integers(1, 200)), depth=int(rng.integers(1, 9)), rng=int(rng.integers(0, 2**31)),
                         exploration_constant=float(rng.choice([0.0, 1.0, 8.0])), k_action=float(rng.choice([2.0, 10.0])),
                         alpha_action=float(rng.choice([0.25, 0.5])), k_state=float(rng.choice([1.0, 10.0])), alpha_state=float(rng.choice([0.2, 0.5])),
                         check_repeat_action=bool(rng.random() < 0.7), check_repeat_state=bool(rng.random() < 0.7),
                         enable_state_pw=bool(rng.random() < 0.8), keep_tree=bool(rng.random() < 0.4), keep_tree_calls=3,
                         estimate_value=(0.5 if est == 0 else m.RolloutEstimator(m.RandomSolver(), max_depth=int(rng.choice([-1, 4])))))
    roots = [tuple(float(v) for v in rng.uniform(-4, 4, size=3)) for _ in range(int(rng.integers(1, 8)))]
    ctx.clear_trees()
    oracle.clear_trees()
    for call in range(2 if solver.keep_tree else 1):
        _run_box_both(ctx, oracle, solver, mdp, roots, base=7, lanes=int(rng.choice([1, 4, 32])) if rng.random() < 0.5 else None)
    ctx.clear_trees()
    oracle.clear_trees()


@pytest.mark.parametrize("seed", range(40 * SCALE))
def test_fuzz_episodes(ctx, oracle, seed):
    rng = np.random.default_rng(7000 + seed)
    mdp = _random_gridworld(rng)
    solver = _random_solver(rng, mdp)
    if hasattr(solver, "enable_tree_vis"):
        solver.enable_tree_vis = False
    starts = [s for s in _random_roots(rng, mdp, int(rng.integers(1, 30)))]
    run_episodes_both(ctx, oracle, solver, mdp, starts, max_steps=int(rng.integers(1, 6)), env_seed=int(rng.integers(0, 1000)), base=int(rng.integers(0, 50)))


@pytest.mark.parametrize("seed", range(40 * SCALE))
def test_fuzz_root_parallel(ctx, oracle, seed):
    """mctsb200_plan_root_parallel: an ensemble of trees of ONE root, per-action sums of N and N*Q -- the exact integer words summed in
    any split and order give the oracle's doubles."""
    rng = np.random.default_rng(3000 + seed)
    mdp = _random_gridworld(rng) if rng.random() < 0.7 else m.MountainCar()
    dense = isinstance(mdp, m.SimpleGridWorld)
    solver = m.MCTSSolver(n_iterations=int(rng.integers(1, 150)), depth=int(rng.integers(1, 14)), rng=int(rng.integers(0, 2**31)),
                          exploration_constant=float(rng.choice([0.5, 5.0])), estimate_value=_random_estimate(rng, mdp, dense))
    root = _random_roots(rng, mdp, 1)[0] if dense else (float(rng.uniform(-1.0, 0.4)), float(rng.uniform(-0.05, 0.05)))
    n_trees, base = int(rng.integers(1, 60)), int(rng.integers(0, 100))
    cfg, keep = m.build_cfg(solver, mdp)
    ctx.configure(mdp.mdp_id, mdp.params())
    oracle.configure(mdp.mdp_id, mdp.params())
    r = mdp.encode_states([root])
    n1, nq1, _ = ctx.plan_root_parallel(cfg, mdp.mdp_id, r, n_trees, tree_index_base=base)
    n2, nq2, _ = oracle.plan_root_parallel(cfg, mdp.mdp_id, r, n_trees, tree_index_base=base, n_threads=3)
    assert np.array_equal(n1, n2) and np.array_equal(nq1.view(np.int64), nq2.view(np.int64))
    # any split of the ensemble, merged through the integer words, gives the same bits
    cut = int(rng.integers(0, n_trees + 1))
    words = np.zeros(5 * len(n1) + 1, dtype=np.int64)
    for lo, hi in ((0, cut), (cut, n_trees)):
        if hi > lo:
            ctx.plan_root_parallel(cfg, mdp.mdp_id, r, hi - lo, tree_index_base=base + lo)
            words = words + ctx.root_sums_words()
    n3, nq3 = ctx.root_sums_merge(words, len(n1))
    assert np.array_equal(n3, n1) and np.array_equal(nq3.view(np.int64), nq1.view(np.int64))
    del keep


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(16 * SCALE))
def test_fuzz_plugins(gpu_ctx, oracle, seed):
    """The run-time plugins (NVRTC-compiled tile kernel; device only) under random DPW / UCT options against their oracle restatements."""
    from test_host_emu_parity import _run_box_both

    rng = np.random.default_rng(11000 + seed)
    which = int(rng.integers(0, 3))
    roots = [(float(rng.uniform(-0.2, 0.2)), float(rng.uniform(-0.1, 0.1))) for _ in range(int(rng.integers(1, 10)))]
    gpu_ctx.clear_trees()
    oracle.clear_trees()
    lanes = int(rng.choice([1, 4, 8])) if rng.random() < 0.5 else None
    if which == 2:
        mdp = m.ContinuousPendulum(noise=float(rng.choice([2.0, 10.0])), effort=float(rng.choice([0.0, 1e-4])))
        solver = m.DPWSolver(n_iterations=int(rng.integers(1, 250)), depth=int(rng.integers(1, 10)), rng=int(rng.integers(0, 2**31)),
                             exploration_constant=float(rng.choice([0.0, 2.0, 10.0])), k_action=float(rng.choice([2.0, 8.0])),
                             alpha_action=float(rng.choice([0.25, 0.5])), alpha_state=float(rng.choice([0.125, 0.5])),
                             check_repeat_action=bool(rng.random() < 0.7), check_repeat_state=bool(rng.random() < 0.7),
                             keep_tree=bool(rng.random() < 0.4), keep_tree_calls=3,
                             estimate_value=m.RolloutEstimator(m.PluginPolicy(100) if rng.random() < 0.5 else m.RandomSolver(), max_depth=int(rng.choice([-1, 6]))))
        for call in range(2 if solver.keep_tree else 1):
            _run_box_both(gpu_ctx, oracle, solver, mdp, roots, base=2, lanes=lanes)
    else:
        mdp = m.NoisyMountainCar() if which == 0 else m.InvertedPendulum()
        if which == 0:
            roots = [(float(rng.uniform(-1.1, 0.4)), float(rng.uniform(-0.05, 0.05))) for _ in roots]
        solver = _random_solver(rng, mdp, dense=False)
        if isinstance(solver, m.MCTSSolver):
            solver.enable_tree_vis = False  # (unbounded branching: _vis_stats needs max_branch >= 1)
        kept = getattr(solver, "keep_tree", False) or getattr(solver, "reuse_tree", False)
        for call in range(2 if kept else 1):
            run_both(gpu_ctx, oracle, solver, mdp, roots, base=3, lanes=lanes, literal=True)
    gpu_ctx.clear_trees()
    oracle.clear_trees()


@pytest.mark.parametrize("seed", range(20 * SCALE))
def test_fuzz_multi_device_ctx(emu_lib_path, seed):
    """mctsb200_ctx_create_multi: the same calls on a ctx over several "devices" (the emulator's one device, twice or three times)
    return what a single-device ctx returns -- plans, exported trees wherever they live, episodes, root-parallel sums."""
    rng = np.random.default_rng(13000 + seed)
    mdp = _random_gridworld(rng) if rng.random() < 0.7 else m.MountainCar()
    dense = isinstance(mdp, m.SimpleGridWorld)
    solver = _random_solver(rng, mdp, dense)
    for attr in ("keep_tree", "reuse_tree"):  # (kept trees pin tree i to its device: covered by the single-device cases)
        if hasattr(solver, attr):
            setattr(solver, attr, False)
    n = int(rng.integers(1, 30))
    roots = _random_roots(rng, mdp, n) if dense else [(float(rng.uniform(-1.0, 0.4)), float(rng.uniform(-0.05, 0.05))) for _ in range(n)]
    one = m.Context(0, lib_path=emu_lib_path)
    many = m.Context([0] * int(rng.integers(2, 4)), lib_path=emu_lib_path)
    cfg, keep = m.build_cfg(solver, mdp)
    r = mdp.encode_states(roots)
    base = int(rng.integers(0, 100))
    for c in (one, many):
        c.configure(mdp.mdp_id, mdp.params())
    a1, q1, s1, _ = one.plan(cfg, mdp.mdp_id, r, base, want_status=True)
    a2, q2, s2, _ = many.plan(cfg, mdp.mdp_id, r, base, want_status=True)
    assert np.array_equal(a1, a2) and np.array_equal(q1.view(np.int64), q2.view(np.int64)) and np.array_equal(s1, s2)
    for i in range(0, n, max(1, n // 4)):
        t1, t2 = one.tree_export(i, mdp.state_dtype, mdp.state_width), many.tree_export(i, mdp.state_dtype, mdp.state_width)
        assert all(np.array_equal(t1[k].view(np.uint8), t2[k].view(np.uint8)) for k in t1)
    if getattr(solver, "enable_tree_vis", False):
        solver.enable_tree_vis = False
        cfg, keep = m.build_cfg(solver, mdp)
    e1 = one.run_episodes(cfg, mdp.mdp_id, r, 3, 17, base)
    e2 = many.run_episodes(cfg, mdp.mdp_id, r, 3, 17, base)
    assert all(np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8)) for x, y in zip(e1[:3], e2[:3]))
    if isinstance(solver, m.MCTSSolver):
        p1 = one.plan_root_parallel(cfg, mdp.mdp_id, r[:1], 11, tree_index_base=base)
        p2 = many.plan_root_parallel(cfg, mdp.mdp_id, r[:1], 11, tree_index_base=base)
        assert np.array_equal(p1[0], p2[0]) and np.array_equal(p1[1].view(np.int64), p2[1].view(np.int64))
    one.close()
    many.close()
    del keep
