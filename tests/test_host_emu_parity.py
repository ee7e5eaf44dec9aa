"""CPU-only checks of the shipped library's sources:

* the host-emulation build of mcts.jl_b200/csrc (tests/emu: the same .cu files compiled by g++ as a
  single-threaded program, one lane per tree) against the CPU oracle, bit-for-bit -- this exercises the host logic
  (validation, table sizing, widening thresholds, lane ordering, export) and the kernel's control flow without
  a GPU;
* the real libmctsb200.so loads and exports every symbol include/mctsb200.h declares (no compute call).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mcts_jl_b200 as m
from helpers import gridworld_roots, mountaincar_roots, run_both, run_episodes_both

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GW = m.SimpleGridWorld()
GW_DET = m.SimpleGridWorld(tprob=1.0)
MC = m.MountainCar()


def test_shipped_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mctsb200.h")).read()
    declared = set(re.findall(r"^(?:int32_t|const char\*)\s+(mctsb200_[a-z_]+)\s*\(", hdr, flags=re.M))
    assert declared >= set(m.abi.EXPORTED_SYMBOLS) and len(declared) >= 15
    m.build_library()
    lib = C.CDLL(m._lib.DEFAULT_LIB)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"libmctsb200.so does not export {sym}"
    assert lib.mctsb200_abi_version() == m.abi.ABI_VERSION


def test_no_device_means_error_not_fallback():
    """Without a CUDA device ctx_create fails loudly (there is no CPU path in the shipped library)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(m.MCTSB200Error) as ei:
        m.Context(0)
    assert ei.value.status == m.abi.ERR_CUDA


def test_emu_uct_deterministic(emu_ctx, oracle):
    roots = gridworld_roots(100)
    for est in (m.RolloutEstimator(m.ConstantPolicy("up")), m.RolloutEstimator(m.SeekTarget((9, 3)))):
        s = m.MCTSSolver(n_iterations=200, depth=20, exploration_constant=5.0, estimate_value=est)
        run_both(emu_ctx, oracle, s, GW_DET, roots, compare_trees=range(0, 100, 3))


def test_emu_uct_stochastic(emu_ctx, oracle):
    s = m.MCTSSolver(n_iterations=250, depth=20, exploration_constant=5.0, rng=7)
    _, _, _, st = run_both(emu_ctx, oracle, s, GW, gridworld_roots(64), base=1000, compare_trees=range(0, 64, 5))
    assert st.kernel_variant == 1  # the shared-memory lane-per-tree kernel


@pytest.mark.parametrize("ahead,det", [("0", "1"), ("1", "0"), ("1", "1")])
def test_emu_uct_both_descent_orders(emu_ctx, oracle, monkeypatch, ahead, det):
    """successors looked up ahead of the selection (lockstep / small batches) or after it, with and without the
    no-random-word transitions of a deterministic MDP: same trees"""
    monkeypatch.setenv("MCTSB200_AHEAD", ahead)
    monkeypatch.setenv("MCTSB200_DET", det)
    s = m.MCTSSolver(n_iterations=150, depth=15, exploration_constant=5.0, rng=3)
    run_both(emu_ctx, oracle, s, GW, gridworld_roots(40), compare_trees=range(0, 40, 7))
    for pol in (m.SeekTarget((9, 3)), m.ConstantPolicy("left"), m.RandomSolver()):
        s = m.MCTSSolver(n_iterations=150, depth=15, exploration_constant=5.0, estimate_value=m.RolloutEstimator(pol))
        run_both(emu_ctx, oracle, s, GW_DET, gridworld_roots(40), compare_trees=range(0, 40, 7))


def test_emu_generic_uct_kernel_still_matches(emu_ctx, oracle, monkeypatch):
    """MCTSB200_FAST=0 forces the tile kernel on the configurations the lean kernel normally takes."""
    monkeypatch.setenv("MCTSB200_FAST", "0")
    s = m.MCTSSolver(n_iterations=200, depth=20, exploration_constant=5.0, rng=7)
    _, _, _, st = run_both(emu_ctx, oracle, s, GW, gridworld_roots(40), compare_trees=range(0, 40, 5))
    assert st.kernel_variant == 0
    s = m.MCTSSolver(n_iterations=150, depth=20, exploration_constant=5.0, estimate_value=m.RolloutEstimator(m.SeekTarget((9, 3))))
    run_both(emu_ctx, oracle, s, GW_DET, gridworld_roots(40), compare_trees=range(0, 40, 5))


@pytest.mark.parametrize("opts", [
    dict(exploration_constant=0.0),
    dict(init_N=3, init_Q=11.73, estimate_value=0.0),
    dict(estimate_value=m.RolloutEstimator(m.RandomSolver(), max_depth=-1)),
    dict(estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"), max_depth=-1)),
    dict(estimate_value=m.RolloutEstimator(m.RandomSolver(), eps=0.2)),
    dict(rollouts_per_leaf=5),
    dict(depth=0),
    dict(n_iterations=1),
])
def test_emu_uct_options(emu_ctx, oracle, opts):
    kw = dict(n_iterations=150, depth=8, exploration_constant=3.0, rng=5)
    kw.update(opts)
    run_both(emu_ctx, oracle, m.MCTSSolver(**kw), GW, gridworld_roots(30) + [(-1, -1)])


def test_emu_dpw_gridworld(emu_ctx, oracle):
    s = m.DPWSolver(n_iterations=300, depth=20, exploration_constant=10.0, k_state=4.0, alpha_state=1 / 8, k_action=4.0, alpha_action=1 / 8, rng=3)
    _, _, _, st = run_both(emu_ctx, oracle, s, GW, gridworld_roots(50), compare_trees=range(0, 50, 4))
    assert st.kernel_variant == 1  # the shared-memory lane-per-tree DPW kernel
    # deterministic model: every action node has one successor, self-loops at the walls repeat (node, slot) pairs on a path
    s = m.DPWSolver(n_iterations=300, depth=20, exploration_constant=5.0, rng=3, estimate_value=m.RolloutEstimator(m.SeekTarget((9, 3))))
    run_both(emu_ctx, oracle, s, GW_DET, gridworld_roots(50), compare_trees=range(0, 50, 4))
    # widening that actually stops: few successors allowed, so transitions are re-sampled from the stored multiset
    s = m.DPWSolver(n_iterations=400, depth=12, exploration_constant=3.0, k_state=1.0, alpha_state=0.2, k_action=1.5, alpha_action=0.3, rng=5)
    _, _, _, st = run_both(emu_ctx, oracle, s, GW, gridworld_roots(30), compare_trees=range(0, 30, 3))
    assert st.kernel_variant == 1 and st.n_transition_samples > 0


def test_emu_generic_dpw_kernel_still_matches(emu_ctx, oracle, monkeypatch):
    monkeypatch.setenv("MCTSB200_FAST", "0")
    s = m.DPWSolver(n_iterations=200, depth=15, exploration_constant=10.0, k_state=4.0, alpha_state=1 / 8, k_action=4.0, alpha_action=1 / 8, rng=3)
    _, _, _, st = run_both(emu_ctx, oracle, s, GW, gridworld_roots(30), compare_trees=range(0, 30, 4))
    assert st.kernel_variant == 0


@pytest.mark.parametrize("opts", [
    dict(),
    dict(enable_action_pw=False),
    dict(k_state=1.0, alpha_state=0.1),
    dict(check_repeat_state=False, k_state=2.0, alpha_state=0.3),
    dict(enable_state_pw=False),
    dict(init_N=2, init_Q=-1.0, estimate_value=4.0),
    dict(exploration_constant=0.0),
    dict(rollouts_per_leaf=3),
])
def test_emu_dpw_options(emu_ctx, oracle, opts):
    kw = dict(n_iterations=200, depth=10, rng=11)
    kw.update(opts)
    run_both(emu_ctx, oracle, m.DPWSolver(**kw), GW, gridworld_roots(24))


def test_emu_mountaincar(emu_ctx, oracle):
    roots = mountaincar_roots(8)
    run_both(emu_ctx, oracle, m.DPWSolver(n_iterations=300, depth=50, rng=8), MC, roots)
    run_both(emu_ctx, oracle, m.MCTSSolver(n_iterations=200, depth=30, exploration_constant=10.0, rng=9), MC, roots)


def test_emu_large_grid_uses_map_layout(emu_ctx, oracle):
    """A grid with far more cells than iterations (node table sized by n_iterations, not by the grid)."""
    big = m.SimpleGridWorld(size=(60, 50), rewards={(30, 20): 5.0, (7, 9): -3.0})
    roots = [(1 + (7 * i) % 60, 1 + (11 * i) % 50) for i in range(20)]
    run_both(emu_ctx, oracle, m.MCTSSolver(n_iterations=80, depth=12, exploration_constant=2.0, rng=4), big, roots)
    run_both(emu_ctx, oracle, m.DPWSolver(n_iterations=80, depth=12, rng=4), big, roots)


def test_emu_terminal_dpw_root_reports_status(emu_ctx, oracle):
    roots = gridworld_roots(10) + [(-1, -1)]
    a, q, st, stats = run_both(emu_ctx, oracle, m.DPWSolver(n_iterations=50, depth=10, rng=2), GW, roots)
    assert st[-1] == m.abi.ERR_TERMINAL_ROOT and (st[:-1] == 0).all()


# ---- edge cases of the lane-per-tree kernels (their own code paths: tables, tags, forwarding, ordering) ----------------
@pytest.mark.parametrize("kind", ["uct", "dpw"])
@pytest.mark.parametrize("case", [
    dict(mdp=dict(size=(11, 11), rewards={(6, 6): 2.0, (1, 11): -1.0}), roots=[(1, 1), (11, 11), (6, 5), (6, 6), (3, 9)]),  # 121 cells: largest fast grid
    dict(mdp=dict(size=(1, 7), rewards={(1, 7): 1.0}), roots=[(1, 1), (1, 4), (1, 7)]),                                       # corridor: most moves fold into "stay"
    dict(mdp=dict(size=(3, 2), rewards={(3, 2): -0.0, (1, 1): 4.0}, tprob=0.4), roots=[(2, 1), (3, 2), (1, 2)]),            # a reward of -0.0
    dict(mdp=dict(tprob=0.0), roots=[(1, 1), (5, 5), (10, 10)]),                                                               # the intended move never happens
    dict(mdp=dict(tprob=1.0, discount=1.0), roots=[(1, 1), (9, 2), (4, 4)]),
    dict(mdp=dict(rewards={(5, 5): 1.0}, terminate_from=[(2, 2)]), roots=[(5, 5), (2, 2), (5, 4)]),                           # reward cell that does not terminate
])
def test_emu_fast_kernels_edge_models(emu_ctx, oracle, kind, case):
    mdp = m.SimpleGridWorld(**case["mdp"])
    S = m.MCTSSolver if kind == "uct" else m.DPWSolver
    for kw in (dict(), dict(init_N=2, init_Q=0.5), dict(estimate_value=m.RolloutEstimator(m.RandomSolver(), max_depth=-1, eps=0.3)),
               dict(estimate_value=m.RolloutEstimator(m.ConstantPolicy("left"), max_depth=7))):
        s = S(n_iterations=120, depth=9, exploration_constant=2.5, rng=13, **kw)
        _, _, _, st = run_both(emu_ctx, oracle, s, mdp, case["roots"] * 3)
        assert st.kernel_variant == 1


@pytest.mark.parametrize("n_roots", [1, 31, 33, 97])
def test_emu_fast_kernels_ragged_batches_and_lane_order(emu_ctx, oracle, monkeypatch, n_roots):
    roots = gridworld_roots(n_roots)
    # lane order on / off; trees spread over idle-lane warps (1 = every lane holds a tree, default: chosen by batch size)
    for order, spread in (("0", "1"), ("1", "1"), ("0", "4"), ("0", "")):
        monkeypatch.setenv("MCTSB200_LANE_ORDER", order)
        monkeypatch.setenv("MCTSB200_SPREAD", spread)
        for s in (m.MCTSSolver(n_iterations=60, depth=10, exploration_constant=5.0, rng=2),
                  m.MCTSSolver(n_iterations=60, depth=10, exploration_constant=5.0, estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"))),
                  m.DPWSolver(n_iterations=60, depth=10, exploration_constant=5.0, rng=2)):
            mdp = GW_DET if isinstance(s.estimate_value.solver, m.ConstantPolicy) else GW
            _, _, _, st = run_both(emu_ctx, oracle, s, mdp, roots, base=7)
            assert st.kernel_variant == 1


def test_emu_grid_above_127_cells_uses_tile_kernel(emu_ctx, oracle):
    mdp = m.SimpleGridWorld(size=(12, 12), rewards={(6, 6): 1.0})
    _, _, _, st = run_both(emu_ctx, oracle, m.MCTSSolver(n_iterations=100, depth=10, exploration_constant=3.0, rng=1), mdp, [(1, 1), (12, 12)])
    assert st.kernel_variant == 0
    _, _, _, st = run_both(emu_ctx, oracle, m.DPWSolver(n_iterations=100, depth=10, rng=1), mdp, [(1, 1), (12, 12)])
    assert st.kernel_variant == 0


def test_emu_max_time_chunks_give_the_same_trees(emu_ctx, oracle):
    """max_time runs the search in several launches; with a generous budget all iterations run and nothing may differ."""
    for s in (m.MCTSSolver(n_iterations=90, depth=10, exploration_constant=5.0, rng=4, max_time=1e6),
              m.DPWSolver(n_iterations=90, depth=10, rng=4, max_time=1e6)):
        _, _, _, st = run_both(emu_ctx, oracle, s, GW, gridworld_roots(20))
        assert st.iterations_done == 90 and st.n_kernel_launches > 2


# ---- trees kept across calls: reuse_tree (src/vanilla.jl:213-217) / keep_tree (src/dpw.jl:31-37) -----------------------
@pytest.mark.parametrize("fast", ["1", "0"])
@pytest.mark.parametrize("kind", ["uct", "dpw"])
def test_emu_kept_trees_continue_across_calls(emu_ctx, oracle, monkeypatch, kind, fast):
    monkeypatch.setenv("MCTSB200_FAST", fast)
    emu_ctx.clear_trees()
    oracle.clear_trees()
    if kind == "uct":
        mk = lambda seed: m.MCTSSolver(n_iterations=70, depth=10, exploration_constant=5.0, rng=seed, reuse_tree=True)
    else:
        mk = lambda seed: m.DPWSolver(n_iterations=70, depth=10, exploration_constant=5.0, rng=seed, keep_tree=True)
    roots = gridworld_roots(40)
    sizes = []
    for call, shift in enumerate((0, 1, 11, 0)):  # an episode-like sequence of root states per tree; the last returns to the start
        rs = [(1 + (x - 1 + shift) % 10, 1 + (y - 1 + shift // 10) % 10) for x, y in roots]
        _, _, _, st = run_both(emu_ctx, oracle, mk(100 + call), GW, rs, compare_trees=range(0, 40, 3))
        assert st.kernel_variant == int(fast)
        sizes.append(emu_ctx.tree_export(0)["total_n"].shape[0])
    assert sizes == sorted(sizes) and sizes[-1] > sizes[0]  # the tree of root 0 only ever grows
    # clear_tree! starts over
    emu_ctx.clear_trees()
    oracle.clear_trees()
    run_both(emu_ctx, oracle, mk(100), GW, roots, compare_trees=range(0, 40, 7))
    assert emu_ctx.tree_export(0)["total_n"].shape[0] == sizes[0]


@pytest.mark.parametrize("kind", ["dpw", "dpw-unmerged", "uct"])
def test_emu_kept_trees_of_hashed_models(emu_ctx, oracle, kind):
    """keep_tree / reuse_tree on a model without a bounded state table (MountainCar: hashed continuous states; DPW's main use with
    kept trees, src/dpw.jl:31-37): the trees get room for keep_tree_calls calls; a 4-call episode-like sequence -- every root moves
    on to a state its own tree already holds, or to a new one -- equals the oracle's bit for bit."""
    emu_ctx.clear_trees()
    oracle.clear_trees()
    mc = m.MountainCar()
    if kind == "uct":
        mk = lambda seed: m.MCTSSolver(n_iterations=60, depth=12, exploration_constant=8.0, rng=seed, reuse_tree=True)
    else:
        mk = lambda seed: m.DPWSolver(n_iterations=60, depth=12, exploration_constant=8.0, rng=seed, keep_tree=True, check_repeat_state=kind == "dpw")
    roots = mountaincar_roots(12)
    for call in range(4):
        a, _, _, _ = run_both(emu_ctx, oracle, mk(300 + call), mc, roots, compare_trees=range(0, 12, 3))
        nxt = []
        for (x, v), ai in zip(roots, a):  # follow the chosen action: the successor is in the tree already (deterministic model)
            sp, _, _ = oracle.gen(mc.mdp_id, (x, v), int(ai))
            nxt.append((float(sp[0]), float(sp[1])) if sp[0] < 0.5 else (x, v))
        roots = nxt
    sizes = emu_ctx.tree_export(0, np.float64, 2)["total_n"].shape[0]
    assert sizes > 61  # the tree kept growing across the calls
    emu_ctx.clear_trees()
    oracle.clear_trees()


def test_emu_kept_unmerged_trees_keep_widening(emu_ctx, oracle):
    """Found by tests/test_fuzz_parity.py: with check_repeat_state = false every transition makes a node, so n_a_children of a kept
    tree's action node grows call after call -- past the length of a widening-threshold table sized for ONE call, where the search
    silently stopped widening. Three calls on a terminate_from root (every action ends the episode: all visits go to one action node
    per call), against the oracle; and a long unmerged search whose visit bound exceeds 2^20 still gets its table."""
    gw = m.SimpleGridWorld(size=(8, 5), rewards={(5, 4): 10.0, (1, 2): 3.0}, terminate_from={(5, 4), (1, 2)})
    s = m.DPWSolver(n_iterations=41, depth=15, exploration_constant=0.0, k_action=4.0, alpha_action=0.8, check_repeat_state=False, keep_tree=True,
                    keep_tree_calls=4, init_N=0, rng=7)
    emu_ctx.clear_trees()
    oracle.clear_trees()
    for _ in range(3):
        run_both(emu_ctx, oracle, s, gw, [(5, 4), (1, 2), (4, 4)])
    # one new (terminal) state node per iteration in every call; the second call also makes a second root node: the first call's root
    # was inserted with maintain_s_lookup = check_repeat_state = false (src/dpw.jl:40), so the kept tree cannot find it (src/dpw.jl:33)
    assert emu_ctx.tree_export(0)["total_n"].shape[0] == 42 + 42 + 41
    emu_ctx.clear_trees()
    oracle.clear_trees()
    big = m.DPWSolver(n_iterations=60000, depth=20, k_state=10.0, alpha_state=1.0, check_repeat_state=False, estimate_value=0.0, rng=1)
    cfg, keep = m.build_cfg(big, gw)
    emu_ctx.configure(gw.mdp_id, gw.params())
    emu_ctx.plan(cfg, gw.mdp_id, gw.encode_states([(5, 4)]))  # (60 000 one-level iterations)
    assert emu_ctx.tree_export(0)["total_n"].shape[0] == 60001


def test_emu_kept_trees_report_capacity_when_their_room_is_used_up(emu_ctx):
    mc = m.MountainCar()
    emu_ctx.clear_trees()
    cfg, keep = m.build_cfg(m.DPWSolver(n_iterations=40, depth=10, keep_tree=True, keep_tree_calls=1), mc)
    emu_ctx.configure(mc.mdp_id, mc.params())
    r = mc.encode_states(mountaincar_roots(3))
    emu_ctx.plan(cfg, mc.mdp_id, r)
    with pytest.raises(m.MCTSB200Error) as ei:
        for _ in range(3):
            emu_ctx.plan(cfg, mc.mdp_id, r)
    assert ei.value.status == m.abi.ERR_CAPACITY
    emu_ctx.clear_trees()
    emu_ctx.plan(cfg, mc.mdp_id, r)  # clear_tree! and start over


# ---- episode driver: simulate(RolloutSimulator(max_steps), mdp, planner, s0) (bench/non-terminal_gw.jl:18,31) ----------
@pytest.mark.parametrize("kind", ["uct", "uct-reuse", "dpw", "dpw-keep"])
def test_emu_episodes_gridworld(emu_ctx, oracle, kind):
    est = m.RolloutEstimator(m.RandomSolver(), max_depth=10)
    if kind.startswith("uct"):
        s = m.MCTSSolver(n_iterations=60, depth=10, exploration_constant=5.0, rng=3, estimate_value=est, reuse_tree=kind.endswith("reuse"))
    else:
        s = m.DPWSolver(n_iterations=60, depth=10, exploration_constant=5.0, rng=3, estimate_value=est, keep_tree=kind.endswith("keep"))
    starts = [p for p in gridworld_roots(100) if not GW.isterminal(p)][:48]
    ret, steps, final, st = run_episodes_both(emu_ctx, oracle, s, GW, starts, max_steps=12, env_seed=99, base=7)
    assert steps.max() == 12 and steps.min() < 12  # some episodes reach a reward cell early, some run to the step limit
    assert (ret != 0).any()
    assert st.n_simulations <= 60 * 48 * 12


def test_emu_episodes_nonterminal_gridworld(emu_ctx, oracle):
    """bench/non-terminal_gw.jl: no terminal states, every episode runs max_steps steps."""
    gw = m.SimpleGridWorld(terminate_from=())
    s = m.MCTSSolver(n_iterations=40, depth=8, exploration_constant=5.0, rng=1)
    ret, steps, final, st = run_episodes_both(emu_ctx, oracle, s, gw, gridworld_roots(33), max_steps=9, env_seed=5)
    assert (steps == 9).all()
    assert st.n_simulations == 40 * 33 * 9


def test_emu_episodes_mountaincar_and_terminal_starts(emu_ctx, oracle):
    s = m.DPWSolver(n_iterations=30, depth=8, exploration_constant=2.0, rng=11, estimate_value=m.RolloutEstimator(m.EnergyPolicy(), max_depth=20))
    starts = mountaincar_roots(6) + [(0.6, 0.0)]  # the last start state is already terminal: zero steps, zero return
    ret, steps, final, st = run_episodes_both(emu_ctx, oracle, s, MC, starts, max_steps=5, env_seed=2)
    assert steps[-1] == 0 and ret[-1] == 0.0 and tuple(final[-1]) == (0.6, 0.0)
    ret0, steps0, _, _ = run_episodes_both(emu_ctx, oracle, s, MC, starts, max_steps=0)
    assert (steps0 == 0).all() and (ret0 == 0).all()


@pytest.mark.parametrize("kind", ["dpw-keep", "uct-reuse"])
def test_emu_episodes_mountaincar_with_kept_trees(emu_ctx, oracle, kind):
    """the closed-loop driver with trees that follow their episodes, on a hashed-state model (see test_emu_kept_trees_of_hashed_models)"""
    est = m.RolloutEstimator(m.RandomSolver(), max_depth=15)
    if kind == "dpw-keep":
        s = m.DPWSolver(n_iterations=40, depth=8, exploration_constant=8.0, rng=11, estimate_value=est, keep_tree=True)
    else:
        s = m.MCTSSolver(n_iterations=40, depth=8, exploration_constant=8.0, rng=11, estimate_value=est, reuse_tree=True)
    ret, steps, final, st = run_episodes_both(emu_ctx, oracle, s, MC, mountaincar_roots(6), max_steps=5, env_seed=2)
    assert steps.max() == 5 and 6 * 40 * 2 < st.n_simulations <= 6 * 40 * 5


def test_emu_simulate_mirror(emu_ctx):
    planner = m.solve(m.MCTSSolver(n_iterations=30, depth=8, rng=4), GW, ctx=emu_ctx)
    sim = m.RolloutSimulator(rng=8, max_steps=6)
    one = m.simulate(sim, GW, planner, (2, 2))
    many = m.simulate_batch(sim, GW, m.solve(m.MCTSSolver(n_iterations=30, depth=8, rng=4), GW, ctx=emu_ctx), [(2, 2), (5, 5)])
    assert isinstance(one, float) and many.shape == (2,) and one == many[0]
    with pytest.raises(m.MCTSB200Error):
        m.simulate(m.RolloutSimulator(), GW, planner, (2, 2))


def test_emu_visit_counts_beyond_the_ucb_tables(emu_ctx, oracle, monkeypatch):
    """a long max_time search outgrows the sqrt(log N) / 1/sqrt(n) tables: those selections take the exact expression"""
    monkeypatch.setenv("MCTSB200_UCB_TABLE_ENTRIES", "2048")
    for s in (m.MCTSSolver(n_iterations=4000, depth=6, exploration_constant=5.0, rng=5),
              m.DPWSolver(n_iterations=4000, depth=6, exploration_constant=5.0, rng=5)):
        for fast in ("1", "0"):
            monkeypatch.setenv("MCTSB200_FAST", fast)
            _, _, _, st = run_both(emu_ctx, oracle, s, GW, [(2, 2), (9, 9)])
            assert st.kernel_variant == int(fast) and st.n_select_exact > 1000


def test_environment_cannot_select_the_emulation_build(emu_lib_path, monkeypatch):
    """MCTSB200_LIB is for alternative CUDA builds; pointing it at the tests' CPU emulation is refused (no CPU path one env var away)."""
    monkeypatch.setenv("MCTSB200_LIB", emu_lib_path)
    with pytest.raises(m.MCTSB200Error) as ei:
        m.load_library()
    assert ei.value.status == m.abi.ERR_CUDA and "emulation" in str(ei.value)
    assert m.load_library(emu_lib_path) is not None  # a test that names it explicitly may


# ---- continuous action spaces (test/discussion_514.jl: actions(m) = Box) ----------------------------------------------------
def _run_box_both(ctx, orc, solver, mdp, roots, base=0, lanes=None, trees=None):
    """DPW over a continuous action space on the library and on the oracle: chosen action vectors, Q, counters and whole trees
    (children lists, vector labels, transitions) bit for bit."""
    from helpers import COUNTER_KEYS, assert_trees_equal, bits

    cfg, keep = m.build_cfg(solver, mdp)
    if lanes is not None:
        cfg.lanes_per_tree = lanes
    oid = getattr(mdp, "oracle_mdp_id", None) or mdp.mdp_id  # (run-time plugins: the id under which the oracle restates the model)
    ctx.configure(mdp.mdp_id, mdp.params())
    orc.configure(oid, mdp.params())
    r = mdp.encode_states(roots)
    a1, q1, s1, st1 = ctx.plan(cfg, mdp.mdp_id, r, base, want_status=True)
    a2, q2, s2, st2 = orc.plan(cfg, oid, r, base, n_threads=4)
    assert np.array_equal(s1, s2) and np.array_equal(a1, a2), (a1, a2)
    assert np.array_equal(bits(q1), bits(q2))
    v1, v2 = ctx.root_action_vectors(0, len(roots), mdp.action_width), orc.root_action_vectors(0, len(roots), mdp.action_width)
    assert np.array_equal(bits(v1), bits(v2))
    d1, d2 = st1.as_dict(), st2.as_dict()
    for k in COUNTER_KEYS:
        if k != "algorithmic_bytes":
            assert d1[k] == d2[k], f"counter {k}: {d1[k]} != {d2[k]}"
    for i in (trees if trees is not None else range(len(roots))):
        t1 = ctx.tree_export(i, mdp.state_dtype, mdp.state_width)
        t2 = orc.tree_export(i, mdp.state_dtype, mdp.state_width)
        assert "a_label_vectors" in t1 and np.all(t1["a_labels"] == -1)
        assert_trees_equal(t1, t2, f"tree {i}")
        rs1, rs2 = ctx.root_stats(i, 4096), orc.root_stats(i, 4096)
        for k in ("action_idx", "n", "q"):
            assert np.array_equal(bits(rs1[k]), bits(rs2[k]))
    del keep
    return v1, q1, st1


BOX_ROOTS = [(0.0, 0.0, 0.0), (1.5, -2.0, 0.5), (-3.0, 3.0, -1.0), (4.0, 4.0, 2.0), (0.25, -0.125, 3.0)]


def test_emu_box_actions_discussion_514(emu_ctx, oracle):
    """The reference's regression test for continuous action spaces (test/discussion_514.jl): the planner returns an action inside
    actions(m) = Box([-5,-5],[5,5]); the trees equal the oracle's."""
    mdp = m.GaussianBox()  # the test's model: no drift, zero reward
    v, _, st = _run_box_both(emu_ctx, oracle, m.DPWSolver(n_iterations=100, depth=10, rng=3), mdp, BOX_ROOTS)
    assert all(mdp.action_in_space(a) for a in v)
    assert st.n_action_inserts > 100 and st.n_simulations == 500


@pytest.mark.parametrize("opts", [
    dict(),
    dict(exploration_constant=5.0, k_action=4.0, alpha_action=0.25, k_state=2.0, alpha_state=0.3),
    dict(check_repeat_state=False), dict(check_repeat_action=False), dict(enable_state_pw=False),
    dict(init_Q=-3.0, init_N=2), dict(estimate_value=-1.5), dict(exploration_constant=0.0),
    dict(estimate_value=m.RolloutEstimator(m.RandomSolver(), max_depth=-1)), dict(estimate_value=m.RolloutEstimator(m.RandomSolver(), max_depth=7, eps=0.3)),
])
def test_emu_box_actions_option_matrix(emu_ctx, oracle, opts):
    """A model where the search has something to find (drift towards the origin is rewarded): every DPW option."""
    mdp = m.GaussianBox(drift=0.2, cost=1.0, var=(0.05, 0.1, 0.02))
    kw = dict(n_iterations=150, depth=8, rng=11)
    kw.update(opts)
    v, q, _ = _run_box_both(emu_ctx, oracle, m.DPWSolver(**kw), mdp, BOX_ROOTS, base=77)
    assert all(mdp.action_in_space(a) for a in v)


@pytest.mark.parametrize("fast", ["1", "0"])
def test_emu_next_action_as_a_state_only_function(emu_ctx, oracle, monkeypatch, fast):
    """DPWSolver(next_action=(mdp, s, snode)->:up) (test/options.jl:50; src/action_gen.jl:14): tabulated over the dense state space, the
    action widening proposes table[s] and draws no random number. Against the oracle."""
    monkeypatch.setenv("MCTSB200_FAST", fast)
    roots = gridworld_roots(24)
    gens = (lambda mdp, s, snode: "up",
            lambda mdp, s, snode: ("right" if s[0] < 9 else "up") if (s[0] + s[1]) % 3 else "left")  # (the terminal state is (-1, -1))
    for gen in gens:
        for kw in (dict(), dict(k_action=2.0, alpha_action=0.3, init_Q=1.5, init_N=2), dict(estimate_value=0.0, enable_state_pw=False)):
            s = m.DPWSolver(n_iterations=80, depth=8, exploration_constant=4.0, rng=12, next_action=gen, **kw)
            _, _, _, st = run_both(emu_ctx, oracle, s, GW, roots, compare_trees=range(0, 24, 5))
            assert st.kernel_variant == 0  # (the tile kernel takes these calls whatever MCTSB200_FAST says)
    t = emu_ctx.tree_export(0)
    assert (np.diff(t["child_offsets"]) <= 1).all()  # one proposal per state: at most one child


def _box_kept_trees(ctx, orc, mdp, roots, lanes=None):
    """keep_tree over a continuous action space (src/dpw.jl:31-37): a 4-call sequence in which every root moves on to a successor its
    own tree already holds (the second state node: the first successor the search generated), trees compared after every call."""
    ctx.clear_trees()
    orc.clear_trees()
    sizes = []
    for call in range(4):
        s = m.DPWSolver(n_iterations=80, depth=8, exploration_constant=3.0, rng=40 + call, keep_tree=True, keep_tree_calls=4)
        _run_box_both(ctx, orc, s, mdp, roots, base=10, lanes=lanes)
        trees = [ctx.tree_export(i, mdp.state_dtype, mdp.state_width) for i in range(len(roots))]
        sizes.append(trees[0]["total_n"].shape[0])
        roots = [mdp.decode_state(t["s_labels"][1]) for t in trees]
    assert sizes == sorted(sizes) and sizes[-1] > sizes[0] + 100  # the kept tree grows across the calls
    cfg, keep = m.build_cfg(m.DPWSolver(n_iterations=80, depth=8, rng=1, keep_tree=True, keep_tree_calls=4), mdp)
    with pytest.raises(m.MCTSB200Error) as ei:  # ... until its room is used up
        for _ in range(8):
            ctx.plan(cfg, mdp.mdp_id, mdp.encode_states(roots))
    assert ei.value.status == m.abi.ERR_CAPACITY
    ctx.clear_trees()
    orc.clear_trees()


def test_emu_box_actions_kept_trees(emu_ctx, oracle):
    _box_kept_trees(emu_ctx, oracle, m.GaussianBox(drift=0.2, cost=1.0, var=(0.05, 0.1, 0.02)), BOX_ROOTS)


@pytest.mark.parametrize("keep", [False, True])
def test_emu_box_actions_episodes(emu_ctx, oracle, keep):
    """Closed-loop episodes over a continuous action space: the step's decision is the action VECTOR of best_sanode, the environment
    steps with it (gen(mdp, s, a, rng)); returns, final states and counters against the oracle."""
    mdp = m.GaussianBox(drift=0.3, cost=0.5, var=(0.05, 0.1, 0.02))
    s = m.DPWSolver(n_iterations=60, depth=6, exploration_constant=3.0, rng=8, keep_tree=keep, keep_tree_calls=6)
    ret, steps, final, st = run_episodes_both(emu_ctx, oracle, s, mdp, BOX_ROOTS * 2, max_steps=5, env_seed=31, base=3)
    assert (steps == 5).all() and (ret < 0).all()
    assert st.n_simulations == 60 * 10 * 5


def test_emu_box_root_stats_state_their_capacity(emu_ctx):
    """A root over a continuous action space has as many children as the search gave it: mctsb200_root_stats ("buffers hold n_actions
    entries") only answers the count there, mctsb200_root_stats_n writes at most `capacity` entries."""
    mdp = m.GaussianBox()
    cfg, keep = m.build_cfg(m.DPWSolver(n_iterations=300, depth=4, rng=1), mdp)
    emu_ctx.configure(mdp.mdp_id, mdp.params())
    emu_ctx.plan(cfg, mdp.mdp_id, mdp.encode_states(BOX_ROOTS[:2]))
    lib, h = emu_ctx.lib, emu_ctx._h
    nch, tot = C.c_int32(), C.c_int64()
    assert lib.mctsb200_root_stats(h, 1, C.byref(nch), None, None, None, C.byref(tot)) == m.abi.OK and nch.value > 16
    q = np.full(nch.value + 8, -7.0)
    assert lib.mctsb200_root_stats(h, 1, C.byref(nch), None, None, q.ctypes.data_as(C.c_void_p), C.byref(tot)) == m.abi.ERR_UNSUPPORTED
    assert lib.mctsb200_root_stats_n(h, 1, 5, C.byref(nch), None, None, q.ctypes.data_as(C.c_void_p), C.byref(tot)) == m.abi.OK
    assert (q[5:] == -7.0).all() and (q[:5] != -7.0).all()
    full = emu_ctx.root_stats(1)
    assert len(full["q"]) == nch.value and np.array_equal(full["q"][:5], q[:5]) and len(emu_ctx.root_stats(1, 3)["n"]) == 3
    # a finite action set through the same entry point
    gw = m.SimpleGridWorld()
    cfg2, keep2 = m.build_cfg(m.MCTSSolver(n_iterations=30, depth=5, rng=1), gw)
    emu_ctx.configure(gw.mdp_id, gw.params())
    emu_ctx.plan(cfg2, gw.mdp_id, gw.encode_states([(2, 2)]))
    assert len(emu_ctx.root_stats(0)["q"]) == 4 and len(emu_ctx.root_stats(0, 2)["q"]) == 2


def test_emu_box_actions_refuses_what_cannot_work(emu_ctx):
    mdp = m.GaussianBox()
    emu_ctx.configure(mdp.mdp_id, mdp.params())
    r = mdp.encode_states(BOX_ROOTS)
    for solver in (m.MCTSSolver(n_iterations=10, depth=5), m.DPWSolver(n_iterations=10, depth=5, enable_action_pw=False)):
        cfg, keep = m.build_cfg(solver, mdp)
        with pytest.raises(m.MCTSB200Error) as ei:
            emu_ctx.plan(cfg, mdp.mdp_id, r)
        assert ei.value.status == m.abi.ERR_UNSUPPORTED


@pytest.mark.parametrize("fast", ["1", "0"])
def test_emu_function_policy_is_a_table_on_the_device(emu_ctx, oracle, monkeypatch, fast):
    """RolloutEstimator(FunctionPolicy(f)) / RolloutEstimator(x -> :up) (test/options.jl:29, src/domain_knowledge.jl:57): a state-only
    host function is tabulated over the dense state space once (MCTSB200_POLICY_TABLE); lane-per-tree and tile kernels, UCT and DPW."""
    monkeypatch.setenv("MCTSB200_FAST", fast)
    f = lambda s: "up" if s[0] < 5 else ("left" if s[1] > 3 else "right")
    for est in (m.RolloutEstimator(m.FunctionPolicy(f), max_depth=30), m.RolloutEstimator(lambda s: "up")):
        run_both(emu_ctx, oracle, m.MCTSSolver(n_iterations=120, depth=12, exploration_constant=5.0, rng=5, estimate_value=est), GW, gridworld_roots(30),
                 compare_trees=range(0, 30, 7))
        run_both(emu_ctx, oracle, m.DPWSolver(n_iterations=120, depth=12, exploration_constant=5.0, rng=5, estimate_value=est), GW, gridworld_roots(30),
                 compare_trees=range(0, 30, 7))
    # x -> :up on the deterministic GridWorld is the bit-exact variant of BASELINE config 2, requested the reference's way
    est = m.RolloutEstimator(lambda s: "up")
    a1, q1, _, _ = run_both(emu_ctx, oracle, m.MCTSSolver(n_iterations=100, depth=20, exploration_constant=5.0, estimate_value=est), GW_DET, gridworld_roots(20))
    a2, q2, _, _ = run_both(emu_ctx, oracle, m.MCTSSolver(n_iterations=100, depth=20, exploration_constant=5.0, estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"))),
                            GW_DET, gridworld_roots(20))
    assert np.array_equal(a1, a2) and np.array_equal(q1.view(np.int64), q2.view(np.int64))
    with pytest.raises(m.MCTSB200Error):  # a hashed-state model has no table to put it in
        m.build_cfg(m.MCTSSolver(estimate_value=m.RolloutEstimator(lambda s: 1.0)), MC)


@pytest.mark.parametrize("mdp_name", ["gridworld", "mountaincar"])
def test_emu_vis_stats_like_record_visit(emu_ctx, oracle, mdp_name):
    """MCTSSolver(enable_tree_vis=true): tree._vis_stats[said => spid] counts every (action node, next state node) edge the search
    walked (record_visit!, src/vanilla.jl:268-270, 328-334) -- what D3Tree(::MCTSTree) renders. Against the oracle, entry for entry;
    the counts of an action node's edges add up to its visits."""
    mdp, roots = (GW, gridworld_roots(12)) if mdp_name == "gridworld" else (MC, mountaincar_roots(6))
    s = m.MCTSSolver(n_iterations=150, depth=12, exploration_constant=5.0, rng=9, enable_tree_vis=True)
    _, _, _, st = run_both(emu_ctx, oracle, s, mdp, roots, compare_trees=range(0, len(roots), 3))
    assert st.kernel_variant == 0  # the tile kernel records the edges
    for i in range(0, len(roots), 3):
        v1, v2 = emu_ctx.tree_vis_stats(i), oracle.tree_vis_stats(i)
        assert v1 == v2 and len(v1) > 10
        t = emu_ctx.tree_export(i, mdp.state_dtype, mdp.state_width)
        per_action = {}
        for (said, _spid), c in v1.items():
            per_action[said] = per_action.get(said, 0) + c
        for said, c in per_action.items():
            assert c == t["n"][said - 1]  # init_N = 0: every visit of the action node walked one of its edges
    # through the mirror: info[:tree]._vis_stats; without enable_tree_vis there is nothing to ask for
    planner = m.solve(s, mdp, ctx=emu_ctx)
    _, info = planner.action_info(roots[0])
    assert sum(info["tree"]._vis_stats.values()) >= 150
    planner2 = m.solve(m.MCTSSolver(n_iterations=20, depth=5), mdp, ctx=emu_ctx)
    _, info2 = planner2.action_info(roots[0])
    with pytest.raises(m.MCTSB200Error):
        info2["tree"]._vis_stats
