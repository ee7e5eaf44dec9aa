"""GPU parity proper: the shipped CUDA library, called through the C ABI, against the CPU oracle on the
same seeded inputs -- actions, best Q, status, work counters and every field of every exported tree
bit-for-bit, for every tile width."""
import numpy as np
import pytest

import mcts_jl_b200 as m
from helpers import run_episodes_both, gridworld_roots, mountaincar_roots, run_both

pytestmark = pytest.mark.gpu

GW = m.SimpleGridWorld()
GW_DET = m.SimpleGridWorld(tprob=1.0)
MC = m.MountainCar()
ALL_LANES = [1, 2, 4, 8, 16, 32]


def test_device_math_matches_host_bit_for_bit(gpu_ctx, oracle):
    n = np.arange(0, 1 << 20, dtype=np.float64)
    host = np.array([oracle.lib.oracle_log(float(x)) for x in n[:50000]])
    dev = gpu_ctx.device_log(n)
    assert np.array_equal(dev[:50000].view(np.int64), host.view(np.int64))
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-3.7, 1.8, 50000), rng.uniform(-1e5, 1e5, 20000), [0.0, -0.0, 0.5235987755982988]])
    hostc = np.array([oracle.lib.oracle_cos(float(v)) for v in x])
    assert np.array_equal(gpu_ctx.device_cos(x).view(np.int64), hostc.view(np.int64))
    xl = np.exp(rng.uniform(-700, 700, 50000))
    hostl = np.array([oracle.lib.oracle_log(float(v)) for v in xl])
    assert np.array_equal(gpu_ctx.device_log(xl).view(np.int64), hostl.view(np.int64))


def test_device_philox_matches_oracle(gpu_ctx, oracle):
    for seed, tree, it, sid in [(0, 0, 0, 0), (12345678901234567, 65535, 999, 1), (2**64 - 1, 2**40 + 17, 2**32 - 1, 33)]:
        assert np.array_equal(gpu_ctx.device_philox(seed, tree, it, sid, 8), oracle.philox(seed, tree, it, sid, 8))


@pytest.mark.parametrize("lanes", ALL_LANES)
def test_uct_gridworld_deterministic_bit_exact(gpu_ctx, oracle, lanes):
    """Config 2, deterministic variant (tprob=1, FunctionPolicy(s->:up) / SeekTarget rollouts), reduced size."""
    roots = gridworld_roots(200)
    for est in (m.RolloutEstimator(m.ConstantPolicy("up")), m.RolloutEstimator(m.SeekTarget((9, 3)))):
        s = m.MCTSSolver(n_iterations=300, depth=20, exploration_constant=5.0, estimate_value=est)
        run_both(gpu_ctx, oracle, s, GW_DET, roots, lanes=lanes, compare_trees=range(0, 200, 7))


@pytest.mark.parametrize("lanes", ALL_LANES)
def test_uct_gridworld_stochastic_bit_exact(gpu_ctx, oracle, lanes):
    s = m.MCTSSolver(n_iterations=250, depth=20, exploration_constant=5.0, rng=7)
    run_both(gpu_ctx, oracle, s, GW, gridworld_roots(128), base=1000, lanes=lanes, compare_trees=range(0, 128, 5))


@pytest.mark.parametrize("lanes", ALL_LANES)
def test_dpw_gridworld_bit_exact(gpu_ctx, oracle, lanes):
    """Config 3 solver (bench/non-terminal_gw_dpw.jl), reduced size."""
    s = m.DPWSolver(n_iterations=300, depth=20, exploration_constant=10.0, k_state=4.0, alpha_state=1 / 8, k_action=4.0, alpha_action=1 / 8, rng=3)
    run_both(gpu_ctx, oracle, s, GW, gridworld_roots(100), lanes=lanes, compare_trees=range(0, 100, 9))


@pytest.mark.parametrize("lanes", [1, 4, 32])
@pytest.mark.parametrize("opts", [
    dict(),
    dict(enable_action_pw=False),
    dict(k_state=1.0, alpha_state=0.1),
    dict(check_repeat_state=False, k_state=2.0, alpha_state=0.3),
    dict(enable_state_pw=False),
    dict(init_N=2, init_Q=-1.0, estimate_value=4.0),
    dict(exploration_constant=0.0),
])
def test_dpw_option_matrix_bit_exact(gpu_ctx, oracle, lanes, opts):
    kw = dict(n_iterations=200, depth=10, rng=11)
    kw.update(opts)
    run_both(gpu_ctx, oracle, m.DPWSolver(**kw), GW, gridworld_roots(40), lanes=lanes)


@pytest.mark.parametrize("lanes", [1, 4, 32])
@pytest.mark.parametrize("opts", [
    dict(exploration_constant=0.0),
    dict(init_N=3, init_Q=11.73, estimate_value=0.0),
    dict(estimate_value=m.RolloutEstimator(m.RandomSolver(), max_depth=-1)),
    dict(estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"), max_depth=-1)),
    dict(estimate_value=m.RolloutEstimator(m.RandomSolver(), eps=0.2)),
    dict(rollouts_per_leaf=8),
    dict(rollouts_per_leaf=40),
])
def test_uct_option_matrix_bit_exact(gpu_ctx, oracle, lanes, opts):
    kw = dict(n_iterations=150, depth=8, exploration_constant=3.0, rng=5)
    kw.update(opts)
    run_both(gpu_ctx, oracle, m.MCTSSolver(**kw), GW, gridworld_roots(40), lanes=lanes)


@pytest.mark.parametrize("lanes", [1, 4, 32])
def test_mountaincar_bit_exact(gpu_ctx, oracle, lanes):
    """Config 4 solver (defaults, depth 50), reduced size; plus vanilla UCT and a heuristic rollout policy."""
    roots = mountaincar_roots(24)
    run_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=400, depth=50, rng=8), MC, roots, lanes=lanes)
    run_both(gpu_ctx, oracle, m.MCTSSolver(n_iterations=300, depth=30, exploration_constant=10.0, rng=9), MC, roots, lanes=lanes)
    run_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=300, depth=50, rng=8, estimate_value=m.RolloutEstimator(m.EnergyPolicy(), max_depth=100)),
             MC, roots, lanes=lanes)


def test_auto_lanes_and_terminal_roots(gpu_ctx, oracle):
    roots = gridworld_roots(30) + [(-1, -1)]
    a, q, st, stats = run_both(gpu_ctx, oracle, m.MCTSSolver(n_iterations=50, depth=20, exploration_constant=5.0, rng=2), GW, roots)
    assert stats.lanes_per_tree in (1, 2, 4, 8, 16, 32)
    a, q, st, stats = run_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=50, depth=10, rng=2), GW, roots)
    assert st[-1] == m.abi.ERR_TERMINAL_ROOT and (st[:-1] == 0).all()


def test_full_size_config2_properties_and_sampled_oracle(gpu_ctx, oracle):
    """BASELINE config 2 at full size (65,536 roots x 1,000 iterations, deterministic variant): size-independent
    properties of the whole batch plus bit-exact trees for a sample of roots against the oracle."""
    roots = gridworld_roots(65536)
    solver = m.MCTSSolver(n_iterations=1000, depth=20, exploration_constant=5.0, rng=0,
                          estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"), max_depth=50))
    cfg, keep = m.build_cfg(solver, GW_DET)
    gpu_ctx.configure(GW_DET.mdp_id, GW_DET.params())
    enc = GW_DET.encode_states(roots)
    a, q, st, stats = gpu_ctx.plan(cfg, GW_DET.mdp_id, enc, want_status=True)
    assert stats.kernel_variant == 1 and stats.n_simulations == 65536 * 1000 and (st == 0).all()
    # deterministic model + deterministic rollout policy: the 655 trees that share a root cell are identical
    cell = np.arange(65536) % 100
    for c in range(100):
        idx = np.nonzero(cell == c)[0]
        assert (a[idx] == a[idx[0]]).all() and (q[idx].view(np.int64) == q[idx[0]].view(np.int64)).all()
    # every simulation backs up through the root at least once unless the root is terminal-adjacent: visit counts add up
    for i in (0, 1234, 65535):
        rs = gpu_ctx.root_stats(i)
        assert int(rs["n"].sum()) == rs["total_n"] >= 1000
    # sampled trees against the oracle, bit for bit (same global root indices)
    sample = [0, 1, 37, 99, 100, 4242, 65535]
    oracle.configure(GW_DET.mdp_id, GW_DET.params())
    for i in sample:
        a2, q2, _, _ = oracle.plan(cfg, GW_DET.mdp_id, enc[i:i + 1], i)
        assert a2[0] == a[i] and q2.view(np.int64)[0] == q.view(np.int64)[i]
        t1 = gpu_ctx.tree_export(i, GW_DET.state_dtype, GW_DET.state_width)
        t2 = oracle.tree_export(0, GW_DET.state_dtype, GW_DET.state_width)
        from helpers import assert_trees_equal
        assert_trees_equal(t1, t2, f"tree {i}")
    del keep


def test_full_size_config3_dpw_sampled_oracle(gpu_ctx, oracle):
    """BASELINE config 3 at full size (16,384 roots, bench/non-terminal_gw_dpw.jl solver): counters and sampled trees."""
    roots = gridworld_roots(16384)
    solver = m.DPWSolver(n_iterations=1000, depth=20, exploration_constant=10.0, k_state=4.0, alpha_state=1 / 8, k_action=4.0, alpha_action=1 / 8, rng=0)
    cfg, keep = m.build_cfg(solver, GW)
    gpu_ctx.configure(GW.mdp_id, GW.params())
    enc = GW.encode_states(roots)
    a, q, st, stats = gpu_ctx.plan(cfg, GW.mdp_id, enc, want_status=True)
    assert stats.kernel_variant == 1 and stats.n_simulations == 16384 * 1000 and (st == 0).all()
    oracle.configure(GW.mdp_id, GW.params())
    from helpers import assert_trees_equal
    for i in (0, 55, 9999, 16383):
        a2, q2, _, _ = oracle.plan(cfg, GW.mdp_id, enc[i:i + 1], i)
        assert a2[0] == a[i] and q2.view(np.int64)[0] == q.view(np.int64)[i]
        assert_trees_equal(gpu_ctx.tree_export(i, GW.state_dtype, GW.state_width), oracle.tree_export(0, GW.state_dtype, GW.state_width), f"tree {i}")
    del keep


def test_full_size_config4_mountaincar_sampled_oracle(gpu_ctx, oracle):
    """BASELINE config 4 at the size one GPU plans (1,024 of the 8,192 MountainCar roots, DPWSolver(n_iterations=10_000, depth=50),
    random rollouts): size-independent properties of every tree and whole sampled trees against the oracle, bit-for-bit. Roots
    are indexed as rank 3 of the 8-GPU run would index them (global root index = Philox stream)."""
    n, base = 1024, 3 * 1024
    roots = mountaincar_roots(8192)[base:base + n]
    solver = m.DPWSolver(n_iterations=10000, depth=50, rng=0)
    cfg, keep = m.build_cfg(solver, MC)
    gpu_ctx.configure(MC.mdp_id, MC.params())
    enc = MC.encode_states(roots)
    a, q, st, stats = gpu_ctx.plan(cfg, MC.mdp_id, enc, base, want_status=True)
    assert stats.kernel_variant == 0 and stats.n_simulations == n * 10000 and (st == 0).all()
    assert stats.n_rollouts >= stats.n_state_inserts - n  # a rollout per new state node (every node but the roots), and per depth-0 leaf
    for i in range(0, n, 37):
        rs = gpu_ctx.root_stats(i)
        assert rs["total_n"] == 10000 == int(rs["n"].sum()) and len(rs["n"]) == 3  # every iteration backs up through the root
        assert q[i] == rs["q"].max() and a[i] == int(rs["action_idx"][int(np.argmax(rs["q"]))])
    oracle.configure(MC.mdp_id, MC.params())
    from helpers import assert_trees_equal
    for i in (0, 511, 1023):
        a2, q2, _, _ = oracle.plan(cfg, MC.mdp_id, enc[i:i + 1], base + i)
        assert a2[0] == a[i] and q2.view(np.int64)[0] == q.view(np.int64)[i]
        assert_trees_equal(gpu_ctx.tree_export(i, MC.state_dtype, MC.state_width), oracle.tree_export(0, MC.state_dtype, MC.state_width), f"tree {i}")
    del keep


@pytest.mark.parametrize("kind,lanes", [("uct", None), ("dpw", None), ("uct", 4), ("dpw", 8)])
def test_kept_trees_continue_across_calls(gpu_ctx, oracle, kind, lanes):
    """reuse_tree (src/vanilla.jl:213-217) / keep_tree (src/dpw.jl:31-37): an episode-like sequence of plan calls grows the
    same trees; lanes=None runs the lane-per-tree kernels, otherwise the tile kernel."""
    gpu_ctx.clear_trees()
    oracle.clear_trees()
    if kind == "uct":
        mk = lambda seed: m.MCTSSolver(n_iterations=120, depth=12, exploration_constant=5.0, rng=seed, reuse_tree=True)
    else:
        mk = lambda seed: m.DPWSolver(n_iterations=120, depth=12, exploration_constant=5.0, rng=seed, keep_tree=True)
    roots = gridworld_roots(96)
    for call, shift in enumerate((0, 1, 11, 0)):
        rs = [(1 + (x - 1 + shift) % 10, 1 + (y - 1 + shift // 10) % 10) for x, y in roots]
        run_both(gpu_ctx, oracle, mk(200 + call), GW, rs, lanes=lanes, compare_trees=range(0, 96, 5))
    gpu_ctx.clear_trees()
    oracle.clear_trees()


# ---- episode driver: simulate(RolloutSimulator(max_steps), mdp, planner, s0) (bench/non-terminal_gw.jl:18,31) ----------
@pytest.mark.parametrize("kind", ["uct", "uct-reuse", "dpw", "dpw-keep"])
def test_episodes_gridworld(gpu_ctx, oracle, kind):
    est = m.RolloutEstimator(m.RandomSolver(), max_depth=10)
    if kind.startswith("uct"):
        s = m.MCTSSolver(n_iterations=150, depth=12, exploration_constant=5.0, rng=3, estimate_value=est, reuse_tree=kind.endswith("reuse"))
    else:
        s = m.DPWSolver(n_iterations=150, depth=12, exploration_constant=5.0, rng=3, estimate_value=est, keep_tree=kind.endswith("keep"))
    starts = [p for p in gridworld_roots(400) if not GW.isterminal(p)][:300]
    ret, steps, final, st = run_episodes_both(gpu_ctx, oracle, s, GW, starts, max_steps=15, env_seed=99, base=7)
    assert steps.max() == 15 and steps.min() < 15
    gpu_ctx.clear_trees()
    oracle.clear_trees()


def test_episodes_nonterminal_gridworld_benchmark_shape(gpu_ctx, oracle):
    """bench/non-terminal_gw.jl:11-18 scaled down: no terminal cells, n_iterations per step, every episode runs max_steps."""
    gw = m.SimpleGridWorld(terminate_from=())
    s = m.MCTSSolver(n_iterations=200, depth=20, exploration_constant=5.0, rng=1)
    ret, steps, final, st = run_episodes_both(gpu_ctx, oracle, s, gw, gridworld_roots(256), max_steps=20, env_seed=5)
    assert (steps == 20).all() and st.n_simulations == 200 * 256 * 20


def test_episodes_mountaincar(gpu_ctx, oracle):
    s = m.DPWSolver(n_iterations=100, depth=10, exploration_constant=2.0, rng=11, estimate_value=m.RolloutEstimator(m.EnergyPolicy(), max_depth=30))
    starts = mountaincar_roots(40) + [(0.6, 0.0)]
    ret, steps, final, st = run_episodes_both(gpu_ctx, oracle, s, MC, starts, max_steps=8, env_seed=2)
    assert steps[-1] == 0 and ret[-1] == 0.0


@pytest.mark.parametrize("model", ["mountaincar", "mountaincar-uct", "noisy-plugin", "noisy-plugin-unmerged"])
def test_kept_trees_of_hashed_models_follow_their_episodes(gpu_ctx, oracle, model):
    """keep_tree (src/dpw.jl:31-37) / reuse_tree on models without a bounded state table -- DPW's main use of kept trees: a continuous
    model planned along an episode. The closed-loop driver plans from every episode's current state with the tree it has grown so
    far (root lookup through the hash map, or a new node when the noise took the environment somewhere new); returns, steps,
    final states and every counter equal the oracle's over the whole episode."""
    est = m.RolloutEstimator(m.RandomSolver(), max_depth=20)
    if model == "mountaincar-uct":
        mdp, s = MC, m.MCTSSolver(n_iterations=80, depth=10, exploration_constant=8.0, rng=11, estimate_value=est, reuse_tree=True)
    else:
        mdp = MC if model == "mountaincar" else m.NoisyMountainCar(noise=0.004)
        s = m.DPWSolver(n_iterations=80, depth=10, exploration_constant=8.0, rng=11, estimate_value=est, keep_tree=True,
                        check_repeat_state=not model.endswith("unmerged"))
    starts = mountaincar_roots(24)
    ret, steps, final, st = run_episodes_both(gpu_ctx, oracle, s, mdp, starts, max_steps=6, env_seed=4)
    assert steps.max() == 6 and 24 * 80 * 3 < st.n_simulations <= 24 * 80 * 6  # (an episode that reaches the goal stops planning)
    # the room a kept tree has is finite: with space for one call only the second plan of the episode overflows, loudly
    gpu_ctx.clear_trees()
    s.keep_tree_calls = 1
    with pytest.raises(m.MCTSB200Error) as ei:
        cfg, keep = m.build_cfg(s, mdp)
        gpu_ctx.run_episodes(cfg, mdp.mdp_id, mdp.encode_states(starts), 6, 4)
    assert ei.value.status == m.abi.ERR_CAPACITY
    gpu_ctx.clear_trees()
    oracle.clear_trees()


@pytest.mark.parametrize("ahead,det", [("0", "1"), ("1", "0"), ("1", "1")])
def test_uct_both_descent_orders(gpu_ctx, oracle, monkeypatch, ahead, det):
    monkeypatch.setenv("MCTSB200_AHEAD", ahead)
    monkeypatch.setenv("MCTSB200_DET", det)
    s = m.MCTSSolver(n_iterations=300, depth=20, exploration_constant=5.0, rng=3)
    run_both(gpu_ctx, oracle, s, GW, gridworld_roots(200), compare_trees=range(0, 200, 17))
    for pol in (m.ConstantPolicy("up"), m.SeekTarget((9, 3)), m.RandomSolver()):
        s = m.MCTSSolver(n_iterations=300, depth=20, exploration_constant=5.0, estimate_value=m.RolloutEstimator(pol))
        run_both(gpu_ctx, oracle, s, GW_DET, gridworld_roots(200), compare_trees=range(0, 200, 17))


def test_visit_counts_beyond_the_ucb_tables(gpu_ctx, oracle, monkeypatch):
    """selections whose visit counts lie beyond the sqrt(log N) / 1/sqrt(n) tables take the reference's exact expression"""
    monkeypatch.setenv("MCTSB200_UCB_TABLE_ENTRIES", "2048")
    for s in (m.MCTSSolver(n_iterations=5000, depth=8, exploration_constant=5.0, rng=5),
              m.DPWSolver(n_iterations=5000, depth=8, exploration_constant=5.0, rng=5)):
        for fast in ("1", "0"):
            monkeypatch.setenv("MCTSB200_FAST", fast)
            _, _, _, st = run_both(gpu_ctx, oracle, s, GW, gridworld_roots(64), compare_trees=range(0, 64, 9))
            assert st.kernel_variant == int(fast) and st.n_select_exact > 1000


# ---- MDP plugins registered at run time (mctsb200_mdp_register: NVRTC-compiled tile kernel) -----------------------------------
def _same_results(r1, r2):
    a1, q1, s1, st1 = r1
    a2, q2, s2, st2 = r2
    assert np.array_equal(a1, a2) and np.array_equal(q1.view(np.int64), q2.view(np.int64)) and np.array_equal(s1, s2)
    from helpers import COUNTER_KEYS

    for k in COUNTER_KEYS:
        assert getattr(st1, k) == getattr(st2, k), k


@pytest.mark.parametrize("lanes", [None, 1, 4, 32])
def test_plugin_mountaincar_matches_builtin_and_oracle(gpu_ctx, oracle, lanes):
    """The built-in MountainCar re-registered as a run-time plugin (plugins/mountaincar.cuh): bit-identical to the oracle -- and so
    to the ahead-of-time kernels -- for DPW, vanilla UCT and a policy implemented by the plugin, for several tile widths."""
    mcp = m.MountainCarPlugin()
    roots = mountaincar_roots(24)
    for mdp, pol in ((MC, m.EnergyPolicy()), (mcp, m.PluginPolicy(100))):
        r_dpw = run_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=400, depth=50, rng=8), mdp, roots, lanes=lanes)
        r_uct = run_both(gpu_ctx, oracle, m.MCTSSolver(n_iterations=300, depth=30, exploration_constant=10.0, rng=9), mdp, roots, lanes=lanes)
        r_pol = run_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=300, depth=50, rng=8, estimate_value=m.RolloutEstimator(pol, max_depth=100)),
                         mdp, roots, lanes=lanes)
        if mdp is MC:
            built_in = (r_dpw, r_uct, r_pol)
        else:
            for x, y in zip(built_in, (r_dpw, r_uct, r_pol)):
                _same_results(x, y)
            assert r_dpw[3].kernel_variant == 0


@pytest.mark.parametrize("lanes", [None, 1, 8])
def test_plugin_noisy_mountaincar_bit_exact(gpu_ctx, oracle, lanes):
    """A continuous-state MDP with stochastic transitions as a plugin (plugins/noisy_mountaincar.cuh, max_branch = 0): state widening
    really widens (every gen call is a new state). DPW defaults, DPW with unmerged states, vanilla UCT, leaf-parallel rollouts."""
    mdp = m.NoisyMountainCar()
    roots = mountaincar_roots(20, seed=3)
    a, q, s, st = run_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=500, depth=30, rng=4), mdp, roots, lanes=lanes)
    assert st.n_transition_samples > 0  # re-sampling of existing transitions happened (src/dpw.jl:140)
    run_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=300, depth=20, k_state=2.0, alpha_state=0.3, check_repeat_state=False, rng=5), mdp, roots, lanes=lanes)
    run_both(gpu_ctx, oracle, m.MCTSSolver(n_iterations=300, depth=20, exploration_constant=5.0, rng=6), mdp, roots, lanes=lanes)
    if lanes in (None, 8):
        run_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=200, depth=20, rng=7, rollouts_per_leaf=8), mdp, roots, lanes=lanes)


def test_plugin_episodes_and_root_parallel(gpu_ctx, oracle):
    mdp = m.NoisyMountainCar(noise=0.004)
    s = m.DPWSolver(n_iterations=100, depth=10, exploration_constant=2.0, rng=11, estimate_value=m.RolloutEstimator(m.PluginPolicy(100), max_depth=30))
    starts = mountaincar_roots(40) + [(0.6, 0.0)]
    ret, steps, final, st = run_episodes_both(gpu_ctx, oracle, s, mdp, starts, max_steps=8, env_seed=2)
    assert steps[-1] == 0 and ret[-1] == 0.0
    # root-parallel ensemble (config 5's call shape) on a plugin MDP
    from helpers import oracle_id

    cfg, _ = m.build_cfg(m.MCTSSolver(n_iterations=200, depth=20, exploration_constant=5.0, rng=3), mdp)
    gpu_ctx.configure(mdp.mdp_id, mdp.params())
    oracle.configure(oracle_id(mdp), mdp.params())
    root = mdp.encode_states([(-0.5, 0.0)])
    n1, nq1, _ = gpu_ctx.plan_root_parallel(cfg, mdp.mdp_id, root, 64)
    n2, nq2, _ = oracle.plan_root_parallel(cfg, oracle_id(mdp), root, 64, n_threads=4)
    assert np.array_equal(n1, n2) and np.array_equal(nq1.view(np.int64), nq2.view(np.int64))  # integer partial sums: exact


def test_plugin_needs_configure_and_refuses_dense_options(gpu_ctx):
    src = m.MountainCarPlugin().source
    mdp = m.PluginMDP(name="Unconfigured", source=src, struct_name="MountainCarPlugin", actions=(-1.0, 0.0, 1.0),
                      params_struct=m.abi.MountainCarParams(0.99, -1.0, 100.0))
    cfg, _ = m.build_cfg(m.MCTSSolver(n_iterations=10, depth=5), mdp)
    with pytest.raises(m.MCTSB200Error) as e:
        gpu_ctx.plan(cfg, mdp.mdp_id, mdp.encode_states([(-0.5, 0.0)]))
    assert e.value.status == m.abi.ERR_STATE
    gpu_ctx.configure(mdp.mdp_id, mdp.params())
    gpu_ctx.plan(cfg, mdp.mdp_id, mdp.encode_states([(-0.5, 0.0)]))
    # kept trees work for hashed-state plugins (bounded room, see test_kept_trees_of_hashed_models_follow_their_episodes); what
    # needs a dense state index does not
    cfg.keep_tree = 1
    gpu_ctx.plan(cfg, mdp.mdp_id, mdp.encode_states([(-0.5, 0.0)]))
    gpu_ctx.clear_trees()
    cfg.keep_tree = 0
    tab = np.zeros(3, dtype=np.float64)
    import ctypes as C

    cfg.init_q_table = tab.ctypes.data_as(C.POINTER(C.c_double))
    with pytest.raises(m.MCTSB200Error) as e:
        gpu_ctx.plan(cfg, mdp.mdp_id, mdp.encode_states([(-0.5, 0.0)]))
    assert e.value.status == m.abi.ERR_UNSUPPORTED


@pytest.mark.parametrize("actions", [(-1.0, 1.0), (-1.0, 0.0, 1.0), (-1.0, -0.5, 0.5, 1.0)])
def test_plugin_every_host_shape(gpu_ctx, oracle, actions):
    """The host half of the library exists for 2- and 4-double states x 2..4 actions (capi.cu: MCTSB200_FOR_PLUGIN). The noisy
    MountainCar with 2, 3 and 4 accelerations against the oracle bit-for-bit (2-double states), and the same model carrying two
    padding coordinates (4-double states): the state map hashes other keys, the search must not notice."""
    m2 = m.NoisyMountainCar(actions=actions, state_width=2)
    m4 = m.NoisyMountainCar(actions=actions, state_width=4)
    roots2 = mountaincar_roots(16, seed=len(actions))
    roots4 = [(x, v, 7.0, -3.0) for x, v in roots2]
    for solver in (m.DPWSolver(n_iterations=300, depth=25, rng=21), m.MCTSSolver(n_iterations=200, depth=15, exploration_constant=3.0, rng=22)):
        r2 = run_both(gpu_ctx, oracle, solver, m2, roots2)
        t2 = [gpu_ctx.tree_export(i, m2.state_dtype, 2) for i in range(len(roots2))]
        cfg, _ = m.build_cfg(solver, m4)
        gpu_ctx.configure(m4.mdp_id, m4.params())
        r4 = gpu_ctx.plan(cfg, m4.mdp_id, m4.encode_states(roots4), 0, want_status=True)
        _same_results(r2, r4)
        for i in range(len(roots4)):
            t4 = gpu_ctx.tree_export(i, m4.state_dtype, 4)
            for k in t2[i]:
                if k == "s_labels":
                    lab = t4[k].reshape(-1, 4)
                    assert np.array_equal(lab[:, :2].view(np.int64), t2[i][k].reshape(-1, 2).view(np.int64))
                    assert (lab[:, 2] == 7.0).all() and (lab[:, 3] == -3.0).all()
                else:
                    assert np.array_equal(t4[k], t2[i][k]), k
    # closed-loop episodes (the run-time compiled episode kernels) on the 4-double shape against the 2-double one
    s = m.DPWSolver(n_iterations=60, depth=10, rng=23)
    ret2, steps2, final2, _ = m.solve(s, m2, ctx=gpu_ctx).run_episodes(roots2, 5, 9)
    ret4, steps4, final4, _ = m.solve(s, m4, ctx=gpu_ctx).run_episodes(roots4, 5, 9)
    assert np.array_equal(ret2.view(np.int64), ret4.view(np.int64)) and np.array_equal(steps2, steps4)
    assert [f[:2] for f in final4] == list(final2) and all(f[2:] == (7.0, -3.0) for f in final4)


def test_config1_root_q_within_3_sigma_over_1000_seeds(gpu_ctx, oracle):
    """BASELINE config 1 (README case: one root, n_iterations=50, depth=20, c=5, random rollouts on the default stochastic
    GridWorld) over 1 000 seeds. Same seeds: the device and the oracle agree bit-for-bit (stronger than the north star asks for a
    stochastic model). Independent seeds -- how a run compares with the reference's own MersenneTwister-driven search: the mean
    root Q of every action lies within 3 sigma of the other ensemble's (tolerance = 3 * sqrt(var_a/n + var_b/n))."""
    n = 1000
    roots = [(1, 2)] * n
    solver_a = m.MCTSSolver(n_iterations=50, depth=20, exploration_constant=5.0, rng=0)
    run_both(gpu_ctx, oracle, solver_a, GW, roots, compare_trees=range(0, n, 97))
    qa = np.array([gpu_ctx.root_stats(i)["q"] for i in range(n)])
    solver_b = m.MCTSSolver(n_iterations=50, depth=20, exploration_constant=5.0, rng=12345)
    cfg, _ = m.build_cfg(solver_b, GW)
    oracle.configure(GW.mdp_id, GW.params())
    oracle.plan(cfg, GW.mdp_id, GW.encode_states(roots), n_threads=4)
    qb = np.array([oracle.root_stats(i)["q"] for i in range(n)])
    assert qa.shape == qb.shape == (n, 4)
    tol = 3.0 * np.sqrt(qa.var(axis=0, ddof=1) / n + qb.var(axis=0, ddof=1) / n)
    assert (np.abs(qa.mean(axis=0) - qb.mean(axis=0)) <= tol).all(), (qa.mean(axis=0), qb.mean(axis=0), tol)
    assert (tol > 0).all() and not np.array_equal(qa, qb)


def test_plugin_max_time_callback_and_value_queries(gpu_ctx, oracle):
    """The host hooks between launches (max_time chunks, per-iteration callback; src/vanilla.jl:229-235, src/dpw.jl:44-56) and the
    tree queries on a plugin MDP: the chunked launches go through the run-time compiled kernels, and a search cut into chunks
    builds the same tree as an unchunked one."""
    mdp = m.NoisyMountainCar()
    root = (-0.5, 0.0)
    p = m.solve(m.DPWSolver(n_iterations=2**63 - 1, depth=20, max_time=0.2), mdp, ctx=gpu_ctx)
    a = p.action(root)
    assert a in mdp.actions and 0 < p.last_stats.iterations_done < 2**24
    calls = []
    pc = m.solve(m.MCTSSolver(n_iterations=40, depth=15, exploration_constant=3.0, rng=4, callback=lambda pl, i: calls.append((i, len(pl.tree.total_n))),
                              callback_every=16), mdp, ctx=gpu_ctx)
    a1, info1 = pc.action_info(root)
    assert [c[0] for c in calls] == [16, 32, 40] and calls[0][1] <= calls[-1][1]
    pu = m.solve(m.MCTSSolver(n_iterations=40, depth=15, exploration_constant=3.0, rng=4), mdp, ctx=gpu_ctx)
    a2, info2 = pu.action_info(root)
    assert a1 == a2 and info1["best_Q"] == info2["best_Q"]
    assert pu.value(root) == info2["best_Q"] and pu.value(root, a2) == info2["best_Q"]
    ranked = m.ranked_actions(pu, root)
    assert ranked[0][0] == a2 and len(ranked) == 3


def test_plugin_params_may_carry_device_pointers(gpu_ctx):
    """A plugin's Params are an opaque POD: they may hold pointers to tables the caller keeps in device memory (a tabular model,
    a learned value table, ...). The noisy MountainCar with its action table behind a device pointer plans exactly like the one
    that carries the table inside Params."""
    import ctypes as C

    import torch

    ref = m.NoisyMountainCar()
    src = ref.source.replace("double push[4];", "const double* push;").replace("p.push[ai]", "__ldg(p.push + ai)")
    assert src != ref.source

    class P(C.Structure):
        _fields_ = [("discount", C.c_double), ("cost", C.c_double), ("jackpot", C.c_double), ("noise", C.c_double), ("push", C.c_void_p),
                    ("n_actions", C.c_int32), ("pad", C.c_int32)]

    table = torch.tensor([ref.params_struct.push[i] for i in range(3)], dtype=torch.float64, device="cuda")
    rp = ref.params_struct
    mdp = m.PluginMDP(name="NoisyMountainCarTable", source=src, struct_name="NoisyMountainCar", actions=ref.actions,
                      params_struct=P(rp.discount, rp.cost, rp.jackpot, rp.noise, table.data_ptr(), 3, 0), state_width=2, max_branch=0)
    roots = mountaincar_roots(24, seed=5)
    out = []
    for model in (ref, mdp):
        res = []
        for solver in (m.DPWSolver(n_iterations=300, depth=25, rng=31), m.MCTSSolver(n_iterations=200, depth=15, exploration_constant=3.0, rng=32)):
            cfg, _ = m.build_cfg(solver, model)
            gpu_ctx.configure(model.mdp_id, model.params())
            res.append(gpu_ctx.plan(cfg, model.mdp_id, model.encode_states(roots), 0, want_status=True))
        out.append(res)
    for x, y in zip(*out):
        _same_results(x, y)
    del table


def _visible_devices():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_multi_device_ctx_equals_one_device(gpu_ctx, oracle, n_dev):
    """mctsb200_ctx_create_multi: the same entry points on a ctx over n_dev devices -- roots sharded inside the
    library, root-parallel partial sums merged by the library's own ncclAllReduce -- return what ONE device returns, bit for
    bit: actions, Q, trees addressed by global index, and the merged ensemble sums (integer words, so no reduction order can
    show). n_dev = 1 runs everywhere; more devices run where the box has them."""
    if _visible_devices() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    from helpers import assert_trees_equal, bits

    multi = m.Context(list(range(n_dev)))
    try:
        for mdp, solver, roots in (
            (GW, m.MCTSSolver(n_iterations=200, depth=20, exploration_constant=5.0, rng=4), gridworld_roots(301)),
            (GW, m.DPWSolver(n_iterations=200, depth=20, exploration_constant=10.0, k_state=4.0, alpha_state=1 / 8, k_action=4.0, alpha_action=1 / 8, rng=4),
             gridworld_roots(301)),
            (MC, m.DPWSolver(n_iterations=300, depth=30, rng=4), mountaincar_roots(67)),
        ):
            cfg, keep = m.build_cfg(solver, mdp)
            r = mdp.encode_states(roots)
            for c in (gpu_ctx, multi):
                c.configure(mdp.mdp_id, mdp.params())
            a1, q1, s1, st1 = gpu_ctx.plan(cfg, mdp.mdp_id, r, 1000, want_status=True)
            a2, q2, s2, st2 = multi.plan(cfg, mdp.mdp_id, r, 1000, want_status=True)
            assert np.array_equal(a1, a2) and np.array_equal(bits(q1), bits(q2)) and np.array_equal(s1, s2)
            for k in ("n_simulations", "n_tree_levels", "n_rollout_steps", "n_state_inserts", "algorithmic_bytes", "n_trees"):
                assert getattr(st1, k) == getattr(st2, k), k
            for i in (0, len(roots) // 2, len(roots) - 1):
                assert_trees_equal(gpu_ctx.tree_export(i, mdp.state_dtype, mdp.state_width), multi.tree_export(i, mdp.state_dtype, mdp.state_width), f"tree {i}")
            # root-parallel ensemble of one root, against one device and against the oracle
            root = mdp.encode_states(roots[:1])
            n1, nq1, _ = gpu_ctx.plan_root_parallel(cfg, mdp.mdp_id, root, 93, tree_index_base=5)
            n2, nq2, _ = multi.plan_root_parallel(cfg, mdp.mdp_id, root, 93, tree_index_base=5)
            oracle.configure(mdp.mdp_id, mdp.params())
            n3, nq3, _ = oracle.plan_root_parallel(cfg, mdp.mdp_id, root, 93, tree_index_base=5, n_threads=4)
            assert np.array_equal(n1, n2) and np.array_equal(bits(nq1), bits(nq2))
            assert np.array_equal(n1, n3) and np.array_equal(bits(nq1), bits(nq3))
            # fewer trees than devices: the empty shards still take part in the allreduce
            n4, nq4, _ = multi.plan_root_parallel(cfg, mdp.mdp_id, root, 1, tree_index_base=5)
            n5, nq5, _ = gpu_ctx.plan_root_parallel(cfg, mdp.mdp_id, root, 1, tree_index_base=5)
            assert np.array_equal(n4, n5) and np.array_equal(bits(nq4), bits(nq5))
            del keep
    finally:
        multi.close()


BOX_ROOTS = [(0.0, 0.0, 0.0), (1.5, -2.0, 0.5), (-3.0, 3.0, -1.0), (4.0, 4.0, 2.0), (0.25, -0.125, 3.0)]


@pytest.mark.parametrize("lanes", [None, 1, 4, 32])
def test_box_actions_bit_exact(gpu_ctx, oracle, lanes):
    """Continuous action spaces (test/discussion_514.jl): a state node owns up to k_action * N^alpha_action action nodes, next_action
    draws a vector, the tile's lanes share the repeat check and the UCB scan over a node's children. Action vectors, Q, counters and
    whole trees against the oracle, for the reference test's model and for one where the search has something to find."""
    from test_host_emu_parity import _run_box_both

    roots = BOX_ROOTS * 8
    v, _, st = _run_box_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=200, depth=10, rng=3), m.GaussianBox(), roots, lanes=lanes, trees=range(0, 40, 7))
    assert all(m.GaussianBox().action_in_space(a) for a in v) and st.n_simulations == 200 * 40
    steer = m.GaussianBox(drift=0.2, cost=1.0, var=(0.05, 0.1, 0.02))
    for opts in (dict(), dict(exploration_constant=5.0, k_action=4.0, alpha_action=0.25, k_state=2.0, alpha_state=0.3), dict(check_repeat_state=False),
                 dict(check_repeat_action=False), dict(init_Q=-3.0, init_N=2), dict(estimate_value=-1.5), dict(exploration_constant=0.0)):
        kw = dict(n_iterations=300, depth=8, rng=11)
        kw.update(opts)
        _run_box_both(gpu_ctx, oracle, m.DPWSolver(**kw), steer, roots, base=77, lanes=lanes, trees=range(0, 40, 9))


def test_box_actions_many_children_at_the_root(gpu_ctx, oracle):
    """A long search from one root: hundreds of action nodes under the root (k_action * N^0.5), children lists and push logs that
    moved through several arena chunks."""
    from test_host_emu_parity import _run_box_both

    steer = m.GaussianBox(drift=0.3, cost=0.5)
    v, _, st = _run_box_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=5000, depth=6, exploration_constant=8.0, rng=2), steer, BOX_ROOTS[:3], trees=[0, 2])
    t = gpu_ctx.tree_export(0, np.float64, 4)
    assert t["child_offsets"][1] - t["child_offsets"][0] > 300


@pytest.mark.parametrize("lanes", [None, 1, 4, 32])
def test_plugin_continuous_actions_bit_exact(gpu_ctx, oracle, lanes):
    """A run-time plugin with a continuous action space (mctsb200_mdp_register_continuous; plugins/continuous_pendulum.cuh: the inverted
    pendulum with the torque as a Box([-50], [50]) action): the plugin's next_action / gen / policy_action run inside the NVRTC-compiled
    dpw_box_iteration; action vectors, Q, counters and whole trees against the oracle's restatement."""
    from test_host_emu_parity import _run_box_both

    cp = m.ContinuousPendulum()
    rng = np.random.default_rng(9)
    roots = [((u - 0.5) * 0.1, (v - 0.5) * 0.1) for u, v in rng.random((24, 2))]
    v, _, st = _run_box_both(gpu_ctx, oracle, m.DPWSolver(n_iterations=400, depth=10, exploration_constant=10.0, alpha_state=1 / 8, rng=4), cp, roots,
                             lanes=lanes, trees=range(0, 24, 5))
    assert all(cp.action_in_space(a) for a in v) and st.n_simulations == 400 * 24
    # a torque cost (the best action is interior), a feedback controller as the rollout policy, unmerged actions and states
    cp2 = m.ContinuousPendulum(effort=1e-4, noise=4.0, gain=150.0, name="ContinuousPendulumEffort")
    for opts in (dict(estimate_value=m.RolloutEstimator(m.PluginPolicy(100), max_depth=30)), dict(check_repeat_action=False),
                 dict(check_repeat_state=False, k_action=6.0, alpha_action=0.3), dict(estimate_value=0.0, init_Q=-0.5, init_N=1)):
        kw = dict(n_iterations=300, depth=12, exploration_constant=2.0, rng=6)
        kw.update(opts)
        _run_box_both(gpu_ctx, oracle, m.DPWSolver(**kw), cp2, roots, base=500, lanes=lanes, trees=range(0, 24, 7))


@pytest.mark.parametrize("which,lanes", [("box", None), ("box", 1), ("box", 32), ("plugin", None), ("plugin", 4)])
def test_continuous_actions_kept_trees(gpu_ctx, oracle, which, lanes):
    """keep_tree (src/dpw.jl:31-37) over a continuous action space, built-in model and run-time plugin: four calls, every root moves on
    to a successor its tree already holds, ERR_CAPACITY once the room for keep_tree_calls calls is used up."""
    from test_host_emu_parity import BOX_ROOTS, _box_kept_trees

    if which == "box":
        _box_kept_trees(gpu_ctx, oracle, m.GaussianBox(drift=0.2, cost=1.0, var=(0.05, 0.1, 0.02)), BOX_ROOTS * 3, lanes=lanes)
    else:
        _box_kept_trees(gpu_ctx, oracle, m.ContinuousPendulum(), [(0.01 * i, -0.02 * i) for i in range(-4, 5)], lanes=lanes)


@pytest.mark.parametrize("which", ["box", "box-keep", "plugin", "plugin-keep", "pendulum-3-torques"])
def test_continuous_actions_episodes(gpu_ctx, oracle, which):
    """mctsb200_run_episodes over a continuous action space: every step plans, the environment steps with the chosen action VECTOR.
    Returns, episode lengths, final states and counters against the oracle (also for the three-torque InvertedPendulum plugin)."""
    keep = which.endswith("keep")
    if which.startswith("box"):
        mdp = m.GaussianBox(drift=0.3, cost=0.5, var=(0.05, 0.1, 0.02))
        from test_host_emu_parity import BOX_ROOTS
        starts = BOX_ROOTS * 6
        s = m.DPWSolver(n_iterations=120, depth=6, exploration_constant=3.0, rng=8, keep_tree=keep, keep_tree_calls=8)
    else:
        mdp = m.ContinuousPendulum(noise=20.0) if which.startswith("plugin") else m.InvertedPendulum(noise=20.0, name="InvertedPendulumNoisy")
        starts = [(0.05 * i, -0.04 * i) for i in range(-12, 13)]
        s = m.DPWSolver(n_iterations=150, depth=8, exploration_constant=5.0, alpha_state=1 / 8, rng=5, keep_tree=keep, keep_tree_calls=8,
                        estimate_value=m.RolloutEstimator(m.PluginPolicy(100), max_depth=15))
    ret, steps, final, st = run_episodes_both(gpu_ctx, oracle, s, mdp, starts, max_steps=7, env_seed=77, base=11)
    assert steps.max() == 7


def test_plugin_continuous_actions_refusals(gpu_ctx):
    """What a continuous action space cannot do, whoever brings the model: vanilla UCT, DPW without action widening, per-action sums."""
    cp = m.ContinuousPendulum()
    gpu_ctx.configure(cp.mdp_id, cp.params())
    r = cp.encode_states([(0.01, 0.0)])
    for solver in (m.MCTSSolver(n_iterations=10, depth=5), m.DPWSolver(n_iterations=10, depth=5, enable_action_pw=False)):
        cfg, keep = m.build_cfg(solver, cp)
        with pytest.raises(m.MCTSB200Error) as ei:
            gpu_ctx.plan(cfg, cp.mdp_id, r)
        assert ei.value.status == m.abi.ERR_UNSUPPORTED
    cfg, keep = m.build_cfg(m.DPWSolver(n_iterations=10, depth=5), cp)
    with pytest.raises(m.MCTSB200Error) as ei:  # the root-parallel sums are defined per action of a finite set
        gpu_ctx.plan_root_parallel(cfg, cp.mdp_id, r[0], 8)
    assert ei.value.status == m.abi.ERR_UNSUPPORTED
    gpu_ctx.plan(cfg, cp.mdp_id, r)  # ... and the ctx is still good
    assert gpu_ctx.root_action_vectors(0, 1, 1).shape == (1, 1)


@pytest.mark.parametrize("fast", ["1", "0"])
def test_next_action_as_a_state_only_function(gpu_ctx, oracle, monkeypatch, fast):
    from test_host_emu_parity import test_emu_next_action_as_a_state_only_function as body

    body(gpu_ctx, oracle, monkeypatch, fast)


@pytest.mark.parametrize("fast", ["1", "0"])
def test_function_policy_is_a_table_on_the_device(gpu_ctx, oracle, monkeypatch, fast):
    from test_host_emu_parity import test_emu_function_policy_is_a_table_on_the_device as body

    body(gpu_ctx, oracle, monkeypatch, fast)


@pytest.mark.parametrize("mdp_name", ["gridworld", "mountaincar"])
def test_vis_stats_like_record_visit(gpu_ctx, oracle, mdp_name):
    from test_host_emu_parity import test_emu_vis_stats_like_record_visit as body

    body(gpu_ctx, oracle, mdp_name)


@pytest.mark.parametrize("lanes", [None, 4])
def test_plugin_inverted_pendulum_bit_exact(gpu_ctx, oracle, lanes):
    """POMDPModels.InvertedPendulum as a shipped run-time plugin (plugins/inverted_pendulum.cuh; det_sin + force noise), with the
    solver of notebooks/Test_Visualization.ipynb cell 4: DPWSolver(depth=10, exploration_constant=10., alpha_state=1/8)."""
    ip = m.InvertedPendulum()
    rng = np.random.default_rng(4)
    roots = [((u - 0.5) * 0.1, (v - 0.5) * 0.1) for u, v in rng.random((24, 2))]  # initialstate: uniform in [-0.05, 0.05]^2
    s = m.DPWSolver(n_iterations=400, depth=10, exploration_constant=10.0, alpha_state=1 / 8, rng=4)
    run_both(gpu_ctx, oracle, s, ip, roots, lanes=lanes, compare_trees=range(0, 24, 5), literal=True)
    s2 = m.MCTSSolver(n_iterations=300, depth=10, exploration_constant=10.0, rng=4, estimate_value=m.RolloutEstimator(m.PluginPolicy(100), max_depth=30))
    run_both(gpu_ctx, oracle, s2, ip, roots, lanes=lanes, compare_trees=range(0, 24, 5))
