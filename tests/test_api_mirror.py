"""The host-side mirror of the reference's public API for this path (mcts.jl_b200/solvers.py), exercised the way
the reference's own tests exercise MCTS.jl -- test/runtests.jl, test/options.jl, test/dpw_test.jl, test/other.jl --
on the host-emulation build (CPU suite) and, marked gpu, on the shipped CUDA library."""
import warnings

import numpy as np
import pytest

import mcts_jl_b200 as m

GW = m.SimpleGridWorld()


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def ctx(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_ctx")
    return request.getfixturevalue("gpu_ctx")


def test_vanilla_smoke_like_runtests(ctx):
    """test/runtests.jl:17-49: action(policy, GWPos(1,1)), root has id 1, info[:tree], clear_tree!."""
    solver = m.MCTSSolver(n_iterations=100, depth=20, exploration_constant=5.0, rng=1)
    policy = m.solve(solver, GW, ctx=ctx)
    state = (1, 1)
    a = policy.action(state)
    assert a in GW.actions
    a2, info = policy.action_info(state)
    assert a2 in GW.actions and "tree" in info and "best_Q" in info
    tree = info["tree"]
    assert tree.state_map[state] == 1  # test/runtests.jl:44-45
    assert len(tree.total_n) <= 101 and len(tree.n) == 4 * len(tree.total_n)
    assert int(tree.total_n[0]) == int(tree.n[:4].sum()) >= 100
    m.clear_tree(policy)
    assert policy.tree is None


def test_readme_usage(ctx):
    """README.md:37-40: MCTSSolver(n_iterations=50, depth=20, exploration_constant=5.0), root SA[1,2]."""
    planner = m.solve(m.MCTSSolver(n_iterations=50, depth=20, exploration_constant=5.0), GW, ctx=ctx)
    assert planner.action((1, 2)) in ("up", "down", "left", "right")


def test_option_matrix_like_options_jl(ctx, capsys):
    """test/options.jl: every device-expressible form of init_Q / init_N / estimate_value runs for both solvers;
    a String is a MethodError (TypeError here); host closures without a device form are refused."""
    for S in (m.MCTSSolver, m.DPWSolver):
        base = dict(n_iterations=20, depth=5)
        for kw in (dict(init_Q=1.0), dict(init_N=3), dict(estimate_value=2.5),
                   dict(init_Q=lambda mdp, s, a: 1.5), dict(init_N=lambda mdp, s, a: 2),
                   dict(estimate_value=lambda mdp, s, d: 0.25),
                   dict(estimate_value=m.RolloutEstimator(m.RandomPolicy())),
                   dict(estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"))),
                   dict(estimate_value=m.RolloutEstimator(m.RandomSolver(), max_depth=-1))):
            assert m.solve(S(**base, **kw), GW, ctx=ctx).action((1, 1)) in GW.actions
        for bad in (dict(init_Q="bad"), dict(init_N="bad"), dict(estimate_value="bad")):
            with pytest.raises(TypeError):
                m.solve(S(**base, **bad), GW, ctx=ctx)
    # enable_tree_vis records tree._vis_stats on the device (src/vanilla.jl:268-270); callback and timer run on the host between
    # launches (tested below)
    _, info = m.solve(m.MCTSSolver(n_iterations=30, depth=5, enable_tree_vis=True), GW, ctx=ctx).action_info((1, 1))
    assert sum(info["tree"]._vis_stats.values()) >= 30
    with pytest.raises(m.MCTSB200Error):
        m.solve(m.DPWSolver(reset_callback=lambda mdp, s: None), GW, ctx=ctx)
    # show_progress (src/dpw.jl:45-53) is a host hook between launches: the same tree as without it
    base_dpw = dict(n_iterations=60, depth=5, exploration_constant=5.0, rng=4)
    a_quiet, i_quiet = m.solve(m.DPWSolver(**base_dpw), GW, ctx=ctx).action_info((2, 2))
    a_loud, i_loud = m.solve(m.DPWSolver(show_progress=True, **base_dpw), GW, ctx=ctx).action_info((2, 2))
    assert a_quiet == a_loud and i_quiet["best_Q"] == i_loud["best_Q"]
    assert "60/60 iterations" in capsys.readouterr().err
    # next_action=(mdp, s, snode)->:up (test/options.jl:50): a generator that looks at the state only is tabulated; every state node then
    # owns the one action it proposes
    p = m.solve(m.DPWSolver(n_iterations=60, depth=5, exploration_constant=5.0, next_action=lambda mdp, s, snode: "up"), GW, ctx=ctx)
    a, info = p.action_info((3, 3), tree_in_info=True)
    assert a == "up" and set(int(x) for x in info["tree"].a_labels) == {GW.action_index("up")}
    with pytest.raises(TypeError):  # test/options.jl:52: next_action="bad" is a MethodError
        m.solve(m.DPWSolver(next_action="bad"), GW, ctx=ctx)
    with pytest.raises(m.MCTSB200Error):  # the node itself is not available to a host function
        m.solve(m.DPWSolver(next_action=lambda mdp, s, snode: snode.tree), GW, ctx=ctx)
    with pytest.raises(m.MCTSB200Error):  # no dense state index to tabulate over
        m.solve(m.DPWSolver(next_action=lambda mdp, s, snode: 1.0), m.MountainCar(), ctx=ctx)
    # RolloutEstimator(x -> :up) (test/options.jl:29): a state-only function is tabulated over the dense state space
    assert m.solve(m.MCTSSolver(n_iterations=20, depth=5, estimate_value=m.RolloutEstimator(lambda s: "up")), GW, ctx=ctx).action((1, 1)) in GW.actions
    with pytest.raises(m.MCTSB200Error):  # ... which a hashed-state model does not have
        m.solve(m.MCTSSolver(estimate_value=m.RolloutEstimator(lambda s: 1.0)), m.MountainCar(), ctx=ctx)


def test_dpw_smoke_like_dpw_test(ctx):
    """test/dpw_test.jl: default, keep_tree + enable_action_pw=false; info keys of src/dpw.jl:58-67."""
    for kw in (dict(), dict(keep_tree=True, enable_action_pw=False), dict(tree_in_info=True)):
        p = m.solve(m.DPWSolver(n_iterations=50, depth=20, exploration_constant=5.0, **kw), GW, ctx=ctx)
        a, info = p.action_info((1, 1), tree_in_info=True)
        assert a in GW.actions
        assert {"search_time", "search_time_us", "tree_queries", "best_Q", "tree"} <= set(info)
        assert info["tree_queries"] == 50


def test_ranked_actions_like_other_jl(ctx):
    """test/other.jl:15,29: the top-ranked action is the chosen action, for both solvers."""
    for S in (m.MCTSSolver, m.DPWSolver):
        p = m.solve(S(n_iterations=200, depth=10, exploration_constant=5.0, rng=3), GW, ctx=ctx)
        a = p.action((3, 3))
        ranked = m.ranked_actions(p, (3, 3))
        assert ranked[0][0] == a
        assert [q for _, q in ranked] == sorted((q for _, q in ranked), reverse=True)


def test_zero_exploration_like_runtests(ctx):
    """test/runtests.jl:136-153: c = 0 runs without NaN for both solvers."""
    for S in (m.MCTSSolver, m.DPWSolver):
        p = m.solve(S(n_iterations=100, depth=10, exploration_constant=0.0), GW, ctx=ctx)
        a, info = p.action_info((2, 2))
        assert a in GW.actions and np.isfinite(info["best_Q"])


def test_dpw_terminal_root_and_default_action(ctx):
    """src/dpw.jl:22-27 raises on a terminal root; with a default_action the exception is captured (src/dpw.jl:68-71,
    test/runtests.jl:155-163)."""
    p = m.solve(m.DPWSolver(n_iterations=10, depth=5), GW, ctx=ctx)
    with pytest.raises(m.MCTSB200Error) as ei:
        p.action((-1, -1))
    assert ei.value.status == m.abi.ERR_TERMINAL_ROOT
    p = m.solve(m.DPWSolver(n_iterations=10, depth=5, default_action="up"), GW, ctx=ctx)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        a, info = p.action_info((-1, -1))
    assert a == "up" and "exception" in info and len(w) == 1


def test_value_like_vanilla_jl(ctx):
    """value(planner, s) = max_a Q(s,a), value(planner, s, a) = Q(s,a) (src/vanilla.jl:169-197)."""
    p = m.solve(m.MCTSSolver(n_iterations=300, depth=10, exploration_constant=5.0, rng=2), GW, ctx=ctx)
    a, info = p.action_info((8, 3))
    v = p.value((8, 3))
    assert v == pytest.approx(info["best_Q"])
    assert p.value((8, 3), a) == pytest.approx(v)
    with pytest.raises(KeyError):
        p.value((1, 10) if (1, 10) not in p.tree.state_map else (-5, -5))


def test_batch_extension_equals_single_calls(ctx):
    """action_batch plans every state in one call; with the same global root indices the trees are the same."""
    solver = m.MCTSSolver(n_iterations=80, depth=10, exploration_constant=5.0, rng=9)
    states = [(1, 1), (5, 5), (9, 2), (4, 6)]
    acts, infos = m.solve(solver, GW, ctx=ctx).action_info_batch(states)
    for i, s in enumerate(states):
        a1, info1 = m.solve(solver, GW, ctx=ctx).action_info_batch([s], root_index_base=i)
        assert a1[0] == acts[i] and info1[0]["best_Q"] == infos[i]["best_Q"]


def test_max_time_runs_at_least_one_iteration(ctx):
    """src/vanilla.jl:229-235: the timer is read after an iteration, so max_time = 0 still runs one."""
    p = m.solve(m.MCTSSolver(n_iterations=10**6, depth=5, max_time=0.0), GW, ctx=ctx)
    p.action((1, 1))
    assert 1 <= p.last_stats.iterations_done < 10**6
    assert int(p.last_stats.n_simulations) == p.last_stats.iterations_done


def test_timing_like_runtests(ctx):
    """test/runtests.jl:90-119: n_iterations = typemax(Int) with max_time: the search stops on the clock."""
    import time
    for S in (m.MCTSSolver, m.DPWSolver):
        p = m.solve(S(n_iterations=2**63 - 1, depth=10, max_time=0.3, exploration_constant=5.0), GW, ctx=ctx)
        p.action((1, 1))  # (first call: tables, module load)
        t0 = time.perf_counter()
        a = p.action((1, 1))
        dt = time.perf_counter() - t0
        assert a in GW.actions and abs(dt - 0.3) < 0.25
        assert 0 < p.last_stats.iterations_done < 2**24


def test_callbacks_like_runtests(ctx):
    """test/runtests.jl:165-180: callback(planner, iteration) runs once per iteration and can look at planner.tree."""
    calls, sizes = [], []

    def callback(planner, iteration_index):
        calls.append(iteration_index)
        sizes.append(len(planner.tree.total_n))

    policy = m.solve(m.MCTSSolver(n_iterations=10, callback=callback), GW, ctx=ctx)
    a = policy.action((1, 1))
    assert a in GW.actions and calls == list(range(1, 11))
    assert sizes == sorted(sizes) and sizes[-1] > 1  # the tree grows under the callback's eyes
    # coarser: every 4 iterations -> after 4, 8 and (the rest) 10
    calls.clear()
    m.solve(m.MCTSSolver(n_iterations=10, callback=callback, callback_every=4), GW, ctx=ctx).action((1, 1))
    assert calls == [4, 8, 10]
    # the hook is gone afterwards, and the chunked search built the same tree as an unchunked one
    calls.clear()
    p1 = m.solve(m.MCTSSolver(n_iterations=50, rng=9), GW, ctx=ctx)
    a1, i1 = p1.action_info((2, 2))
    assert calls == []
    p2 = m.solve(m.MCTSSolver(n_iterations=50, rng=9, callback=lambda p, i: None, callback_every=7), GW, ctx=ctx)
    a2, i2 = p2.action_info((2, 2))
    assert a1 == a2 and i1["best_Q"] == i2["best_Q"]


def test_custom_timer_decides_when_max_time_is_up(ctx):
    """solver.timer (src/vanilla.jl:25,226-232; src/dpw.jl:44-52): a fake clock that advances 1 s per reading stops the
    search after the first chunk -- at least one iteration always runs."""
    for S in (m.MCTSSolver, m.DPWSolver):
        ticks = [0.0]

        def fake():
            ticks[0] += 1.0
            return ticks[0]

        p = m.solve(S(n_iterations=100000, depth=5, max_time=1.5, timer=fake), GW, ctx=ctx)
        a = p.action((1, 1))
        assert a in GW.actions and 1 <= p.last_stats.iterations_done < 100000


def test_continuous_action_space_like_discussion_514(ctx):
    """test/discussion_514.jl: states(m) = Box, actions(m) = Box([-5,-5],[5,5]), transition = MvNormal(s, Diagonal([.1,.1,.1])),
    reward 0, discount 0.9; `policy = solve(DPWSolver(max_time=1.0), m); a = action(policy, [0.0,0.0,0.0]); @test a in actions(m)`."""
    mdp = m.GaussianBox()
    policy = m.solve(m.DPWSolver(max_time=0.2, n_iterations=200), mdp, ctx=ctx)
    a = policy.action((0.0, 0.0, 0.0))
    assert len(a) == 2 and mdp.action_in_space(a)
    a2, info = policy.action_info((1.0, -1.0, 0.5), tree_in_info=True)
    assert mdp.action_in_space(a2) and info["tree_queries"] >= 1
    tree = info["tree"]
    # the chosen action is the label of the root child with the largest Q (src/dpw.jl:163-173, :65)
    kids = tree.children(1)
    best = kids[int(np.argmax(tree.q[kids - 1]))]
    assert tuple(tree.a_label_vectors[best - 1]) == tuple(a2) and info["best_Q"] == tree.q[best - 1]
    # vanilla UCT enumerates actions(mdp, s): refused, not approximated
    with pytest.raises(m.MCTSB200Error) as ei:
        m.solve(m.MCTSSolver(n_iterations=10, depth=5), mdp, ctx=ctx).action((0.0, 0.0, 0.0))
    assert ei.value.status == m.abi.ERR_UNSUPPORTED
    # a model with something to find: the planner steers towards the origin (drift * a reduces |s|^2)
    steer = m.GaussianBox(drift=0.5, cost=1.0, var=(0.01, 0.01, 0.01))
    p2 = m.solve(m.DPWSolver(n_iterations=3000, depth=3, exploration_constant=20.0, rng=5), steer, ctx=ctx)
    a3 = p2.action((4.0, -4.0, 0.0))
    assert a3[0] < 0.0 < a3[1]


@pytest.mark.gpu
def test_continuous_action_space_of_a_caller_supplied_model(gpu_ctx):
    """The same through a run-time plugin (mctsb200_mdp_register_continuous): `action` returns the torque DPW chose, the exported tree
    carries vector labels, kept trees continue, closed-loop episodes keep the pendulum up longer than no control."""
    cp = m.ContinuousPendulum()
    policy = m.solve(m.DPWSolver(n_iterations=2000, depth=10, exploration_constant=10.0, alpha_state=1 / 8, rng=3, keep_tree=True), cp, ctx=gpu_ctx)
    (u,) = policy.action((0.03, 0.0))
    assert cp.action_in_space((u,))
    a2, info = policy.action_info((0.03, 0.0), tree_in_info=True)
    tree = info["tree"]
    kids = tree.children(1)
    best = kids[int(np.argmax(tree.q[kids - 1]))]
    assert tuple(tree.a_label_vectors[best - 1]) == tuple(a2) and len(kids) > 20
    assert tree.total_n[0] >= 4000  # the second call continued the first call's tree
    policy.clear_tree()
    # falling over costs -1 and ends the episode: planned episodes from a tilted start last longer than uncontrolled ones
    starts = [(0.15, 0.05)] * 16  # (the torque limit can hold the pendulum only below about 0.5 rad)
    planned = m.solve(m.DPWSolver(n_iterations=400, depth=10, exploration_constant=1.0, alpha_state=1 / 8, rng=1,
                                  estimate_value=m.RolloutEstimator(m.PluginPolicy(100), max_depth=20)), cp, ctx=gpu_ctx)
    _, steps, _, _ = planned.run_episodes(starts, max_steps=60, env_seed=5)
    idle = m.solve(m.DPWSolver(n_iterations=1, depth=1, rng=1, estimate_value=0.0), m.ContinuousPendulum(u_max=1e-9, name="IdlePendulum"), ctx=gpu_ctx)
    _, steps0, _, _ = idle.run_episodes(starts, max_steps=60, env_seed=5)
    assert steps.mean() > steps0.mean() + 3, (steps, steps0)  # (a 1 s search horizon sees the fall late: it delays it, no more)


def test_several_devices_behind_the_same_planner(emu_lib_path):
    """MCTSB200.jl: `solve(B200(solver, [0, 1]), mdp)` -- a planner over several devices of one process (mctsb200_ctx_create_multi):
    batch calls shard inside the library and return what one device returns; `action_root_parallel` merges the devices' trees.
    (Two ctxs on the emulator's one device here; real devices in tests/test_gpu_parity.py::test_multi_device_ctx_equals_one_device.)"""
    solver = m.MCTSSolver(n_iterations=60, depth=10, exploration_constant=5.0, rng=2)
    one = m.solve(solver, GW, ctx=m.Context(0, lib_path=emu_lib_path))
    two = m.solve(solver, GW, ctx=m.Context([0, 0], lib_path=emu_lib_path))
    states = [(1 + i % 10, 1 + i // 10) for i in range(23)]
    a1, i1 = one.action_info_batch(states)
    a2, i2 = two.action_info_batch(states)
    assert a1 == a2 and [x["best_Q"] for x in i1] == [x["best_Q"] for x in i2]
    assert np.array_equal(i1[17]["tree"].q.view(np.int64), i2[17]["tree"].q.view(np.int64))  # tree 17 lives on the second device
    r1, d1 = one.action_root_parallel((1, 1), 13)
    r2, d2 = two.action_root_parallel((1, 1), 13)
    assert r1 == r2 and d1["best_Q"] == d2["best_Q"] and np.array_equal(d1["sum_n"], d2["sum_n"])
    assert np.array_equal(d1["q"].view(np.int64), d2["q"].view(np.int64))
