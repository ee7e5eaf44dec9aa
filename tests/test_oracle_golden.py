"""Pins the CPU oracle against every known-answer vector the reference holds for this path
(tests/golden/reference_vectors.json, extracted from /root/reference notebooks and test/value_test.jl by
tests/golden/extract_notebook_vectors.py) and against independent restatements written here in numpy.

No reference test pins a whole seeded search bit-for-bit (they all ride on Julia's MersenneTwister), so
the stored values are reproduced through scenarios whose arithmetic does not depend on the RNG.
"""
import json
import os

import numpy as np
import pytest

import mcts_jl_b200 as m

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
GW = m.SimpleGridWorld()


def plan(oracle, solver, mdp, roots, **kw):
    cfg, keep = m.build_cfg(solver, mdp)
    oracle.configure(mdp.mdp_id, mdp.params())
    out = oracle.plan(cfg, mdp.mdp_id, mdp.encode_states(roots), **kw)
    del keep
    return out


# ---- Philox4x32-10 known answers (Random123 kat_vectors) -----------------------------------------
def test_philox_known_answers(oracle):
    kat = [
        ([0, 0], [0, 0, 0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 2, [0xFFFFFFFF] * 4, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0xA4093822, 0x299F31D0], [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for key, ctr, want in kat:
        assert list(oracle.philox_raw(key, ctr)) == want
    # stream contract: counter = (block, iteration, tree lo32, tree hi16<<16 | stream)
    w = oracle.philox(seed=(7 << 32) | 5, tree=(3 << 32) | 9, iteration=11, stream_id=2, n_blocks=2)
    assert list(w[:4]) == list(oracle.philox_raw([5, 7], [0, 11, 9, (3 << 16) | 2]))
    assert list(w[4:]) == list(oracle.philox_raw([5, 7], [1, 11, 9, (3 << 16) | 2]))


# ---- notebook Domain_Knowledge_Example cell 3: incremental-mean arithmetic -------------------------
def test_notebook_cell3_incremental_mean_values(oracle):
    """init_N=3, init_Q=11.73 and zero returns give exactly the stored Q for each N."""
    stored = {r["N"]: r["Q"] for r in VEC["dk_cell3"]["rows"]}
    assert stored == {3: 11.73, 4: 8.7975, 5: 7.037999999999999, 6: 5.864999999999999}
    seen = set()
    reentered = False
    for seed in range(40):
        s = m.MCTSSolver(n_iterations=3, depth=4, init_N=3, init_Q=11.73, estimate_value=0.0, rng=seed)
        plan(oracle, s, GW, [(1, 1)])
        t = oracle.tree_export(0)
        assert len(t["total_n"]) <= 4 and len(t["n"]) == 4 * len(t["total_n"])  # the notebook prints 4 state nodes x 4 actions
        for q, n in zip(t["q"], t["n"]):
            if n in stored:
                assert q == stored[int(n)], (seed, n, q)  # exact doubles
                seen.add(int(n))
        root_visits = int(t["total_n"][0]) - 12
        reentered |= root_visits > 3  # notebook: 6 root visits from 3 iterations (self-loops at the wall)
    assert seen == {3, 4, 5, 6}
    assert reentered


# ---- notebook cell 6: DPW backup with gamma = 0.95 -------------------------------------------------------
def manhattan_value(mdp, s, depth):
    d = abs(s[0] - 9) + abs(s[1] - 3)
    return 10.0 / d if d else float("inf")  # Julia: 10.0/0 == Inf


def test_notebook_cell6_dpw_backup_values(oracle):
    rows = {(tuple(r["s"]), r["a"]): r for r in VEC["dk_cell6"]["rows"]}
    set_values = {(a, b): v for a, b, v in VEC["dk_cell6"]["set_value"]}
    for (x, y), v in set_values.items():  # the estimator restated here prints the same values
        assert manhattan_value(None, (x, y), 0) == v
    det = m.SimpleGridWorld(tprob=1.0)
    special_q = lambda mdp, s, a: 11.73 if s == (1, 2) else 0.0
    special_n = lambda mdp, s, a: 3 if s == (1, 2) else 0
    mk = lambda: m.DPWSolver(n_iterations=1, depth=4, init_N=special_n, init_Q=special_q, estimate_value=manhattan_value, enable_action_pw=False)
    # s:[2,2] a:up -> new node [2,3], value 10/7: Q = 0 + 0.95*1.4285714285714286, N = 1
    plan(oracle, mk(), det, [(2, 2)])
    t = oracle.tree_export(0)
    assert t["a_labels"][0] == 0 and t["n"][0] == 1
    assert t["q"][0] == rows[((2, 2), "up")]["Q"] == 1.3571428571428572
    # s:[1,2] a:up -> new node [1,3], value 1.25: Q = 11.73 + (0.95*1.25 - 11.73)/4, N = 4
    plan(oracle, mk(), det, [(1, 2)])
    t = oracle.tree_export(0)
    assert t["n"][0] == 4 and t["q"][0] == rows[((1, 2), "up")]["Q"] == 9.094375
    assert list(t["n"][1:4]) == [3, 3, 3] and t["total_n"][0] == 13
    # s:[2,1] a:right -> [3,1], value 1.25: Q = 0.95*1.25 (reached with a constant-zero prior on the other actions)
    only_right = lambda mdp, s, a: 0.0 if a == "right" else -1.0
    s = m.DPWSolver(n_iterations=1, depth=4, init_Q=only_right, estimate_value=manhattan_value, enable_action_pw=False)
    plan(oracle, s, det, [(2, 1)])
    t = oracle.tree_export(0)
    assert t["q"][3] == rows[((2, 1), "right")]["Q"] == 1.1875 and t["n"][3] == 1


# ---- notebook Test_Visualization cell 3: the vanilla tree is a DAG keyed by state ------------------------
def _check_dag_invariants(t, root, n_iterations):
    for i in range(len(t["total_n"])):  # every state node: total_n == sum of its children's n (init_N = 0)
        ch = t["child_ids"][t["child_offsets"][i]:t["child_offsets"][i + 1]]
        assert t["total_n"][i] == t["n"][ch - 1].sum()
    assert tuple(t["s_labels"][0]) == tuple(root)  # tree.state_map[root] == 1 (test/runtests.jl:44)
    assert t["total_n"][0] > n_iterations  # the root is re-entered deeper inside a simulation
    assert len(t["total_n"]) <= 101
    assert len(np.unique(t["s_labels"], axis=0)) == len(t["s_labels"])  # one node per state


def test_visualization_notebook_dag_invariants(oracle):
    tv = VEC["tv_cell3"]
    assert sum(tv["root_children_N"]) == tv["root_N"]  # what the notebook shows
    s = m.MCTSSolver(n_iterations=tv["n_iterations"], depth=10, exploration_constant=10.0, rng=4)
    _, _, _, st = plan(oracle, s, GW, [tuple(tv["root"])])
    assert st.n_simulations == tv["n_iterations"]
    _check_dag_invariants(oracle.tree_export(0), tv["root"], tv["n_iterations"])


def test_visualization_notebook_statistics_reproduced(oracle):
    """The stored tree (root [1,3]: N = 264175 = 2.64 visits/iteration, action :left, per-child N and Q two
    levels deep) was produced by a release that returned 0.0 at depth 0 instead of calling estimate_value
    (the stored Q of about -0.05 is impossible with 50-step random rollouts at every depth-0 leaf, which give
    about -0.8). With estimate_value = 0.0 today's loop (src/vanilla.jl:247-250) has exactly that behaviour,
    apart from <= 101 new-node rollouts, and the oracle must then land on the stored numbers up to seed noise."""
    tv = VEC["tv_cell3"]
    nodes = tv["nodes"]
    n_seeds = 8
    acc = {k: {"N": [], "frac": [], "q": []} for k in nodes}
    actions = []
    for seed in range(n_seeds):
        s = m.MCTSSolver(n_iterations=tv["n_iterations"], depth=10, exploration_constant=10.0, rng=seed, estimate_value=0.0)
        a, _, _, _ = plan(oracle, s, GW, [tuple(tv["root"])])
        actions.append(GW.actions[int(a[0])])
        t = oracle.tree_export(0)
        _check_dag_invariants(t, tv["root"], tv["n_iterations"])
        labels = [tuple(x) for x in t["s_labels"].tolist()]
        for key in nodes:
            i = labels.index(tuple(map(int, key.split(","))))
            ch = t["child_ids"][t["child_offsets"][i]:t["child_offsets"][i + 1]]
            acc[key]["N"].append(t["total_n"][i])
            acc[key]["frac"].append(t["n"][ch - 1] / t["total_n"][i])
            acc[key]["q"].append(t["q"][ch - 1])
    assert actions.count(tv["action"]) >= n_seeds - 1  # stored decision: :left
    for key, ref in nodes.items():
        N = np.mean(acc[key]["N"])
        assert abs(N / ref["N"] - 1.0) < 0.08, (key, N, ref["N"])  # root: 2.52-2.61 vs 2.64 visits per iteration
        frac = np.mean(acc[key]["frac"], axis=0)
        ref_frac = np.array(ref["child_N"]) / ref["N"]
        sd = np.std(acc[key]["frac"], axis=0, ddof=1)  # the notebook is ONE run: allow 3 sigma of seed-to-seed spread
        assert (np.abs(frac - ref_frac) < 3 * sd + 0.01).all(), (key, frac, ref_frac, sd)
        q = np.mean(acc[key]["q"], axis=0)
        ref_q = np.array(ref["child_Q"])
        sdq = np.std(acc[key]["q"], axis=0, ddof=1)
        tol = 3 * sdq + 0.02  # stored values are rounded to 2 decimals; rarely visited children are noisy
        assert (np.abs(q - ref_q) < tol).all(), (key, q, ref_q, sdq)


# ---- SimpleGridWorld restatement vs the reference's hard-coded value-iteration table -------------------------
def gridworld_value_iteration(gw: m.SimpleGridWorld, iters=2000):
    """Independent numpy restatement of SimpleGridWorld (third implementation, for cross-checking)."""
    nx, ny = gw.size
    nS = nx * ny + 1
    T = np.zeros((nS, 4, nS))
    R = np.zeros(nS)
    dirs = [(0, 1), (0, -1), (-1, 0), (1, 0)]
    for i in range(nS - 1):
        s = gw.state_from_index(i)
        R[i] = gw.rewards.get(s, 0.0)
        for a in range(4):
            if s in gw.terminate_from:
                T[i, a, nS - 1] = 1.0
                continue
            for k, (dx, dy) in enumerate(dirs):
                p = gw.tprob if k == a else (1.0 - gw.tprob) / 3
                d = (s[0] + dx, s[1] + dy)
                if 1 <= d[0] <= nx and 1 <= d[1] <= ny:
                    T[i, a, gw.state_index(d)] += p
                else:
                    T[i, a, i] += p
    T[nS - 1, :, nS - 1] = 1.0
    V = np.zeros(nS)
    for _ in range(iters):
        Q = R[:, None] + gw.discount * T @ V
        Q[nS - 1] = 0.0
        V = Q.max(axis=1)
    return Q


def test_value_iteration_table_matches_reference(oracle):
    """test/value_test.jl:30-45 hard-codes VI-optimal actions for 16 states; an independent VI over the
    restated model must reproduce them, and one-step oracle gen() must agree with that model's support."""
    Q = gridworld_value_iteration(GW)
    exact = 0
    for key, act in VEC["value_test"]["vi_actions"].items():
        x, y = map(int, key.split(","))
        q = Q[GW.state_index((x, y))]
        # the table predates SimpleGridWorld (generated under `using Base.Test`); two of its 16 entries are
        # near-ties (dQ = 0.006 and 0.037) that a converged VI on today's model resolves the other way
        assert q[GW.action_index(act)] >= q.max() - 0.05, (key, act, q)
        exact += int(q.argmax() == GW.action_index(act))
    assert exact >= 14
    # oracle gen(): support and frequencies against the numpy model
    oracle.configure(GW.mdp_id, GW.params())
    for s, a in [((1, 3), 0), ((5, 5), 3), ((10, 10), 0), ((1, 1), 1)]:
        counts = {}
        for it in range(4000):
            sp, r, used = oracle.gen(GW.mdp_id, s, a, seed=99, tree=0, iteration=it)
            counts[tuple(sp)] = counts.get(tuple(sp), 0) + 1
            assert r == 0.0 and used == 1
        dirs = [(0, 1), (0, -1), (-1, 0), (1, 0)]
        exp = {}
        for k, (dx, dy) in enumerate(dirs):
            d = (s[0] + dx, s[1] + dy)
            p = 0.7 if k == a else 0.1
            tgt = d if (1 <= d[0] <= 10 and 1 <= d[1] <= 10) else s
            exp[tgt] = exp.get(tgt, 0.0) + p
        assert set(counts) == set(exp)
        for k, p in exp.items():
            assert abs(counts[k] / 4000 - p) < 4 * np.sqrt(p * (1 - p) / 4000) + 1e-9
    # reward is collected when acting FROM a reward cell, which then moves to the terminal state without a draw
    sp, r, used = oracle.gen(GW.mdp_id, (9, 3), 0)
    assert tuple(sp) == (-1, -1) and r == 10.0 and used == 0
    sp, r, used = oracle.gen(GW.mdp_id, (-1, -1), 2)
    assert tuple(sp) == (-1, -1) and r == 0.0 and used == 0


def test_value_test_statistical_bound(oracle):
    """test/value_test.jl:47-65: mean |Q(s, a_mcts) - Q(s, a_VI)| < 1.0 at 10,000 iterations, c = 20."""
    states, vi_a = [], []
    for key, act in VEC["value_test"]["vi_actions"].items():
        states.append(tuple(map(int, key.split(","))))
        vi_a.append(GW.action_index(act))
    s = m.MCTSSolver(n_iterations=10_000, depth=20, exploration_constant=20.0, rng=43)
    actions, best_q, _, _ = plan(oracle, s, GW, states, n_threads=8)
    diffs = []
    for i in range(len(states)):
        rs = oracle.root_stats(i)
        q_by_action = dict(zip(rs["action_idx"].tolist(), rs["q"].tolist()))
        assert best_q[i] == q_by_action[int(actions[i])] == max(q_by_action.values())  # decision = argmax Q
        diffs.append(abs(q_by_action[int(actions[i])] - q_by_action[vi_a[i]]))
    assert np.mean(diffs) < VEC["value_test"]["bound_mean_abs_dq"]


# ---- rollouts: closed forms --------------------------------------------------------------------------
def test_rollout_known_answers(oracle):
    det = m.SimpleGridWorld(tprob=1.0)
    oracle.configure(det.mdp_id, det.params())
    cfg, _ = m.build_cfg(m.MCTSSolver(estimate_value=m.RolloutEstimator(m.SeekTarget((9, 3)), max_depth=50)), det)
    v, steps = oracle.rollout(cfg, det.mdp_id, (5, 1))
    expect = 0.0
    disc = 1.0
    for _ in range(6):  # right x4, up x2: six zero-reward moves
        expect += disc * 0.0
        disc *= 0.95
    expect += disc * 10.0  # then acting from (9,3) collects +10 and terminates
    assert v == expect and steps == 7
    cfg, _ = m.build_cfg(m.MCTSSolver(estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"), max_depth=50)), det)
    v, steps = oracle.rollout(cfg, det.mdp_id, (1, 1))
    assert v == 0.0 and steps == 50  # walks into the wall until max_depth
    v, steps = oracle.rollout(cfg, det.mdp_id, (4, 1))  # up, up -> (4,3) reward -10
    assert v == 0.95 * 0.95 * -10.0 and steps == 3
    cfg, _ = m.build_cfg(m.MCTSSolver(estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"), max_depth=-1)), det)
    assert oracle.rollout(cfg, det.mdp_id, (1, 1), remaining_depth=0) == (0.0, 0)
    assert oracle.rollout(cfg, det.mdp_id, (4, 2), remaining_depth=2)[1] == 2
    assert oracle.rollout(cfg, det.mdp_id, (-1, -1), remaining_depth=5) == (0.0, 0)  # terminal: loop guard
    # eps: stop once disc <= eps
    cfg, _ = m.build_cfg(m.MCTSSolver(estimate_value=m.RolloutEstimator(m.ConstantPolicy("up"), max_depth=50, eps=0.9)), det)
    assert oracle.rollout(cfg, det.mdp_id, (1, 1))[1] == 3  # disc: 1, .95, .9025 run; .857 stops


def test_mountaincar_restatement(oracle):
    mc = m.MountainCar()
    oracle.configure(mc.mdp_id, mc.params())
    # initial state (-0.5, 0), a = +1 (index 2): v' = 0.001 + cos(-1.5)*-0.0025
    sp, r, used = oracle.gen(mc.mdp_id, (-0.5, 0.0), 2)
    v = (0.0 + 1.0 * 0.001) + float(np.cos(-1.5)) * -0.0025
    assert abs(sp[1] - v) < 1e-17 and abs(sp[0] - (-0.5 + v)) < 1e-16 and r == -1.0 and used == 0
    sp, r, _ = oracle.gen(mc.mdp_id, (-1.19, -0.07), 0)  # inelastic wall
    assert tuple(sp) == (-1.2, 0.0) and r == -1.0
    sp, r, _ = oracle.gen(mc.mdp_id, (0.45, 0.07), 2)  # reaches the goal: jackpot on the transition INTO a terminal state
    assert sp[0] >= 0.5 and r == 100.0
    sp, r, _ = oracle.gen(mc.mdp_id, (-0.6, 0.0695), 2)  # speed clamp (cos(-1.8) < 0 adds to v)
    assert sp[1] == 0.07


def test_decision_is_argmax_q_first_max(oracle):
    """best_sanode_Q: strict '>' from -Inf, ties -> first child (src/vanilla.jl:339-349)."""
    s = m.MCTSSolver(n_iterations=0, depth=5, init_Q=2.5)
    a, q, _, _ = plan(oracle, s, GW, [(3, 3)])
    assert a[0] == 0 and q[0] == 2.5
    a, q, st, _ = plan(oracle, m.DPWSolver(n_iterations=5, depth=5), GW, [(3, 3), (-1, -1)])
    assert st[1] == m.abi.ERR_TERMINAL_ROOT and st[0] == 0


def test_libm_oracle_takes_the_same_decisions(oracle, oracle_libm):
    """The deterministic log/cos only matter for last-ulp ties: with glibc's functions the chosen actions
    are the same on a deterministic config-2 sample."""
    det = m.SimpleGridWorld(tprob=1.0)
    s = m.MCTSSolver(n_iterations=500, depth=20, exploration_constant=5.0, estimate_value=m.RolloutEstimator(m.ConstantPolicy("up")))
    roots = [(1 + i % 10, 1 + (i // 10) % 10) for i in range(100)]
    a1, q1, _, _ = plan(oracle, s, det, roots, n_threads=4)
    a2, q2, _, _ = plan(oracle_libm, s, det, roots, n_threads=4)
    assert (a1 == a2).mean() >= 0.97
    assert np.allclose(q1, q2, rtol=0, atol=0.5)


def test_literal_and_multiset_transition_sampling_agree(oracle):
    """src/dpw.jl:131,140 keeps every (spnode, r) push and samples by index; the multiset form samples the
    same distribution. Identical trees when nothing is sampled; statistically equal root Q otherwise."""
    # init_N = 1 keeps n >= 1, so n_a_children (<= 5) <= 10*sqrt(n) always holds: every visit widens, nothing is sampled
    s = m.DPWSolver(n_iterations=300, depth=10, rng=1, init_N=1)
    plan(oracle, s, GW, [(2, 2)], literal_transitions=True)
    t1 = oracle.tree_export(0)
    _, _, _, st = plan(oracle, s, GW, [(2, 2)], literal_transitions=False)
    t2 = oracle.tree_export(0)
    assert st.n_transition_samples == 0
    for k in t1:
        assert np.array_equal(t1[k], t2[k])
    qs = {True: [], False: []}
    n_samples = 0
    for lit in (True, False):
        for seed in range(60):
            s = m.DPWSolver(n_iterations=300, depth=10, k_state=1.0, alpha_state=0.1, rng=seed)
            _, q, _, st = plan(oracle, s, GW, [(2, 2)], literal_transitions=lit)
            qs[lit].append(q[0])
            n_samples += st.n_transition_samples
    assert n_samples > 1000
    a, b = np.array(qs[True]), np.array(qs[False])
    assert abs(a.mean() - b.mean()) < 3 * np.sqrt(a.var() / len(a) + b.var() / len(b)) + 1e-12
