"""Shared helpers for the parity tests: drive the library and the oracle with identical bytes and
compare actions, Q, per-root status, work counters and whole exported trees bit-for-bit."""
import numpy as np

import mcts_jl_b200 as m

COUNTER_KEYS = ["n_simulations", "n_tree_levels", "n_dpw_children_seen", "n_state_inserts", "n_action_inserts", "n_rollouts",
                "n_rollout_steps", "n_transition_samples", "algorithmic_bytes"]


def oracle_id(mdp) -> int:
    """The oracle's id of a model (a run-time plugin's library id is assigned at registration; the oracle restates the shipped
    plugins under fixed ids)."""
    oid = getattr(mdp, "oracle_mdp_id", None)
    return mdp.mdp_id if oid is None else oid


def bits(a: np.ndarray) -> np.ndarray:
    return a.view(np.int64) if a.dtype == np.float64 else a


def assert_trees_equal(t1: dict, t2: dict, what: str = "") -> None:
    assert set(t1) == set(t2)
    for k in t1:
        assert t1[k].shape == t2[k].shape, f"{what}: field {k} shape {t1[k].shape} != {t2[k].shape}"
        assert np.array_equal(bits(t1[k]), bits(t2[k])), f"{what}: field {k} differs"


from mcts_jl_b200.workloads import gridworld_roots, mountaincar_roots  # noqa: E402,F401  (the BASELINE root formulas live in the package)


def run_both(ctx, orc, solver, mdp, roots, base: int = 0, compare_trees="all", lanes=None, literal=False):
    """Plan with the library (ctx) and with the oracle; assert bit-exact agreement; return library outputs."""
    cfg, keep = m.build_cfg(solver, mdp)
    if lanes is not None:
        cfg.lanes_per_tree = lanes
    ctx.configure(mdp.mdp_id, mdp.params())
    orc.configure(oracle_id(mdp), mdp.params())
    r = mdp.encode_states(roots)
    a1, q1, s1, st1 = ctx.plan(cfg, mdp.mdp_id, r, base, want_status=True)
    a2, q2, s2, st2 = orc.plan(cfg, oracle_id(mdp), r, base, literal_transitions=literal, n_threads=4)
    assert np.array_equal(s1, s2), "per-root status differs"
    assert np.array_equal(a1, a2), f"actions differ at {np.nonzero(a1 != a2)[0][:10]}"
    assert np.array_equal(bits(q1), bits(q2)), "best_Q differs"
    d1, d2 = st1.as_dict(), st2.as_dict()
    for k in COUNTER_KEYS:
        assert d1[k] == d2[k], f"counter {k}: {d1[k]} != {d2[k]}"
    idx = range(len(roots)) if compare_trees == "all" else compare_trees
    for i in idx:
        t1 = ctx.tree_export(i, mdp.state_dtype, mdp.state_width)
        t2 = orc.tree_export(i, mdp.state_dtype, mdp.state_width)
        assert_trees_equal(t1, t2, f"tree {i}")
        rs1, rs2 = ctx.root_stats(i), orc.root_stats(i)
        for k in ("action_idx", "n", "q"):
            assert np.array_equal(bits(rs1[k]), bits(rs2[k]))
        assert rs1["total_n"] == rs2["total_n"]
    del keep
    return a1, q1, s1, st1


def run_episodes_both(ctx, orc, solver, mdp, starts, max_steps: int, env_seed: int = 0, base: int = 0):
    """The closed-loop episode driver (plan at every step, then gen) on the library and on the oracle, bit-for-bit."""
    cfg, keep = m.build_cfg(solver, mdp)
    ctx.configure(mdp.mdp_id, mdp.params())
    orc.configure(oracle_id(mdp), mdp.params())
    ctx.clear_trees()
    orc.clear_trees()
    s0 = mdp.encode_states(starts)
    r1, n1, f1, st1 = ctx.run_episodes(cfg, mdp.mdp_id, s0, max_steps, env_seed, base)
    r2, n2, f2, st2 = orc.run_episodes(cfg, oracle_id(mdp), s0, max_steps, env_seed, base, n_threads=4)
    assert np.array_equal(n1, n2), f"episode lengths differ at {np.nonzero(n1 != n2)[0][:10]}"
    assert np.array_equal(bits(r1), bits(r2)), f"returns differ at {np.nonzero(bits(r1) != bits(r2))[0][:10]}"
    assert np.array_equal(f1.view(np.uint8), f2.view(np.uint8)), "final states differ"
    d1, d2 = st1.as_dict(), st2.as_dict()
    for k in COUNTER_KEYS:
        assert d1[k] == d2[k], f"counter {k}: {d1[k]} != {d2[k]}"
    del keep
    return r1, n1, f1, st1
