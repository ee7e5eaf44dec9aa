"""The BASELINE.json configurations as synthetic workloads: root-state formulas, models and solver settings, shared by bench.py,
the scripts and the tests (SURVEY.md 8d "Config 1..5").

  config 2   SimpleGridWorld vanilla UCT, 65 536 roots, n_iterations=1000, depth=20, c=5, deterministic rollout policy
             FunctionPolicy(s -> :up) (test/options.jl:29), rollout max_depth=50. The headline is the DEFAULT SimpleGridWorld
             (tprob = 0.7: stochastic transitions); the tprob = 1.0 variant (fully deterministic, lock-step trees) and the
             random-rollout variant are named secondaries.
  config 3   bench/non-terminal_gw_dpw.jl solver, 16 384 roots
  config 4   MountainCar DPWSolver(n_iterations=10_000, depth=50), 8 192 roots in all (1 024 per GPU on 8 GPUs)
  config 5   root-parallel UCT of ONE GridWorld root, 10^7 simulations per decision = 9 472 trees x 1 056 iterations
"""
from __future__ import annotations

import numpy as np

HEADLINE = "config2-uct-gridworld-default"

WORKLOADS = {
    # name: mdp, mdp kwargs, solver kind, rollout policy, solver kwargs, roots (per GPU for the weak-scaled headline line)
    "config2-uct-gridworld-default": dict(mdp="gw", mdp_kw=dict(tprob=0.7), kind="uct", policy="up",
                                          solver_kw=dict(n_iterations=1000, depth=20, exploration_constant=5.0), n_roots=65536),
    "config2-uct-gridworld-deterministic": dict(mdp="gw", mdp_kw=dict(tprob=1.0), kind="uct", policy="up",
                                                solver_kw=dict(n_iterations=1000, depth=20, exploration_constant=5.0), n_roots=65536),
    "config2-uct-gridworld-stochastic": dict(mdp="gw", mdp_kw=dict(tprob=0.7), kind="uct", policy="random",
                                             solver_kw=dict(n_iterations=1000, depth=20, exploration_constant=5.0), n_roots=65536),
    "config3-dpw-gridworld": dict(mdp="gw", mdp_kw=dict(tprob=0.7), kind="dpw", policy="random",
                                  solver_kw=dict(n_iterations=1000, depth=20, exploration_constant=10.0, k_state=4.0, alpha_state=1 / 8,
                                                 k_action=4.0, alpha_action=1 / 8), n_roots=16384),
    "config4-dpw-mountaincar": dict(mdp="mc", mdp_kw=dict(), kind="dpw", policy="random", solver_kw=dict(n_iterations=10000, depth=50),
                                    n_roots=1024),
    # (not a BASELINE config) vanilla UCT on MountainCar: the tile kernel's successor cache for single-successor hashed models
    "uct-mountaincar": dict(mdp="mc", mdp_kw=dict(), kind="uct", policy="random",
                            solver_kw=dict(n_iterations=10000, depth=50, exploration_constant=10.0), n_roots=1024),
    # config 4 with MountainCar registered at run time as a plugin (mctsb200_mdp_register: the NVRTC-compiled tile kernel), and
    # its noisy variant (continuous-state stochastic transitions: state widening really widens)
    "config4-dpw-mountaincar-plugin": dict(mdp="mc-plugin", mdp_kw=dict(), kind="dpw", policy="random",
                                           solver_kw=dict(n_iterations=10000, depth=50), n_roots=1024),
    "config4-dpw-noisy-mountaincar-plugin": dict(mdp="noisy-plugin", mdp_kw=dict(), kind="dpw", policy="random",
                                                 solver_kw=dict(n_iterations=10000, depth=50), n_roots=1024),
    # (not a BASELINE config) DPW over a continuous action space (test/discussion_514.jl's model with drift and cost): hundreds of action
    # nodes under a state node, the case where a tile of lanes per tree pays (children scans)
    "box-dpw-continuous-actions": dict(mdp="box", mdp_kw=dict(drift=0.3, cost=0.5), kind="dpw", policy="random",
                                       solver_kw=dict(n_iterations=4000, depth=6, exploration_constant=8.0), n_roots=4096),
    # config 5's trees: one GridWorld root, default model, random rollouts (the README solver)
    "config5-root-parallel-uct-gridworld": dict(mdp="gw", mdp_kw=dict(tprob=0.7), kind="uct", policy="random",
                                                solver_kw=dict(n_iterations=1056, depth=20, exploration_constant=5.0), n_roots=9472),
}

CONFIG4_TOTAL_ROOTS = 8192
CONFIG5_ROOT = (1, 1)
CONFIG5_TREES = 9472        # 8 x 1 184: the same ensemble whatever the number of GPUs
CONFIG5_ITERATIONS = 1056   # 9 472 x 1 056 = 10 002 432 simulations per decision


def gridworld_roots(n: int):
    """BASELINE config 2/3 root formula: root i = GWPos(1 + i mod 10, 1 + (i div 10) mod 10)."""
    return [(1 + i % 10, 1 + (i // 10) % 10) for i in range(n)]


def mountaincar_roots(n: int, seed: int = 0):
    """Config 4 roots: x = -1.2 + 1.7 u1 (rejecting x >= 0.5), v = -0.07 + 0.14 u2."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        x, v = -1.2 + 1.7 * rng.random(), -0.07 + 0.14 * rng.random()
        if x < 0.5:
            out.append((x, v))
    return out


def make_workload(name, n_roots=None, lanes=0, n_iterations=None, same_root=None):
    """-> (mdp, solver, roots as the C-contiguous array the ABI takes)."""
    from . import (ConstantPolicy, DPWSolver, GaussianBox, MCTSSolver, MountainCar, MountainCarPlugin, NoisyMountainCar, RandomSolver,
                   RolloutEstimator, SimpleGridWorld)

    w = WORKLOADS[name]
    n = n_roots or w["n_roots"]
    mdp = {"gw": SimpleGridWorld, "mc": MountainCar, "mc-plugin": MountainCarPlugin, "noisy-plugin": NoisyMountainCar, "box": GaussianBox}[w["mdp"]](**w["mdp_kw"])
    pol = {"up": ConstantPolicy("up"), "random": RandomSolver()}[w["policy"]]
    kw = dict(w["solver_kw"])
    if n_iterations:
        kw["n_iterations"] = n_iterations
    kw["estimate_value"] = RolloutEstimator(pol, max_depth=50)
    kw["lanes_per_tree"] = lanes
    solver = (MCTSSolver if w["kind"] == "uct" else DPWSolver)(rng=0, **kw)
    if w["mdp"] == "box":
        rng = np.random.default_rng(0)
        roots = [tuple(float(x) for x in r) for r in rng.uniform(-4.0, 4.0, (n, 3))]
    else:
        roots = gridworld_roots(n) if w["mdp"] == "gw" else mountaincar_roots(n)
    if same_root:  # every tree plans from the same root state
        roots = [tuple(int(v) for v in same_root.split(","))] * n if isinstance(same_root, str) else [tuple(same_root)] * n
    return mdp, solver, mdp.encode_states(roots)
