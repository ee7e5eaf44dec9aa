"""Small plans through every kernel family, meant to be run under `compute-sanitizer --tool memcheck` (scripts/gpu_sanitize.sh):
the lane-per-tree UCT and DPW kernels, the tile kernel (built-in MountainCar, every tile width), a run-time plugin (NVRTC-compiled
tile kernel, unbounded branching), kept trees, the episode driver, tree export."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mcts_jl_b200 as m
from mcts_jl_b200.workloads import gridworld_roots, mountaincar_roots

ctx = m.Context(0)
gw, mc, noisy = m.SimpleGridWorld(), m.MountainCar(), m.NoisyMountainCar()


def plan(solver, mdp, roots, lanes=0, export=True):
    cfg, keep = m.build_cfg(solver, mdp)
    cfg.lanes_per_tree = lanes
    ctx.configure(mdp.mdp_id, mdp.params())
    a, q, s, st = ctx.plan(cfg, mdp.mdp_id, mdp.encode_states(roots), 0, want_status=True)
    if export:
        for i in (0, len(roots) - 1):
            ctx.tree_export(i, mdp.state_dtype, mdp.state_width)
            ctx.root_stats(i)
    return st


print("uct fast", plan(m.MCTSSolver(n_iterations=60, depth=20, exploration_constant=5.0, rng=1), gw, gridworld_roots(70)).kernel_variant)
print("dpw fast", plan(m.DPWSolver(n_iterations=60, depth=15, rng=2), gw, gridworld_roots(70)).kernel_variant)
for lanes in (1, 2, 4, 8, 16, 32):
    plan(m.DPWSolver(n_iterations=80, depth=30, rng=3), mc, mountaincar_roots(9), lanes)
    plan(m.MCTSSolver(n_iterations=60, depth=20, exploration_constant=3.0, rng=4), mc, mountaincar_roots(9), lanes)
    plan(m.DPWSolver(n_iterations=60, depth=10, rng=5), gw, gridworld_roots(9), lanes)  # tile kernel on the dense-state model
print("tile widths done")
for lanes in (0, 1, 8):
    plan(m.DPWSolver(n_iterations=120, depth=20, rng=6), noisy, mountaincar_roots(9), lanes)
    plan(m.DPWSolver(n_iterations=80, depth=20, rng=7, check_repeat_state=False), noisy, mountaincar_roots(9), lanes)
    plan(m.MCTSSolver(n_iterations=80, depth=20, exploration_constant=3.0, rng=8), noisy, mountaincar_roots(9), lanes)
print("plugin done")
p = m.solve(m.MCTSSolver(n_iterations=40, depth=10, exploration_constant=5.0, reuse_tree=True, rng=9), gw, ctx=ctx)
for _ in range(3):
    p.action_batch(gridworld_roots(40))
p.run_episodes(gridworld_roots(40), 4, 3)
m.solve(m.DPWSolver(n_iterations=40, depth=10, rng=10), noisy, ctx=ctx).run_episodes(mountaincar_roots(12), 3, 3)
print("episodes done")
ctx.close()
print("sanitize_small OK")
