"""Small plans through every kernel family: on the GPU box `python scripts/sanitize_small.py` (under compute-sanitizer where the pool
allows it), here `python scripts/sanitize_small.py --emu <host-emulation library>` under AddressSanitizer (scripts/asan_emu.sh; run-time
plugins are device-only and skipped there):
the lane-per-tree UCT and DPW kernels, the tile kernel (built-in MountainCar, every tile width), a run-time plugin (NVRTC-compiled
tile kernel, unbounded branching), continuous action spaces (built-in model and plugin; kept trees until their room is used up), kept trees of hashed
models, _vis_stats, the episode driver, tree export."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mcts_jl_b200 as m
from mcts_jl_b200.workloads import gridworld_roots, mountaincar_roots

EMU = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--emu" else None
ctx = m.Context(0, lib_path=EMU) if EMU else m.Context(0)
gw, mc = m.SimpleGridWorld(), m.MountainCar()
noisy = None if EMU else m.NoisyMountainCar()


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
for lanes in (0, 1, 8) if noisy else ():
    plan(m.DPWSolver(n_iterations=120, depth=20, rng=6), noisy, mountaincar_roots(9), lanes)
    plan(m.DPWSolver(n_iterations=80, depth=20, rng=7, check_repeat_state=False), noisy, mountaincar_roots(9), lanes)
    plan(m.MCTSSolver(n_iterations=80, depth=20, exploration_constant=3.0, rng=8), noisy, mountaincar_roots(9), lanes)
print("plugin done")
p = m.solve(m.MCTSSolver(n_iterations=40, depth=10, exploration_constant=5.0, reuse_tree=True, rng=9), gw, ctx=ctx)
for _ in range(3):
    p.action_batch(gridworld_roots(40))
p.run_episodes(gridworld_roots(40), 4, 3)
m.solve(m.DPWSolver(n_iterations=40, depth=10, rng=10), noisy or mc, ctx=ctx).run_episodes(mountaincar_roots(12), 3, 3)
p.action_root_parallel((1, 1), 37)
print("episodes done")
# continuous action spaces: action-node table + arena, every tile width, kept trees up to MCTSB200_ERR_CAPACITY, episodes
box, cp = m.GaussianBox(drift=0.2, cost=1.0, var=(0.05, 0.1, 0.02)), None if EMU else m.ContinuousPendulum()
box_roots = [(0.0, 0.0, 0.0), (1.5, -2.0, 0.5), (-3.0, 3.0, -1.0)]
for lanes in (0, 1, 4, 32):
    plan(m.DPWSolver(n_iterations=150, depth=8, rng=11), box, box_roots, lanes)
    plan(m.DPWSolver(n_iterations=150, depth=8, rng=12, check_repeat_action=False, check_repeat_state=False), box, box_roots, lanes)
    if cp:
        plan(m.DPWSolver(n_iterations=150, depth=8, rng=13, estimate_value=m.RolloutEstimator(m.PluginPolicy(100), max_depth=12)), cp, [(0.02, 0.0), (-0.1, 0.05)], lanes)
for mdp, roots in ((box, box_roots), (cp, [(0.02, 0.0), (-0.1, 0.05)])):
    if mdp is None:
        continue
    ctx.clear_trees()
    try:
        for call in range(6):
            plan(m.DPWSolver(n_iterations=60, depth=6, rng=20 + call, keep_tree=True, keep_tree_calls=3), mdp, roots)
        raise SystemExit("kept trees never ran out of room")
    except m.MCTSB200Error as e:
        assert e.status == m.abi.ERR_CAPACITY, e
    ctx.clear_trees()
    m.solve(m.DPWSolver(n_iterations=40, depth=6, rng=30, keep_tree=True), mdp, ctx=ctx).run_episodes(roots, 4, 3)
print("continuous actions done")
# kept trees of a hashed model, _vis_stats
ctx.clear_trees()
for call in range(3):
    plan(m.DPWSolver(n_iterations=60, depth=12, rng=40 + call, keep_tree=True), mc, mountaincar_roots(6))
ctx.clear_trees()
plan(m.MCTSSolver(n_iterations=80, depth=10, exploration_constant=5.0, rng=50, enable_tree_vis=True), gw, gridworld_roots(12))
ctx.tree_vis_stats(3)
plan(m.MCTSSolver(n_iterations=80, depth=10, exploration_constant=3.0, rng=51, enable_tree_vis=True), mc, mountaincar_roots(5))
ctx.tree_vis_stats(2)
print("kept hashed trees / vis stats done")
ctx.close()
print("sanitize_small OK")
