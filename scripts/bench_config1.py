"""BASELINE config 1 -- the reference's own README / bench/non-terminal_gw.jl case: SimpleGridWorld, MCTSSolver(n_iterations=50,
depth=20, exploration_constant=5.0), ONE root state. A latency measurement, not a throughput one: time per `action(planner, s)`
through the Python mirror and through the bare C-ABI call (host buffers in, host buffers out), next to the CPU oracle on one
thread (the reference is single-threaded). Also the same decision for a batch of 1 000 seeds (one tree per seed) in one call.
Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import mcts_jl_b200 as m
from oracle.oracle_binding import Oracle

REPS = int(os.environ.get("REPS", "300"))
gw = m.SimpleGridWorld()
solver = m.MCTSSolver(n_iterations=50, depth=20, exploration_constant=5.0, rng=0)
root = (1, 2)
ctx = m.Context(0)
planner = m.solve(solver, gw, ctx=ctx)
for _ in range(20):
    planner.action(root)
t = []
for _ in range(REPS):
    t0 = time.perf_counter()
    a = planner.action(root)
    t.append(time.perf_counter() - t0)
api_us = np.array(t) * 1e6

cfg, keep = m.build_cfg(solver, gw)
r = gw.encode_states([root])
t, kms = [], []
for i in range(REPS):
    cfg.seed = i
    t0 = time.perf_counter()
    _, _, _, st = ctx.plan(cfg, gw.mdp_id, r)
    t.append(time.perf_counter() - t0)
    kms.append(st.kernel_ms)
abi_us = np.array(t) * 1e6

orc = Oracle()
orc.configure(gw.mdp_id, gw.params())
t = []
for i in range(REPS):
    cfg.seed = i
    t0 = time.perf_counter()
    orc.plan(cfg, gw.mdp_id, r, n_threads=1, keep_trees=False)
    t.append(time.perf_counter() - t0)
cpu_us = np.array(t) * 1e6

# 1 000 seeds of the same decision in one call (tree i uses Philox stream (seed, i, ...))
roots = gw.encode_states([root] * 1000)
cfg.seed = 0
ctx.plan(cfg, gw.mdp_id, roots)
t0 = time.perf_counter()
_, _, _, stb = ctx.plan(cfg, gw.mdp_id, roots)
batch_s = time.perf_counter() - t0
t0 = time.perf_counter()
orc.plan(cfg, gw.mdp_id, roots, n_threads=1, keep_trees=False)
cpu_batch_s = time.perf_counter() - t0
print(json.dumps({
    "config": "config1-uct-gridworld-single-root", "n_iterations": 50, "depth": 20, "exploration_constant": 5.0, "reps": REPS,
    "action_call_us": {"median": float(np.median(api_us)), "p10": float(np.percentile(api_us, 10)), "p90": float(np.percentile(api_us, 90))},
    "c_abi_plan_us": {"median": float(np.median(abi_us)), "p10": float(np.percentile(abi_us, 10)), "p90": float(np.percentile(abi_us, 90))},
    "search_kernel_us_median": float(np.median(kms) * 1e3), "lanes_per_tree": int(st.lanes_per_tree), "kernel_variant": int(st.kernel_variant),
    "cpu_oracle_1thread_us": {"median": float(np.median(cpu_us)), "p10": float(np.percentile(cpu_us, 10)), "p90": float(np.percentile(cpu_us, 90))},
    "batch_1000_seeds": {"gpu_ms": batch_s * 1e3, "gpu_kernel_ms": float(stb.kernel_ms), "cpu_oracle_1thread_ms": cpu_batch_s * 1e3,
                         "simulations": int(stb.n_simulations)},
}))
