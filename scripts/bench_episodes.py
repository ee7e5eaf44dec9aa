#!/usr/bin/env python
"""Closed-loop episode benchmark: the loop bench/non-terminal_gw.jl and bench/non-terminal_gw_dpw.jl time
(simulate(RolloutSimulator(max_steps=100), SimpleGridWorld(), planner) repeated N times), run as ONE batch of
episodes through mctsb200_run_episodes, with the CPU oracle timed on a sample beside it.

    python scripts/bench_episodes.py [--kind uct|dpw] [--episodes 65536] [--max-steps 100] [--reuse]

Prints one JSON line: episodes/s, environment steps/s and simulations/s of the whole batch (host buffers in, host
buffers out, everything between on the device). Not the driver's bench (bench.py is); a measurement of SURVEY §8(f).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="uct", choices=["uct", "dpw"])
    ap.add_argument("--episodes", type=int, default=65536)
    ap.add_argument("--max-steps", type=int, default=100)
    ap.add_argument("--reuse", action="store_true", help="reuse_tree / keep_tree: the trees follow their episodes")
    ap.add_argument("--cpu-episodes", type=int, default=0, help="oracle sample size (0: skip the CPU leg)")
    ap.add_argument("--check", action="store_true", help="compare the CPU sample's returns with the device's bit-for-bit")
    args = ap.parse_args()

    import mcts_jl_b200 as m
    from mcts_jl_b200.workloads import gridworld_roots

    mdp = m.SimpleGridWorld()
    if args.kind == "uct":  # bench/non-terminal_gw.jl:13-15
        solver = m.MCTSSolver(depth=20, n_iterations=100, exploration_constant=10.0, rng=8, reuse_tree=args.reuse)
    else:  # bench/non-terminal_gw_dpw.jl:13-22
        solver = m.DPWSolver(depth=20, n_iterations=1000, exploration_constant=10.0, k_state=4.0, alpha_state=1 / 8, k_action=4.0,
                             alpha_action=1 / 8, rng=8, keep_tree=args.reuse)
    starts = [p for p in gridworld_roots(100) if not mdp.isterminal(p)]
    starts = [starts[i % len(starts)] for i in range(args.episodes)]
    cfg, keep = m.build_cfg(solver, mdp)
    s0 = mdp.encode_states(starts)

    ctx = m.Context(0)
    ctx.configure(mdp.mdp_id, mdp.params())
    ctx.run_episodes(cfg, mdp.mdp_id, s0[:1024], 3, 7)  # warm-up: module load, buffers
    ctx.clear_trees()
    t0 = time.perf_counter()
    ret, steps, final, st = ctx.run_episodes(cfg, mdp.mdp_id, s0, args.max_steps, 7)
    dt = time.perf_counter() - t0
    line = {
        "benchmark": f"episodes-{args.kind}-gridworld" + ("-reuse" if args.reuse else ""), "episodes": args.episodes, "max_steps": args.max_steps,
        "n_iterations": int(cfg.n_iterations), "seconds": dt, "episodes_per_s": args.episodes / dt, "env_steps_per_s": float(steps.sum()) / dt,
        "simulations_per_s": st.n_simulations / dt, "mean_return": float(ret.mean()), "mean_steps": float(steps.mean()),
        "kernel_ms": st.kernel_ms, "kernel_launches": int(st.n_kernel_launches), "kernel_variant": int(st.kernel_variant),
    }
    if args.cpu_episodes:
        from oracle.oracle_binding import Oracle

        orc = Oracle()
        orc.configure(mdp.mdp_id, mdp.params())
        threads = os.cpu_count() or 1
        n = args.cpu_episodes
        t0 = time.perf_counter()
        r2, n2, f2, st2 = orc.run_episodes(cfg, mdp.mdp_id, s0[:n], args.max_steps, 7, n_threads=threads)
        dtc = time.perf_counter() - t0
        line["cpu"] = {"episodes": n, "threads": threads, "seconds": dtc, "episodes_per_s": n / dtc, "simulations_per_s": st2.n_simulations / dtc}
        if args.check:
            line["cpu"]["bit_exact"] = bool(np.array_equal(ret[:n].view(np.int64), r2.view(np.int64)) and np.array_equal(steps[:n], n2))
    del keep
    print(json.dumps(line))


if __name__ == "__main__":
    main()
