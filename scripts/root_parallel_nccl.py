"""BASELINE config 5 on N GPUs (torchrun): root-parallel UCT of ONE root, 10^7 simulations per decision split over the
ranks, per-rank [sum N | sum NQ] merged with one NCCL allreduce. Prints one JSON line from rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import mcts_jl_b200 as m
from mcts_jl_b200 import distributed as D

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local)
gw = m.SimpleGridWorld()
total_sims = int(float(os.environ.get("TOTAL_SIMS", "1e7")))
# ENSEMBLE=fixed (default): a FIXED ensemble (8 GPUs x 1 184 trees, ~1 056 iterations each), sharded by global
#   tree index: the merged sums -- and the decision -- do not depend on the number of GPUs. One tree's 1 056 iterations are a
#   serial chain, so the time per decision does not shrink with more GPUs either.
# ENSEMBLE=scaled: 9 472 trees PER GPU and 10^7 / trees iterations each: the same number of simulations per decision, the
#   serial chain of every tree N times shorter (a different, statistically equivalent ensemble for every N).
mode = os.environ.get("ENSEMBLE", "fixed")
n_trees = int(os.environ.get("TREES", str(8 * 148 * 8))) * (world if mode == "scaled" else 1)
n_iter = max(1, total_sims // n_trees)
solver = m.MCTSSolver(n_iterations=n_iter, depth=20, exploration_constant=5.0, rng=5)
for rep in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, best, q, sum_n, stats = D.plan_root_parallel(ctx, solver, gw, (1, 1), n_trees)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"config": "config5-root-parallel", "ensemble": mode, "n_gpus": world, "trees": n_trees, "iterations_per_tree": n_iter,
                      "simulations": int(n_trees * n_iter), "seconds_per_decision": dt, "simulations_per_s": n_trees * n_iter / dt,
                      "action": gw.actions[a], "best_q": best, "q": [float(x) for x in q], "sum_n": [float(x) for x in sum_n],
                      "kernel_ms_rank0": float(stats.kernel_ms)}))
if world > 1:
    dist.destroy_process_group()
