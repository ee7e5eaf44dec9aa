"""Multi-GPU plumbing: one process per GPU (torch.distributed), no data-path collective for batched
planning, one allreduce per decision for root-parallel search.

  * Batched planning (BASELINE configs 2-4): every root state's tree is independent (the reference plans one
    state per call, src/vanilla.jl:203-207), so ranks take contiguous blocks of roots. Philox streams are keyed
    by GLOBAL root index (`root_index_base`), hence the sharded result equals the single-GPU result bit-for-bit.
    Results are gathered with all_gather (4 + 8 B per root) only if the caller wants the full vectors.
  * Root-parallel search of one root (config 5): every rank grows `n_trees` independent trees of the same root
    with distinct Philox tree indices; per-rank partial sums [sum N_a | sum N_a*Q_a] (2*|A| doubles = 64 B for
    GridWorld) are merged with ONE allreduce over NVLink (NCCL) -- or gloo in the CPU tests -- and the decision
    is argmax_a sum(NQ)/sum(N), first max, as best_sanode_Q (src/vanilla.jl:339-349).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _abi as abi
from ._lib import Context
from .solvers import build_cfg


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of rank `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist

    return dist


def plan_sharded(ctx: Context, solver, mdp, states: Sequence, rank: Optional[int] = None, world: Optional[int] = None,
                 gather: bool = True, device=None):
    """Plan for `states` with the roots split across the ranks of the default process group.

    Returns (action_idx, best_q, stats): the full vectors on every rank if gather=True, else this rank's slice."""
    import torch

    dist = _dist()
    initialized = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if (rank is None and initialized) else (rank or 0)
    world = dist.get_world_size() if (world is None and initialized) else (world or 1)
    cfg, keep = build_cfg(solver, mdp)
    ctx.configure(mdp.mdp_id, mdp.params())
    roots = mdp.encode_states(states)
    lo, hi = shard_bounds(len(roots), rank, world)
    if hi > lo:
        act, bq, _, stats = ctx.plan(cfg, mdp.mdp_id, roots[lo:hi], root_index_base=lo)
    else:
        act, bq, stats = np.empty(0, np.int32), np.empty(0, np.float64), abi.PlanStats()
    del keep
    if not gather or world == 1:
        return act, bq, stats
    # ragged all_gather via padding to the largest shard
    width = shard_bounds(len(roots), 0, world)[1]
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    pa = torch.zeros(width, dtype=torch.int32, device=dev)
    pq = torch.zeros(width, dtype=torch.float64, device=dev)
    pa[: hi - lo] = torch.from_numpy(act).to(dev)
    pq[: hi - lo] = torch.from_numpy(bq).to(dev)
    ga = [torch.empty_like(pa) for _ in range(world)]
    gq = [torch.empty_like(pq) for _ in range(world)]
    dist.all_gather(ga, pa)
    dist.all_gather(gq, pq)
    acts, qs = [], []
    for r in range(world):
        l, h = shard_bounds(len(roots), r, world)
        acts.append(ga[r][: h - l].cpu().numpy())
        qs.append(gq[r][: h - l].cpu().numpy())
    return np.concatenate(acts), np.concatenate(qs), stats


def plan_root_parallel(ctx: Context, solver, mdp, state, n_trees_total: int, rank: Optional[int] = None,
                       world: Optional[int] = None, device=None):
    """Root-parallel decision for ONE root: `n_trees_total` trees of `solver.n_iterations` iterations each, split
    across ranks; one allreduce(sum) of 2*|A| doubles. Returns (action_idx, best_q, q_per_action, sum_n, stats)."""
    import torch

    dist = _dist()
    initialized = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if (rank is None and initialized) else (rank or 0)
    world = dist.get_world_size() if (world is None and initialized) else (world or 1)
    cfg, keep = build_cfg(solver, mdp)
    ctx.configure(mdp.mdp_id, mdp.params())
    root = mdp.encode_states([state])
    lo, hi = shard_bounds(n_trees_total, rank, world)
    na = len(mdp.actions)
    if hi > lo:
        sum_n, sum_nq, stats = ctx.plan_root_parallel(cfg, mdp.mdp_id, root, hi - lo, tree_index_base=lo)
    else:
        sum_n, sum_nq, stats = np.zeros(na), np.zeros(na), abi.PlanStats()
    del keep
    if world > 1:
        dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
        buf = torch.from_numpy(np.concatenate([sum_n, sum_nq])).to(dev)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)  # the only inter-GPU traffic of a decision: 2*|A| doubles
        merged = buf.cpu().numpy()
        sum_n, sum_nq = merged[:na], merged[na:]
    init_q = float(cfg.init_Q)
    a, bq, q = ctx.root_parallel_decide(sum_n, sum_nq, init_q)
    return a, bq, q, sum_n, stats
