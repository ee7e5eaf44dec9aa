"""Multi-GPU plumbing: one process per GPU (torch.distributed), no data-path collective for batched
planning, one allreduce per decision for root-parallel search.

  * Batched planning (BASELINE configs 2-4): every root state's tree is independent (the reference plans one
    state per call, src/vanilla.jl:203-207), so ranks take contiguous blocks of roots. Philox streams are keyed
    by GLOBAL root index (`root_index_base`), hence the sharded result equals the single-GPU result bit-for-bit.
    Results are gathered with all_gather (4 + 8 B per root) only if the caller wants the full vectors.
  * Root-parallel search of one root (config 5): every rank grows `n_trees` independent trees of the same root
    with distinct Philox tree indices; the per-rank partial sums are INTEGER words (include/mctsb200.h:
    mctsb200_plan_root_parallel), merged with ONE allreduce per decision, and the decision is
    argmax_a sum(NQ)/sum(N), first max, as best_sanode_Q (src/vanilla.jl:339-349). On GPUs the allreduce is the
    library's own ncclAllReduce over NVLink (`attach_comm` below joins the ranks' ctxs to one communicator; the
    data never leaves the device before it is merged). Without a communicator (gloo in the CPU tests, MPI hosts)
    this module adds the int64 words with the process group and converts them with mctsb200_root_sums_merge --
    integer addition, so either way the merged sums equal a single-device run with the same trees bit for bit.
  * One process driving several GPUs needs none of this: `Context([0, 1, ...])` (mctsb200_ctx_create_multi)
    shards inside the library.
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


def attach_comm(ctx: Context, rank: Optional[int] = None, world: Optional[int] = None) -> None:
    """Joins this rank's ctx to an NCCL communicator over all ranks of the default process group: rank 0 makes the id
    (mctsb200_comm_unique_id), torch.distributed only carries its 128 bytes."""
    import torch

    dist = _dist()
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if world == 1 or ctx.comm_ranks > 1:
        return
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(buf, src=0)
    ctx.comm_init(bytes(buf.cpu().numpy().tobytes()), world, rank)


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
    across ranks; one allreduce(sum) of 5*|A|+1 int64 words (exact integer partial sums: the result does not depend on the number of
    ranks) -- inside the library over its NCCL communicator after `attach_comm`, else over the process group. Returns
    (action_idx, best_q, q_per_action, sum_n, stats)."""
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
    in_comm = ctx.comm_ranks > 1  # the library allreduces itself (attach_comm): every rank calls, also with no tree
    if hi > lo or in_comm:
        sum_n, sum_nq, stats = ctx.plan_root_parallel(cfg, mdp.mdp_id, root, hi - lo, tree_index_base=lo)
    else:
        sum_n, sum_nq, stats = np.zeros(na), np.zeros(na), abi.PlanStats()
    del keep
    if world > 1 and not in_comm:
        # no communicator (gloo, CPU tests): the process group adds the integer words, the library converts them
        words = ctx.root_sums_words() if hi > lo else np.zeros(5 * na + 1, dtype=np.int64)
        dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
        buf = torch.from_numpy(words).to(dev)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)  # the only inter-rank traffic of a decision: 5*|A|+1 int64 words
        sum_n, sum_nq = ctx.root_sums_merge(buf.cpu().numpy(), na)
    init_q = float(cfg.init_Q)
    a, bq, q = ctx.root_parallel_decide(sum_n, sum_nq, init_q)
    return a, bq, q, sum_n, stats
