"""ctypes binding of the CPU oracle (oracle/mcts_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg --
never from the product package.  Reuses the POD struct definitions of mcts.jl_b200/_abi.py (the oracle
takes the same mctsb200_solver_cfg so both sides are driven by identical bytes).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import mcts_jl_b200 as m  # noqa: E402

abi = m.abi
from mcts_jl_b200._lib import export_into_numpy  # noqa: E402

LIB = os.path.join(_HERE, "liboracle_mcts.so")
LIB_LIBM = os.path.join(_HERE, "liboracle_mcts_libm.so")


def build(force: bool = False) -> None:
    if force or not (os.path.exists(LIB) and os.path.exists(LIB_LIBM)):
        res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)


_cache: dict = {}


def load(libm: bool = False) -> C.CDLL:
    path = LIB_LIBM if libm else LIB
    if path in _cache:
        return _cache[path]
    build()
    lib = C.CDLL(path)
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    P = C.POINTER
    lib.oracle_plan.argtypes = [P(abi.SolverCfg), i32, vp, vp, i64, i64, i32, i32, i32, i32, vp, vp, vp, P(abi.PlanStats)]
    lib.oracle_plan.restype = i32
    lib.oracle_root_stats.argtypes = [i64, P(i32), vp, vp, vp, P(i64)]
    lib.oracle_root_stats.restype = i32
    lib.oracle_tree_sizes_query.argtypes = [i64, P(abi.TreeSizes)]
    lib.oracle_tree_sizes_query.restype = i32
    lib.oracle_tree_export.argtypes = [i64, P(abi.TreeBuffers)]
    lib.oracle_tree_vis_stats.argtypes = [i64, vp, vp, vp, i64, P(i64)]
    lib.oracle_tree_vis_stats.restype = i32
    lib.oracle_root_action_vectors.argtypes = [i64, i64, vp]
    lib.oracle_root_action_vectors.restype = i32
    lib.oracle_tree_export.restype = i32
    lib.oracle_plan_root_parallel.argtypes = [P(abi.SolverCfg), i32, vp, vp, i64, i64, i32, vp, vp, P(abi.PlanStats)]
    lib.oracle_plan_root_parallel.restype = i32
    lib.oracle_run_episodes.argtypes = [P(abi.SolverCfg), i32, vp, vp, i64, i32, u64, i64, i32, vp, vp, P(abi.PlanStats)]
    lib.oracle_run_episodes.restype = i32
    lib.oracle_philox.argtypes = [u64, u64, u32, u32, i32, vp]
    lib.oracle_philox_raw.argtypes = [vp, vp, vp]
    for f in ("oracle_log", "oracle_cos", "oracle_libm_log", "oracle_libm_cos"):
        getattr(lib, f).argtypes = [dbl]
        getattr(lib, f).restype = dbl
    lib.oracle_gen.argtypes = [i32, vp, vp, i32, u64, u64, u32, u32, vp, P(dbl)]
    lib.oracle_gen.restype = i32
    lib.oracle_rollout.argtypes = [P(abi.SolverCfg), i32, vp, vp, u64, u32, i64, P(i64)]
    lib.oracle_rollout.restype = dbl
    lib.oracle_hardware_threads.restype = i32
    _cache[path] = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Same call shapes as mcts_jl_b200.Context so tests can drive both with the same arguments."""

    def __init__(self, libm: bool = False):
        self.lib = load(libm)
        self._params = {}

    def configure(self, mdp_id: int, params) -> None:
        if mdp_id in self._params and bytes(self._params[mdp_id]) != bytes(params):
            self.clear_trees()  # like mctsb200_mdp_configure: a new model invalidates the trees kept for reuse
        self._params[mdp_id] = params

    def clear_trees(self) -> None:
        self.lib.oracle_clear_trees()

    def plan(self, cfg, mdp_id: int, roots: np.ndarray, root_index_base: int = 0, n_threads: int = 1, keep_trees: bool = True,
             literal_transitions: bool = False, same_root: bool = False, n_trees: int | None = None):
        roots = np.ascontiguousarray(roots)
        n = roots.shape[0] if n_trees is None else n_trees
        actions = np.empty(n, dtype=np.int32)
        best_q = np.empty(n, dtype=np.float64)
        status = np.zeros(n, dtype=np.int32)
        stats = abi.PlanStats()
        p = self._params[mdp_id]
        rc = self.lib.oracle_plan(C.byref(cfg), mdp_id, C.byref(p), _ptr(roots), n, int(root_index_base), int(n_threads), int(keep_trees),
                                  int(literal_transitions), int(same_root), _ptr(actions), _ptr(best_q), _ptr(status), C.byref(stats))
        if rc != 0:
            raise RuntimeError(f"oracle_plan failed with {rc}")
        return actions, best_q, status, stats

    def run_episodes(self, cfg, mdp_id: int, states: np.ndarray, max_steps: int, env_seed: int = 0, episode_index_base: int = 0,
                     n_threads: int = 1):
        """simulate(RolloutSimulator(max_steps), mdp, planner, s0) for every row of states; same shapes as Context.run_episodes."""
        final = np.array(states, copy=True, order="C")
        n = final.shape[0]
        returns = np.empty(n, dtype=np.float64)
        steps = np.empty(n, dtype=np.int32)
        stats = abi.PlanStats()
        rc = self.lib.oracle_run_episodes(C.byref(cfg), mdp_id, C.byref(self._params[mdp_id]), _ptr(final), n, int(max_steps), int(env_seed),
                                          int(episode_index_base), int(n_threads), _ptr(returns), _ptr(steps), C.byref(stats))
        if rc != 0:
            raise RuntimeError(f"oracle_run_episodes failed with {rc}")
        return returns, steps, final, stats

    def plan_root_parallel(self, cfg, mdp_id: int, root: np.ndarray, n_trees: int, tree_index_base: int = 0, n_threads: int = 1):
        root = np.ascontiguousarray(root)
        na = 4 if mdp_id == abi.MDP_GRIDWORLD else 3
        sum_n = np.zeros(na)
        sum_nq = np.zeros(na)
        stats = abi.PlanStats()
        rc = self.lib.oracle_plan_root_parallel(C.byref(cfg), mdp_id, C.byref(self._params[mdp_id]), _ptr(root), int(n_trees),
                                                int(tree_index_base), int(n_threads), _ptr(sum_n), _ptr(sum_nq), C.byref(stats))
        if rc != 0:
            raise RuntimeError(f"oracle_plan_root_parallel failed with {rc}")
        return sum_n, sum_nq, stats

    def tree_vis_stats(self, root_i: int) -> dict:
        n = C.c_int64()
        assert self.lib.oracle_tree_vis_stats(int(root_i), None, None, None, 0, C.byref(n)) == 0
        said, spid, cnt = (np.zeros(n.value, dtype=np.int64) for _ in range(3))
        assert self.lib.oracle_tree_vis_stats(int(root_i), _ptr(said), _ptr(spid), _ptr(cnt), n.value, C.byref(n)) == 0
        return {(int(a), int(b)): int(c) for a, b, c in zip(said, spid, cnt)}

    def root_action_vectors(self, first_root: int, n: int, action_doubles: int = abi.BOX_ACTION_DOUBLES) -> np.ndarray:
        out = np.zeros((int(n), int(action_doubles)), dtype=np.float64)
        rc = self.lib.oracle_root_action_vectors(int(first_root), int(n), _ptr(out))
        assert rc == 0, rc
        return out

    def root_stats(self, root_i: int, n_actions: int = 16) -> dict:
        nch, tot = C.c_int32(), C.c_int64()
        act = np.zeros(n_actions, dtype=np.int32)
        n = np.zeros(n_actions, dtype=np.int64)
        q = np.zeros(n_actions, dtype=np.float64)
        rc = self.lib.oracle_root_stats(int(root_i), C.byref(nch), _ptr(act), _ptr(n), _ptr(q), C.byref(tot))
        assert rc == 0
        k = nch.value
        return {"action_idx": act[:k].copy(), "n": n[:k].copy(), "q": q[:k].copy(), "total_n": tot.value}

    def tree_export(self, root_i: int, state_dtype=np.int64, state_width: int = 2) -> dict:
        sz = abi.TreeSizes()
        rc = self.lib.oracle_tree_sizes_query(int(root_i), C.byref(sz))
        assert rc == 0, rc

        def call(b):
            assert self.lib.oracle_tree_export(int(root_i), C.byref(b)) == 0

        return export_into_numpy(call, sz, state_dtype, state_width)

    def philox(self, seed: int, tree: int, iteration: int, stream_id: int, n_blocks: int) -> np.ndarray:
        out = np.empty(4 * n_blocks, dtype=np.uint32)
        self.lib.oracle_philox(seed, tree, iteration, stream_id, n_blocks, _ptr(out))
        return out

    def philox_raw(self, key, ctr) -> np.ndarray:
        key = np.asarray(key, dtype=np.uint32)
        ctr = np.asarray(ctr, dtype=np.uint32)
        out = np.empty(4, dtype=np.uint32)
        self.lib.oracle_philox_raw(_ptr(key), _ptr(ctr), _ptr(out))
        return out

    def gen(self, mdp_id: int, s, a: int, seed=0, tree=0, iteration=0, stream_id=0):
        dt = np.int64 if mdp_id == abi.MDP_GRIDWORLD else np.float64
        s = np.ascontiguousarray(s, dtype=dt)
        sp = np.empty(2, dtype=dt)
        r = C.c_double()
        used = self.lib.oracle_gen(mdp_id, C.byref(self._params[mdp_id]), _ptr(s), int(a), seed, tree, iteration, stream_id, _ptr(sp), C.byref(r))
        return sp, r.value, used

    def rollout(self, cfg, mdp_id: int, s, tree=0, iteration=0, remaining_depth=0):
        dt = np.int64 if mdp_id == abi.MDP_GRIDWORLD else np.float64
        s = np.ascontiguousarray(s, dtype=dt)
        steps = C.c_int64()
        v = self.lib.oracle_rollout(C.byref(cfg), mdp_id, C.byref(self._params[mdp_id]), _ptr(s), tree, iteration, remaining_depth, C.byref(steps))
        return v, steps.value

    def hardware_threads(self) -> int:
        return int(self.lib.oracle_hardware_threads())
