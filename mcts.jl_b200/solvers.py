"""Host mirror of MCTS.jl's public API for the hot path, on top of the C ABI.

This is what the Julia glue (mcts.jl_b200/julia/MCTSB200.jl) does in Julia, written in Python because
Julia is not installed in the build image: solver structs with the reference's field names and
defaults (src/vanilla.jl:28-62, src/dpw_types.jl:45-100), `solve` (src/vanilla.jl:152, src/dpw.jl:1),
`action` / `action_info` (src/vanilla.jl:158-164, src/dpw.jl:13-74), `value` (src/vanilla.jl:169-197),
`clear_tree!` (src/vanilla.jl:148, src/dpw.jl:6-8), `ranked_actions` (src/util.jl:8-24), the
`default_action` handling around the DPW search (src/dpw.jl:68-71, src/default_action.jl) -- plus the
batch extension `action_batch` / `action_info_batch` (one tree per root state, all on the device).

Everything that iterates (the bodies of build_tree and of the DPW search block) happens inside
`mctsb200_plan`; nothing here computes a search step on the host.
"""
from __future__ import annotations

import math
import time
import warnings
from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Sequence

import numpy as np

from . import _abi as abi
from ._lib import Context, MCTSB200Error

_INF = float("inf")


# ----------------------------------------------------------------------------------------------
# rollout policies / estimators (src/domain_knowledge.jl:27-57)
# ----------------------------------------------------------------------------------------------
class RandomSolver:
    """POMDPTools.RandomSolver: solve() gives a RandomPolicy."""


class RandomPolicy:
    """POMDPTools.RandomPolicy: rand(rng, actions(mdp, s))."""


@dataclass
class ConstantPolicy:
    """FunctionPolicy(s -> a) for a fixed action `a` (e.g. `RolloutEstimator(x->:up)`, test/options.jl:29)."""

    action: Any


@dataclass
class FunctionPolicy:
    """POMDPTools.FunctionPolicy(f): action(policy, s) = f(s). On the device it becomes a table over the dense state space."""

    f: Callable


@dataclass
class SeekTarget:
    """notebooks/Domain_Knowledge_Example.ipynb cell 12-13: move towards `target` (x first, then y)."""

    target: tuple


class EnergyPolicy:
    """MountainCar heuristic: accelerate in the direction of motion (v < 0 ? -1 : +1)."""


@dataclass
class PluginPolicy:
    """A rollout policy implemented by a run-time MDP plugin's `policy_action` (id >= abi.POLICY_PLUGIN, up to 4 int params)."""

    policy_id: int = 100
    params: tuple = ()


@dataclass
class RolloutEstimator:
    """src/domain_knowledge.jl:27-37."""

    solver: Any = field(default_factory=RandomSolver)
    max_depth: int = 50
    eps: float = 0.0


class ExceptionRethrow:
    """src/default_action.jl: rethrow the search exception (the DPWSolver default)."""


# ----------------------------------------------------------------------------------------------
# solver structs
# ----------------------------------------------------------------------------------------------
def _seed_of(rng) -> int:
    if rng is None:
        return 0
    if isinstance(rng, (int, np.integer)):
        return int(rng) & 0xFFFFFFFFFFFFFFFF
    if isinstance(rng, np.random.Generator):
        return int(rng.integers(0, 2**63 - 1))
    raise TypeError(f"rng must be an int seed or numpy Generator, got {type(rng).__name__}")


@dataclass
class MCTSSolver:
    """Mirror of MCTSSolver (src/vanilla.jl:28-62); same keyword names and defaults."""

    n_iterations: int = 100
    max_time: float = _INF
    depth: int = 10
    exploration_constant: float = 1.0
    rng: Any = None
    estimate_value: Any = field(default_factory=RolloutEstimator)
    init_Q: Any = 0.0
    init_N: Any = 0
    reuse_tree: bool = False
    enable_tree_vis: bool = False
    timer: Optional[Callable] = None
    callback: Optional[Callable] = None
    # device extensions (not in the reference)
    callback_every: int = 1  # iterations between two calls of `callback` (1 = the reference's per-iteration callback)
    rollouts_per_leaf: int = 1
    lanes_per_tree: int = 0
    keep_tree_calls: int = 0  # reuse_tree on a model without a bounded state table: calls a tree has room for (0 = 8)


@dataclass
class DPWSolver:
    """Mirror of DPWSolver (src/dpw_types.jl:45-100); same keyword names and defaults."""

    depth: int = 10
    exploration_constant: float = 1.0
    n_iterations: int = 100
    max_time: float = _INF
    k_action: float = 10.0
    alpha_action: float = 0.5
    k_state: float = 10.0
    alpha_state: float = 0.5
    keep_tree: bool = False
    enable_action_pw: bool = True
    enable_state_pw: bool = True
    check_repeat_state: bool = True
    check_repeat_action: bool = True
    tree_in_info: bool = False
    rng: Any = None
    estimate_value: Any = field(default_factory=RolloutEstimator)
    init_Q: Any = 0.0
    init_N: Any = 0
    next_action: Any = None  # None = RandomActionGenerator(rng); f(mdp, s, snode) looking at s only: tabulated (dense-state MDPs)
    default_action: Any = field(default_factory=ExceptionRethrow)
    reset_callback: Optional[Callable] = None
    show_progress: bool = False
    timer: Optional[Callable] = None
    rollouts_per_leaf: int = 1
    lanes_per_tree: int = 0
    keep_tree_calls: int = 0  # keep_tree on a model without a bounded state table: calls a tree has room for (0 = 8)


def _unsupported(what: str) -> MCTSB200Error:
    return MCTSB200Error(abi.ERR_UNSUPPORTED, f"{what} cannot run on the device and there is no CPU fallback")


def _tabulate(f: Callable, mdp, per_action: bool, dtype) -> np.ndarray:
    """Evaluate a pure host function over the dense state space once (device-expressible stand-in for the
    Function/object forms of init_Q / init_N / estimate_value, src/domain_knowledge.jl:9,81,90)."""
    n = mdp.n_states()
    if per_action:
        out = np.zeros((n, len(mdp.actions)), dtype=dtype)
        for i in range(n):
            s = mdp.state_from_index(i)
            for j, a in enumerate(mdp.actions):
                out[i, j] = f(mdp, s, a)
    else:
        out = np.zeros(n, dtype=dtype)
        for i in range(n):
            out[i] = f(mdp, mdp.state_from_index(i), 0)
    return np.ascontiguousarray(out)


def build_cfg(solver, mdp) -> tuple[abi.SolverCfg, list]:
    """Flatten a solver struct into mctsb200_solver_cfg; raises for options that need host closures.
    Returns (cfg, keepalive) where keepalive holds the numpy tables the cfg points into."""
    import ctypes as C

    cfg = abi.SolverCfg()
    keep: list = []
    cfg.struct_size = C.sizeof(abi.SolverCfg)
    is_dpw = isinstance(solver, DPWSolver)
    cfg.kind = abi.KIND_DPW if is_dpw else abi.KIND_UCT
    for name in ("n_iterations", "depth"):
        if not isinstance(getattr(solver, name), (int, np.integer)):
            raise TypeError(f"{name} must be an Int")
    cfg.n_iterations = int(solver.n_iterations)
    cfg.depth = int(solver.depth)
    cfg.exploration_constant = float(solver.exploration_constant)
    cfg.max_time = float(solver.max_time)
    cfg.rollouts_per_leaf = int(solver.rollouts_per_leaf)
    cfg.lanes_per_tree = int(solver.lanes_per_tree)
    cfg.keep_tree_calls = int(getattr(solver, "keep_tree_calls", 0))
    cfg.seed = _seed_of(solver.rng)

    # `timer` and vanilla's `callback` run on the host between launches (Context.set_timer / set_progress_callback, installed
    # by _Planner._plan); the other host callbacks have no device form
    if is_dpw:
        if solver.reset_callback is not None:
            raise _unsupported("`reset_callback`")
        if solver.next_action is not None:
            # next_action(f::Function, mdp, s, snode) = f(mdp, s, snode) (src/action_gen.jl:14; test/options.jl:50 `(mdp, s, snode)->:up`):
            # a host function cannot run inside the search, but one that looks at the state only is a table over a dense state space
            if not callable(solver.next_action):
                raise TypeError("next_action must be callable as f(mdp, s, snode)")  # the reference's MethodError for `next_action="bad"`
            if not getattr(mdp, "dense", False) or getattr(mdp, "action_width", 0):
                raise _unsupported("a custom `next_action` on an MDP without a dense state index (write it into the device plugin's next_action)")
            try:
                tab = np.ascontiguousarray([mdp.action_index(solver.next_action(mdp, mdp.state_from_index(i), None)) for i in range(mdp.n_states())],
                                           dtype=np.int32)
            except (AttributeError, TypeError) as e:
                raise _unsupported(f"a `next_action` that inspects the tree node (snode is not available on the host: {e})")
            keep.append(tab)
            cfg.next_action_table = tab.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.k_action, cfg.alpha_action = float(solver.k_action), float(solver.alpha_action)
        cfg.k_state, cfg.alpha_state = float(solver.k_state), float(solver.alpha_state)
        cfg.enable_action_pw = int(bool(solver.enable_action_pw))
        cfg.enable_state_pw = int(bool(solver.enable_state_pw))
        cfg.check_repeat_state = int(bool(solver.check_repeat_state))
        cfg.check_repeat_action = int(bool(solver.check_repeat_action))
        cfg.keep_tree = int(bool(solver.keep_tree))
    else:
        cfg.keep_tree = int(bool(solver.reuse_tree))
        cfg.enable_tree_vis = int(bool(solver.enable_tree_vis))  # record_visit! (src/vanilla.jl:268-270): tree._vis_stats

    # init_Q / init_N: Number, ndarray table, or (dense MDPs) a pure function tabulated once
    def table_or_const(val, name, dtype, ctype):
        if isinstance(val, bool) or isinstance(val, str):
            raise TypeError(f"no method matching {name}(::{type(val).__name__}, ...)")  # MethodError parity (test/options.jl:18)
        if isinstance(val, (int, float, np.integer, np.floating)):
            return val, None
        if isinstance(val, np.ndarray) or callable(val):
            if not mdp.dense:
                raise _unsupported(f"a table/function-valued `{name}` on a non-dense MDP")
            tab = val if isinstance(val, np.ndarray) else _tabulate(val, mdp, True, dtype)
            tab = np.ascontiguousarray(tab, dtype=dtype).reshape(mdp.n_states(), len(mdp.actions))
            keep.append(tab)
            return None, tab.ctypes.data_as(C.POINTER(ctype))
        raise TypeError(f"no method matching {name}(::{type(val).__name__}, ...)")

    qc, qt = table_or_const(solver.init_Q, "init_Q", np.float64, C.c_double)
    nc, nt = table_or_const(solver.init_N, "init_N", np.int64, C.c_int64)
    cfg.init_Q = float(qc) if qc is not None else 0.0
    if nc is not None and int(nc) != nc:
        raise TypeError("init_N must be an Int (InexactError)")
    cfg.init_N = int(nc) if nc is not None else 0
    if qt is not None:
        cfg.init_q_table = qt
    if nt is not None:
        cfg.init_n_table = nt

    ev = solver.estimate_value
    cfg.rollout_max_depth = 50
    if isinstance(ev, bool) or isinstance(ev, str):
        raise TypeError(f"no method matching estimate_value(::{type(ev).__name__}, ...)")
    if isinstance(ev, (int, float, np.integer, np.floating)):
        cfg.estimate_kind = abi.EST_CONSTANT
        cfg.estimate_const = float(ev)
    elif isinstance(ev, RolloutEstimator):
        cfg.estimate_kind = abi.EST_ROLLOUT
        cfg.rollout_max_depth = int(ev.max_depth)
        cfg.rollout_eps = float(ev.eps)
        pol = ev.solver
        if isinstance(pol, type):
            pol = pol()
        if isinstance(pol, (RandomSolver, RandomPolicy)):
            cfg.rollout_policy = abi.POLICY_RANDOM
        elif isinstance(pol, ConstantPolicy):
            cfg.rollout_policy = abi.POLICY_CONSTANT
            cfg.rollout_policy_param[0] = mdp.action_index(pol.action)
        elif isinstance(pol, SeekTarget):
            cfg.rollout_policy = abi.POLICY_GW_SEEK
            cfg.rollout_policy_param[0], cfg.rollout_policy_param[1] = int(pol.target[0]), int(pol.target[1])
        elif isinstance(pol, EnergyPolicy):
            cfg.rollout_policy = abi.POLICY_MC_ENERGY
        elif isinstance(pol, PluginPolicy):
            if pol.policy_id < abi.POLICY_PLUGIN or len(pol.params) > 4:
                raise ValueError("PluginPolicy: policy_id >= 100 and at most 4 integer parameters")
            cfg.rollout_policy = int(pol.policy_id)
            for i, v in enumerate(pol.params):
                cfg.rollout_policy_param[i] = int(v)
        elif isinstance(pol, FunctionPolicy) or callable(pol):
            # FunctionPolicy(f) (src/domain_knowledge.jl:57; `RolloutEstimator(x->:up)`, test/options.jl:29): a state-only host function
            # cannot run inside the search, but on a dense-state MDP it is a table over the states
            f = pol.f if isinstance(pol, FunctionPolicy) else pol
            if not mdp.dense:
                raise _unsupported("a Function-valued rollout policy on a non-dense MDP (register it as a device rollout-policy plugin)")
            tab = np.ascontiguousarray([mdp.action_index(f(mdp.state_from_index(i))) for i in range(mdp.n_states())], dtype=np.int32)
            keep.append(tab)
            cfg.rollout_policy = abi.POLICY_TABLE
            cfg.policy_table = tab.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            raise TypeError(f"RolloutEstimator: unsupported policy {type(pol).__name__}")
    elif isinstance(ev, np.ndarray) or callable(ev):
        if not mdp.dense:
            raise _unsupported("a table/function-valued `estimate_value` on a non-dense MDP")
        tab = ev if isinstance(ev, np.ndarray) else _tabulate(lambda m, s, _a: ev(m, s, 0), mdp, False, np.float64)
        tab = np.ascontiguousarray(tab, dtype=np.float64).reshape(mdp.n_states())
        keep.append(tab)
        cfg.estimate_kind = abi.EST_TABLE
        cfg.value_table = tab.ctypes.data_as(C.POINTER(C.c_double))
    else:
        raise TypeError(f"no method matching estimate_value(::{type(ev).__name__}, ...)")
    return cfg, keep


# ----------------------------------------------------------------------------------------------
# tree views
# ----------------------------------------------------------------------------------------------
class TreeView:
    """info[:tree]: the arrays of MCTSTree (src/vanilla.jl:64-94) / DPWTree (src/dpw_types.jl:123-159),
    exported from the device on first access. Node ids are 1-based like in Julia."""

    def __init__(self, planner, root_i: int):
        self._planner, self._root_i, self._data, self._vis = planner, root_i, None, None

    def _load(self) -> dict:
        if self._data is None:
            m = self._planner.mdp
            self._data = self._planner.ctx.tree_export(self._root_i, m.state_dtype, m.state_width)
        return self._data

    def __getattr__(self, name):
        d = self._load()
        if name in d:
            return d[name]
        raise AttributeError(name)

    @property
    def state_map(self) -> dict:
        """state => state-node id (first occurrence; DPW without check_repeat_state may repeat states)."""
        d = self._load()
        m = self._planner.mdp
        out = {}
        for i, row in enumerate(d["s_labels"]):
            out.setdefault(m.decode_state(row), i + 1)
        return out

    s_lookup = state_map

    @property
    def _vis_stats(self) -> dict:
        """(said, spid) => visits of a vanilla tree planned with enable_tree_vis (src/vanilla.jl:75, 328-334): what
        D3Tree(::MCTSTree) draws below an action node (src/visualization.jl:52-122)."""
        if self._vis is None:
            self._vis = self._planner.ctx.tree_vis_stats(self._root_i)
        return self._vis

    def children(self, snode: int) -> np.ndarray:
        d = self._load()
        return d["child_ids"][d["child_offsets"][snode - 1]:d["child_offsets"][snode]]


# ----------------------------------------------------------------------------------------------
# planners
# ----------------------------------------------------------------------------------------------
class _Planner:
    def __init__(self, solver, mdp, ctx: Optional[Context] = None, device: int = 0):
        self.solver, self.mdp = solver, mdp
        self.ctx = ctx if ctx is not None else Context(device)
        self.ctx.configure(mdp.mdp_id, mdp.params())
        self.tree = None
        self._roots = None
        self._action_vectors = None
        self._calls = 0
        self.last_stats = None

    # one plan call over a batch of root states
    def _plan(self, states: Sequence, root_index_base: int = 0, want_status: bool = False):
        cfg, keep = build_cfg(self.solver, self.mdp)
        cfg.seed = (cfg.seed + self._calls) & 0xFFFFFFFFFFFFFFFF  # successive calls see fresh randomness
        self.ctx.configure(self.mdp.mdp_id, self.mdp.params())
        roots = self.mdp.encode_states(states)
        self._roots = [self.mdp.decode_state(r) for r in roots]
        cb = getattr(self.solver, "callback", None)
        if cb is None and getattr(self.solver, "show_progress", False):
            # DPWSolver.show_progress (src/dpw.jl:45-53: a ProgressMeter advanced once per iteration): a host hook between launches,
            # about twenty updates per search
            import sys

            def cb(_planner, done, total=int(cfg.n_iterations)):
                sys.stderr.write(f"\rMCTS (DPW) {done}/{total} iterations")
                if done >= total:
                    sys.stderr.write("\n")
                sys.stderr.flush()
        if cb is not None:  # callback(planner, iteration) (src/vanilla.jl:231); planner.tree is live inside it
            def progress(done, _total):
                self.tree = TreeView(self, 0)
                cb(self, done)
                return False

            every = int(getattr(self.solver, "callback_every", 0)) or max(1, int(cfg.n_iterations) // 20)
            self.ctx.set_progress_callback(progress, max(1, every))
        if self.solver.timer is not None:
            self.ctx.set_timer(self.solver.timer)
        try:
            actions, best_q, status, stats = self.ctx.plan(cfg, self.mdp.mdp_id, roots, root_index_base, want_status)
        finally:
            if cb is not None:
                self.ctx.set_progress_callback(None)
            if self.solver.timer is not None:
                self.ctx.set_timer(None)
        del keep
        self._calls += 1
        self.last_stats = stats
        # a continuous action space: the decisions are vectors (a_labels[sanode] of src/dpw.jl:65), fetched once per call
        self._action_vectors = self.ctx.root_action_vectors(0, len(roots), self.mdp.action_width) if getattr(self.mdp, "action_width", 0) else None
        return actions, best_q, status, stats

    def _label(self, idx: int, root_i: int = 0):
        if self._action_vectors is not None:
            return tuple(float(x) for x in self._action_vectors[root_i])
        return self.mdp.actions[int(idx)]

    def run_episodes(self, states: Sequence, max_steps: int, env_seed: int = 0, episode_index_base: int = 0):
        """One closed-loop episode per start state, planning at every step (mctsb200_run_episodes).
        Returns (discounted returns float64[n], steps int32[n], final states, stats)."""
        cfg, keep = build_cfg(self.solver, self.mdp)
        cfg.seed = (cfg.seed + self._calls) & 0xFFFFFFFFFFFFFFFF
        self.ctx.configure(self.mdp.mdp_id, self.mdp.params())
        s0 = self.mdp.encode_states(states)
        returns, steps, final, stats = self.ctx.run_episodes(cfg, self.mdp.mdp_id, s0, max_steps, env_seed, episode_index_base)
        del keep
        self._calls += int(steps.max()) if len(steps) else 0
        self._roots = None
        self.tree = None
        self.last_stats = stats
        return returns, steps, [self.mdp.decode_state(r) for r in final], stats

    def clear_tree(self):
        """clear_tree!(planner)"""
        self.tree = None
        self._roots = None
        self.ctx.clear_trees()

    def ranked_actions(self, state):
        """src/util.jl:8-24: children of the root sorted by Q, best first -> [(action, Q)]."""
        i = self._root_index(state)
        rs = self.ctx.root_stats(i, len(self.mdp.actions))
        order = sorted(range(len(rs["q"])), key=lambda j: -rs["q"][j])
        return [(self._label(rs["action_idx"][j]), float(rs["q"][j])) for j in order]

    def _root_index(self, state) -> int:
        if self._roots is None:
            self.action(state)
        st = self.mdp.decode_state(self.mdp.encode_states([state])[0])
        try:
            return self._roots.index(st)
        except ValueError:
            raise KeyError(f"State {state} was not a root of the last plan call.")


class MCTSPlanner(_Planner):
    """solve(::MCTSSolver, mdp) (src/vanilla.jl:137-152)."""

    def action_info_batch(self, states: Sequence, root_index_base: int = 0):
        actions, best_q, _, stats = self._plan(states, root_index_base)
        infos = [{"tree": TreeView(self, i), "best_Q": float(best_q[i])} for i in range(len(actions))]
        self.tree = infos[0]["tree"] if infos else None
        return [self._label(a) for a in actions], infos

    def action_batch(self, states: Sequence):
        return self.action_info_batch(states)[0]

    def action_info(self, state):
        acts, infos = self.action_info_batch([state])
        return acts[0], infos[0]

    def action(self, state):
        return self.action_info(state)[0]

    def action_root_parallel(self, state, n_trees: int, tree_index_base: int = 0):
        """Root-parallel decision for ONE root (BASELINE config 5; MCTSB200.jl: action_root_parallel): `n_trees` independent trees of
        `solver.n_iterations` iterations each, their root statistics merged in integer arithmetic -- on a multi-device ctx
        (`solve(..., device=[0, 1, ...])`) split over the devices and merged by one ncclAllReduce inside the library, with the same
        result whatever the number of devices. Returns (action, {best_Q, q, sum_n})."""
        cfg, keep = build_cfg(self.solver, self.mdp)
        self.ctx.configure(self.mdp.mdp_id, self.mdp.params())
        sn, snq, stats = self.ctx.plan_root_parallel(cfg, self.mdp.mdp_id, self.mdp.encode_states([state]), int(n_trees), tree_index_base)
        a, bq, q = self.ctx.root_parallel_decide(sn, snq, float(cfg.init_Q))
        del keep
        self.last_stats = stats
        return self._label(a), {"best_Q": bq, "q": q, "sum_n": sn}

    def value(self, state, a=None):
        """value(planner, s[, a]) (src/vanilla.jl:169-197)."""
        if self.tree is None:
            self.action(state)
        tree = self.tree
        sid = tree.state_map.get(self.mdp.decode_state(self.mdp.encode_states([state])[0]), 0)
        if sid == 0:
            raise KeyError(f"State {state} not present in MCTS tree.")
        ch = tree.children(sid)
        if a is None:
            return float(max(tree.q[c - 1] for c in ch))
        ai = self.mdp.action_index(a)
        for c in ch:
            if tree.a_labels[c - 1] == ai:
                return float(tree.q[c - 1])
        return None


class DPWPlanner(_Planner):
    """solve(::DPWSolver, mdp) (src/dpw_types.jl:211-227, src/dpw.jl:1)."""

    def action_info_batch(self, states: Sequence, tree_in_info: bool = False, root_index_base: int = 0):
        t0 = time.perf_counter()
        infos = []
        try:
            for s in states:
                if self.mdp.isterminal(s):
                    raise MCTSB200Error(abi.ERR_TERMINAL_ROOT, f"MCTS cannot handle terminal states. action was called with\ns = {s}")
            actions, best_q, _, stats = self._plan(states, root_index_base)
            dt = time.perf_counter() - t0
            out = []
            for i, a in enumerate(actions):
                info = {"search_time": dt, "search_time_us": dt * 1e6, "tree_queries": int(stats.iterations_done), "best_Q": float(best_q[i])}
                if self.solver.tree_in_info or tree_in_info:
                    info["tree"] = TreeView(self, i)
                infos.append(info)
                out.append(self._label(a, i))
            self.tree = TreeView(self, 0)
            return out, infos
        except MCTSB200Error as ex:
            # src/dpw.jl:68-71 + src/default_action.jl
            da = self.solver.default_action
            if isinstance(da, ExceptionRethrow):
                raise
            out = []
            for s in states:
                a = da(self.mdp, s, ex) if callable(da) else da
                warnings.warn(f"Using default action {a} because of the following error:\n{ex}")
                out.append(a)
                infos.append({"exception": ex})
            return out, infos

    def action_batch(self, states: Sequence):
        return self.action_info_batch(states)[0]

    def action_info(self, state, tree_in_info: bool = False):
        acts, infos = self.action_info_batch([state], tree_in_info=tree_in_info)
        return acts[0], infos[0]

    def action(self, state):
        return self.action_info(state)[0]


def solve(solver, mdp, ctx: Optional[Context] = None, device: int = 0):
    """POMDPs.solve: no computation, just binds solver and model (src/vanilla.jl:151-152)."""
    if isinstance(solver, DPWSolver):
        build_cfg(solver, mdp)  # reject unsupported options up front
        return DPWPlanner(solver, mdp, ctx, device)
    if isinstance(solver, MCTSSolver):
        build_cfg(solver, mdp)
        return MCTSPlanner(solver, mdp, ctx, device)
    raise TypeError(f"solve: unknown solver type {type(solver).__name__}")


class RolloutSimulator:
    """POMDPTools.RolloutSimulator(rng, max_steps): the simulator bench/non-terminal_gw.jl:18 and test/performance.jl:33 drive
    the planners with. eps is not supported (the device loop stops on max_steps or a terminal state only)."""

    def __init__(self, rng=None, max_steps: Optional[int] = None, eps=None):
        if eps is not None:
            raise _unsupported("RolloutSimulator(eps=...)")
        self.rng, self.max_steps = rng, max_steps


def simulate(sim: RolloutSimulator, mdp, planner: _Planner, initialstate):
    """simulate(sim, mdp, planner, s0) -> discounted return of one episode."""
    if sim.max_steps is None:
        raise _unsupported("RolloutSimulator without max_steps")
    assert mdp is planner.mdp
    return float(planner.run_episodes([initialstate], sim.max_steps, _seed_of(sim.rng))[0][0])


def simulate_batch(sim: RolloutSimulator, mdp, planner: _Planner, initialstates: Sequence) -> np.ndarray:
    """Batch extension: one independent episode per start state, all on the device in lockstep."""
    if sim.max_steps is None:
        raise _unsupported("RolloutSimulator without max_steps")
    assert mdp is planner.mdp
    return planner.run_episodes(initialstates, sim.max_steps, _seed_of(sim.rng))[0]


def clear_tree(planner: _Planner) -> None:
    planner.clear_tree()


def ranked_actions(planner: _Planner, state):
    return planner.ranked_actions(state)
