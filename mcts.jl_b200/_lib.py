"""Loader and thin object wrapper for libmctsb200.so (the C ABI of include/mctsb200.h).

There is no CPU fallback: if the CUDA library has not been built, or no B200 is visible, every
compute entry point raises.  (The CPU-only test-suite links a host emulation of the same sources,
tests/emu, and hands its path to `load_library` explicitly; nothing in this package looks for it.)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
DEFAULT_LIB = os.path.join(CSRC_DIR, "libmctsb200.so")

_lib_cache: dict[str, C.CDLL] = {}


class MCTSB200Error(RuntimeError):
    """Raised for every non-zero status of the C ABI (the Julia glue calls error() the same way)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libmctsb200 [{abi.STATUS_NAMES.get(status, status)}]: {message}")
        self.status = status
        self.message = message


def build_library(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with nvcc (in-tree, so the .so travels with the repo snapshot)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libmctsb200.so failed")
    return DEFAULT_LIB


def load_library(path: Optional[str] = None) -> C.CDLL:
    # MCTSB200_LIB: tuning experiments only (an alternative build of the same CUDA library, e.g. other -D switches)
    explicit = path is not None
    path = os.path.abspath(path or os.environ.get("MCTSB200_LIB") or DEFAULT_LIB)
    if path in _lib_cache and (explicit or not hasattr(_lib_cache[path], "mctsb200_host_emulation")):
        return _lib_cache[path]
    if not os.path.exists(path):
        raise MCTSB200Error(
            abi.ERR_CUDA,
            f"{path} not found: the CUDA extension has not been built (run `python __graft_entry__.py` or "
            f"`make -C {CSRC_DIR}`). There is no CPU fallback.",
        )
    lib = abi.declare(C.CDLL(path))
    if not explicit and hasattr(lib, "mctsb200_host_emulation"):
        # the tests' host-emulation build is a CPU program: only a test that passes its path explicitly may load it
        raise MCTSB200Error(abi.ERR_CUDA, f"{path} is the tests' host-emulation build, not the CUDA library. There is no CPU fallback.")
    got = lib.mctsb200_abi_version()
    if got != abi.ABI_VERSION:
        raise MCTSB200Error(abi.ERR_INVALID_ARG, f"ABI version {got} != {abi.ABI_VERSION}")
    _lib_cache[path] = lib
    return lib


def register_mdp(name: str, cuda_source: str, plugin_struct: str, state_doubles: int, n_actions: int, max_branch: int = 0,
                 lib_path: Optional[str] = None) -> int:
    """mctsb200_mdp_register: compile the tile kernel around a device-side MDP plugin (NVRTC, sm_100a) and return its
    process-wide mdp_id. Compiler errors raise MCTSB200Error with the NVRTC log. Needs no GPU."""
    lib = load_library(lib_path)
    out = C.c_int32(-1)
    rc = lib.mctsb200_mdp_register(name.encode(), cuda_source.encode(), plugin_struct.encode(), int(state_doubles), int(n_actions),
                                   int(max_branch), C.byref(out))
    if rc != abi.OK:
        raise MCTSB200Error(rc, lib.mctsb200_last_error().decode("utf-8", "replace"))
    return out.value


def register_mdp_continuous(name: str, cuda_source: str, plugin_struct: str, state_doubles: int, action_doubles: int,
                            lib_path: Optional[str] = None) -> int:
    """mctsb200_mdp_register_continuous: the same for a plugin whose actions are vectors of `action_doubles` float64 values
    (actions(mdp) = a Box); its DPW program is compiled here."""
    lib = load_library(lib_path)
    out = C.c_int32(-1)
    rc = lib.mctsb200_mdp_register_continuous(name.encode(), cuda_source.encode(), plugin_struct.encode(), int(state_doubles), int(action_doubles),
                                              C.byref(out))
    if rc != abi.OK:
        raise MCTSB200Error(rc, lib.mctsb200_last_error().decode("utf-8", "replace"))
    return out.value


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One mctsb200_ctx, not re-entrant (like a reference planner). `device` is a CUDA ordinal, or a sequence of ordinals for a
    multi-device ctx (mctsb200_ctx_create_multi): the same calls then shard roots / episodes / ensemble trees over the devices
    inside the library and return what one device would, bit for bit."""

    def __init__(self, device=0, stream: int = 0, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        self._h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int32 * len(device))(*[int(d) for d in device])
            self._check(self.lib.mctsb200_ctx_create_multi(ids, len(device), C.byref(self._h)))
            self.devices = [int(d) for d in device]
            self.device = self.devices[0]
        else:
            self._check(self.lib.mctsb200_ctx_create(int(device), C.c_void_p(stream or None), C.byref(self._h)))
            self.device = int(device)
            self.devices = [self.device]
        self._configured: dict[int, bytes] = {}
        self.comm_ranks = 1

    # -- one process per GPU: join a communicator (the id travels by the host's own means) --
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.lib.mctsb200_comm_unique_id(buf, 128))
        return buf.raw

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int) -> None:
        """After this, plan_root_parallel on this ctx is a collective call (one ncclAllReduce inside the library)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.mctsb200_comm_init(self._h, buf, 128, int(n_ranks), int(rank)))
        self.comm_ranks = int(n_ranks)

    def root_sums_words(self) -> np.ndarray:
        """Integer words behind the sums of the last root-parallel call (for hosts that merge with their own transport)."""
        n = C.c_int32()
        self._check(self.lib.mctsb200_root_sums_words(self._h, None, 0, C.byref(n)))
        w = np.zeros(n.value, dtype=np.int64)
        self._check(self.lib.mctsb200_root_sums_words(self._h, _ptr(w), n.value, C.byref(n)))
        return w

    def root_sums_merge(self, words: np.ndarray, n_actions: int):
        words = np.ascontiguousarray(words, dtype=np.int64)
        assert len(words) == 5 * n_actions + 1
        sn, snq = np.zeros(n_actions), np.zeros(n_actions)
        self._check(self.lib.mctsb200_root_sums_merge(_ptr(words), int(n_actions), _ptr(sn), _ptr(snq)))
        return sn, snq

    # -- plumbing --
    def _check(self, status: int) -> None:
        if status != abi.OK:
            raise MCTSB200Error(status, self.lib.mctsb200_last_error().decode("utf-8", "replace"))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.mctsb200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- MDP plugins --
    def mdp_lookup(self, name: str) -> int:
        out = C.c_int32()
        self._check(self.lib.mctsb200_mdp_lookup(name.encode(), C.byref(out)))
        return out.value

    def mdp_info(self, mdp_id: int) -> tuple[int, int]:
        na, sb = C.c_int32(), C.c_int64()
        self._check(self.lib.mctsb200_mdp_info(mdp_id, C.byref(na), C.byref(sb)))
        return na.value, sb.value

    def configure(self, mdp_id: int, params: C.Structure) -> None:
        raw = bytes(params)
        if self._configured.get(mdp_id) == raw:
            return
        self._check(self.lib.mctsb200_mdp_configure(self._h, mdp_id, C.byref(params), C.sizeof(params)))
        self._configured[mdp_id] = raw

    # -- planning --
    def plan(self, cfg: abi.SolverCfg, mdp_id: int, roots: np.ndarray, root_index_base: int = 0,
             want_status: bool = False, out=None):
        """roots: C-contiguous array whose rows are statetype(mdp) values. Returns (actions, best_q, status, stats).
        With want_status=False a terminal DPW root raises (like src/dpw.jl:22-27); with True it is reported per root.
        out: optional preallocated (actions int32[n], best_q float64[n], status int32[n]) host arrays (e.g. pinned)."""
        roots = np.ascontiguousarray(roots)
        n = roots.shape[0]
        state_bytes = roots.dtype.itemsize * int(np.prod(roots.shape[1:], dtype=np.int64)) if n else 0
        if out is not None:
            actions, best_q, status = out
            assert actions.dtype == np.int32 and best_q.dtype == np.float64 and len(actions) == n and len(best_q) == n
        else:
            actions = np.empty(n, dtype=np.int32)
            best_q = np.empty(n, dtype=np.float64)
            status = np.zeros(n, dtype=np.int32) if want_status else None
        stats = abi.PlanStats()
        self._check(self.lib.mctsb200_plan(self._h, C.byref(cfg), mdp_id, _ptr(roots), n, state_bytes, int(root_index_base),
                                           _ptr(actions), _ptr(best_q), _ptr(status), C.byref(stats)))
        return actions, best_q, status, stats

    def plan_device(self, cfg: abi.SolverCfg, mdp_id: int, d_roots: int, n_roots: int, state_bytes: int,
                    d_actions: int, d_best_q: int, d_status: int = 0, root_index_base: int = 0) -> abi.PlanStats:
        """Device-pointer variant (e.g. torch tensors' data_ptr())."""
        stats = abi.PlanStats()
        self._check(self.lib.mctsb200_plan_device(self._h, C.byref(cfg), mdp_id, C.c_void_p(d_roots), int(n_roots), int(state_bytes),
                                                  int(root_index_base), C.c_void_p(d_actions or None), C.c_void_p(d_best_q or None),
                                                  C.c_void_p(d_status or None), C.byref(stats)))
        return stats

    def run_episodes(self, cfg: abi.SolverCfg, mdp_id: int, states: np.ndarray, max_steps: int, env_seed: int = 0,
                     episode_index_base: int = 0):
        """simulate(RolloutSimulator(max_steps=max_steps), mdp, planner, s0) for every row of states, on the device.
        Returns (discounted returns, steps taken, final states, stats summed over the plan calls)."""
        states = np.ascontiguousarray(states)
        n = states.shape[0]
        state_bytes = states.dtype.itemsize * int(np.prod(states.shape[1:], dtype=np.int64)) if n else 0
        returns = np.empty(n, dtype=np.float64)
        steps = np.empty(n, dtype=np.int32)
        final = np.empty_like(states)
        stats = abi.PlanStats()
        self._check(self.lib.mctsb200_run_episodes(self._h, C.byref(cfg), mdp_id, _ptr(states), n, state_bytes, int(max_steps), int(env_seed),
                                                   int(episode_index_base), _ptr(returns), _ptr(steps), _ptr(final), C.byref(stats)))
        return returns, steps, final, stats

    # -- host hooks between launches --
    def set_progress_callback(self, fn=None, every: int = 1) -> None:
        """fn(iterations_done, n_iterations) -> truthy to stop; called after every `every` iterations of each plan call. None clears."""
        if fn is None:
            self._progress = abi.PROGRESS_FN()  # NULL
            self._check(self.lib.mctsb200_set_progress_callback(self._h, self._progress, None, 0))
            return

        def thunk(_user, done, total):
            return 1 if fn(int(done), int(total)) else 0

        self._progress = abi.PROGRESS_FN(thunk)  # (kept alive as long as it is installed)
        self._check(self.lib.mctsb200_set_progress_callback(self._h, self._progress, None, int(every)))

    def set_timer(self, fn=None) -> None:
        """fn() -> seconds: the clock max_time is measured with (solver.timer). None restores the monotonic clock."""
        self._timer = abi.TIMER_FN((lambda _user: float(fn())) if fn is not None else 0)
        self._check(self.lib.mctsb200_set_timer(self._h, self._timer, None))

    def plan_root_parallel(self, cfg: abi.SolverCfg, mdp_id: int, root: np.ndarray, n_trees: int, tree_index_base: int = 0,
                           d_sums: int = 0):
        root = np.ascontiguousarray(root)
        na, _ = self.mdp_info(mdp_id)
        sum_n = np.zeros(na, dtype=np.float64)
        sum_nq = np.zeros(na, dtype=np.float64)
        stats = abi.PlanStats()
        self._check(self.lib.mctsb200_plan_root_parallel(self._h, C.byref(cfg), mdp_id, _ptr(root), root.nbytes, int(n_trees),
                                                         int(tree_index_base), _ptr(sum_n), _ptr(sum_nq), C.c_void_p(d_sums or None),
                                                         C.byref(stats)))
        return sum_n, sum_nq, stats

    def root_parallel_decide(self, sum_n: np.ndarray, sum_nq: np.ndarray, init_q: float = 0.0):
        sum_n = np.ascontiguousarray(sum_n, dtype=np.float64)
        sum_nq = np.ascontiguousarray(sum_nq, dtype=np.float64)
        q = np.zeros_like(sum_n)
        a, bq = C.c_int32(), C.c_double()
        self._check(self.lib.mctsb200_root_parallel_decide(_ptr(sum_n), _ptr(sum_nq), len(sum_n), float(init_q), C.byref(a), C.byref(bq), _ptr(q)))
        return a.value, bq.value, q

    # -- tree queries --
    def root_stats(self, root_i: int, n_actions: Optional[int] = None) -> dict:
        """Children of the root of tree root_i (mctsb200_root_stats_n, two calls: the count, then arrays of exactly that size -- a
        continuous action space puts no bound on it). `n_actions` caps what is returned."""
        nch, tot = C.c_int32(), C.c_int64()
        self._check(self.lib.mctsb200_root_stats_n(self._h, int(root_i), 0, C.byref(nch), None, None, None, C.byref(tot)))
        cap = nch.value if n_actions is None else min(int(n_actions), nch.value)
        act = np.zeros(cap, dtype=np.int32)
        n = np.zeros(cap, dtype=np.int64)
        q = np.zeros(cap, dtype=np.float64)
        if cap:
            self._check(self.lib.mctsb200_root_stats_n(self._h, int(root_i), cap, C.byref(nch), _ptr(act), _ptr(n), _ptr(q), C.byref(tot)))
        return {"action_idx": act, "n": n, "q": q, "total_n": tot.value}

    def tree_export(self, root_i: int, state_dtype=np.int64, state_width: int = 2) -> dict:
        """Two-call export into numpy arrays named after MCTSTree / DPWTree fields (ids are 1-based)."""
        sz = abi.TreeSizes()
        self._check(self.lib.mctsb200_tree_sizes_query(self._h, int(root_i), C.byref(sz)))
        return export_into_numpy(lambda b: self._check(self.lib.mctsb200_tree_export(self._h, int(root_i), C.byref(b))), sz,
                                 state_dtype, state_width)

    def clear_trees(self) -> None:
        self._check(self.lib.mctsb200_clear_trees(self._h))

    def tree_vis_stats(self, root_i: int) -> dict:
        """tree._vis_stats of a vanilla tree planned with enable_tree_vis: {(said, spid): visits}, ids 1-based (src/vanilla.jl:328-334)."""
        n = C.c_int64()
        self._check(self.lib.mctsb200_tree_vis_stats(self._h, int(root_i), None, None, None, 0, C.byref(n)))
        said, spid, cnt = (np.zeros(n.value, dtype=np.int64) for _ in range(3))
        self._check(self.lib.mctsb200_tree_vis_stats(self._h, int(root_i), _ptr(said), _ptr(spid), _ptr(cnt), n.value, C.byref(n)))
        return {(int(a), int(b)): int(c) for a, b, c in zip(said, spid, cnt)}

    def root_action_vectors(self, first_root: int, n: int, action_doubles: int = abi.BOX_ACTION_DOUBLES) -> np.ndarray:
        """Continuous action spaces: the action vectors best_sanode chose for roots [first_root, first_root + n) of the last plan call."""
        out = np.zeros((int(n), int(action_doubles)), dtype=np.float64)
        self._check(self.lib.mctsb200_root_action_vectors(self._h, int(first_root), int(n), _ptr(out)))
        return out

    # -- numerical-contract probes --
    def device_log(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        self._check(self.lib.mctsb200_device_log(self._h, _ptr(x), x.size, _ptr(out)))
        return out

    def device_cos(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        self._check(self.lib.mctsb200_device_cos(self._h, _ptr(x), x.size, _ptr(out)))
        return out

    def device_philox(self, seed: int, tree: int, iteration: int, stream_id: int, n_blocks: int) -> np.ndarray:
        out = np.empty(4 * n_blocks, dtype=np.uint32)
        self._check(self.lib.mctsb200_device_philox(self._h, seed, tree, iteration, stream_id, n_blocks, _ptr(out)))
        return out


def export_into_numpy(call, sz: abi.TreeSizes, state_dtype, state_width: int) -> dict:
    ns, na, nt = sz.n_state_nodes, sz.n_action_nodes, sz.n_transitions
    t = {
        "total_n": np.zeros(ns, dtype=np.int64),
        "s_labels": np.zeros((ns, state_width), dtype=state_dtype),
        "child_offsets": np.zeros(ns + 1, dtype=np.int64),
        "child_ids": np.zeros(na, dtype=np.int64),
        "n": np.zeros(na, dtype=np.int64),
        "q": np.zeros(na, dtype=np.float64),
        "a_labels": np.zeros(na, dtype=np.int32),
        "n_a_children": np.zeros(na, dtype=np.int64),
        "trans_offsets": np.zeros(na + 1, dtype=np.int64),
        "trans_spnode": np.zeros(nt, dtype=np.int64),
        "trans_r": np.zeros(nt, dtype=np.float64),
        "trans_count": np.zeros(nt, dtype=np.int64),
    }
    if sz.action_doubles > 0:  # a continuous action space: the labels are vectors (a_labels holds -1)
        t["a_label_vectors"] = np.zeros((na, int(sz.action_doubles)), dtype=np.float64)
    b = abi.TreeBuffers(**{k: v.ctypes.data for k, v in t.items()})
    call(b)
    return t
