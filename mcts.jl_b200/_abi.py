"""ctypes mirror of include/mctsb200.h (struct layouts, status codes, enum values).

Kept separate from the loader so that test infrastructure (oracle/oracle_binding.py) can reuse the POD
definitions without touching the CUDA library.
"""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 2

OK = 0
ERR_INVALID_ARG = -1
ERR_UNSUPPORTED = -2
ERR_TERMINAL_ROOT = -3
ERR_CAPACITY = -4
ERR_CUDA = -5
ERR_STATE = -6

STATUS_NAMES = {
    OK: "OK",
    ERR_INVALID_ARG: "INVALID_ARG",
    ERR_UNSUPPORTED: "UNSUPPORTED",
    ERR_TERMINAL_ROOT: "TERMINAL_ROOT",
    ERR_CAPACITY: "CAPACITY",
    ERR_CUDA: "CUDA",
    ERR_STATE: "STATE",
}

MDP_GRIDWORLD = 0
MDP_MOUNTAINCAR = 1
MDP_GAUSSIAN_BOX = 2
BOX_STATE_DOUBLES = 4
BOX_ACTION_DOUBLES = 2
MDP_PLUGIN_BASE = 100      # ids of MDP plugins registered at run time (mctsb200_mdp_register)
PLUGIN_PARAM_BYTES = 256

KIND_UCT = 0
KIND_DPW = 1

EST_ROLLOUT = 0
EST_CONSTANT = 1
EST_TABLE = 2

POLICY_RANDOM = 0
POLICY_CONSTANT = 1
POLICY_GW_SEEK = 2
POLICY_MC_ENERGY = 3
POLICY_TABLE = 4
POLICY_PLUGIN = 100        # ids >= this go to the plugin's policy_action

GW_MAX_SPECIAL = 32


class GridWorldParams(C.Structure):
    _fields_ = [
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("tprob", C.c_double),
        ("discount", C.c_double),
        ("n_rewards", C.c_int32),
        ("n_terminate", C.c_int32),
        ("reward_x", C.c_int32 * GW_MAX_SPECIAL),
        ("reward_y", C.c_int32 * GW_MAX_SPECIAL),
        ("reward_v", C.c_double * GW_MAX_SPECIAL),
        ("term_x", C.c_int32 * GW_MAX_SPECIAL),
        ("term_y", C.c_int32 * GW_MAX_SPECIAL),
    ]


class MountainCarParams(C.Structure):
    _fields_ = [("discount", C.c_double), ("cost", C.c_double), ("jackpot", C.c_double)]


class GaussianBoxParams(C.Structure):
    _fields_ = [("discount", C.c_double), ("a_lo", C.c_double * BOX_ACTION_DOUBLES), ("a_hi", C.c_double * BOX_ACTION_DOUBLES),
                ("var", C.c_double * 3), ("drift", C.c_double), ("cost", C.c_double)]


class SolverCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("kind", C.c_int32),
        ("n_iterations", C.c_int64),
        ("depth", C.c_int32),
        ("rollouts_per_leaf", C.c_int32),
        ("exploration_constant", C.c_double),
        ("max_time", C.c_double),
        ("k_action", C.c_double),
        ("alpha_action", C.c_double),
        ("k_state", C.c_double),
        ("alpha_state", C.c_double),
        ("enable_action_pw", C.c_int32),
        ("enable_state_pw", C.c_int32),
        ("check_repeat_state", C.c_int32),
        ("check_repeat_action", C.c_int32),
        ("keep_tree", C.c_int32),
        ("lanes_per_tree", C.c_int32),
        ("init_Q", C.c_double),
        ("init_N", C.c_int64),
        ("estimate_kind", C.c_int32),
        ("rollout_policy", C.c_int32),
        ("estimate_const", C.c_double),
        ("rollout_policy_param", C.c_int32 * 4),
        ("rollout_max_depth", C.c_int32),
        ("keep_tree_calls", C.c_int32),
        ("rollout_eps", C.c_double),
        ("seed", C.c_uint64),
        ("init_q_table", C.POINTER(C.c_double)),
        ("init_n_table", C.POINTER(C.c_int64)),
        ("value_table", C.POINTER(C.c_double)),
        ("policy_table", C.POINTER(C.c_int32)),
        ("enable_tree_vis", C.c_int32),
        ("reserved1", C.c_int32),
        ("next_action_table", C.POINTER(C.c_int32)),
    ]


class PlanStats(C.Structure):
    _fields_ = [
        ("n_trees", C.c_int64),
        ("n_simulations", C.c_int64),
        ("n_tree_levels", C.c_int64),
        ("n_dpw_children_seen", C.c_int64),
        ("n_state_inserts", C.c_int64),
        ("n_action_inserts", C.c_int64),
        ("n_rollouts", C.c_int64),
        ("n_rollout_steps", C.c_int64),
        ("n_transition_samples", C.c_int64),
        ("n_select_exact", C.c_int64),
        ("algorithmic_bytes", C.c_int64),
        ("kernel_ms", C.c_double),
        ("total_ms", C.c_double),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("lanes_per_tree", C.c_int32),
        ("n_kernel_launches", C.c_int32),
        ("iterations_done", C.c_int32),
        ("kernel_variant", C.c_int32),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_ if not name.startswith("reserved")}


class TreeSizes(C.Structure):
    _fields_ = [
        ("n_state_nodes", C.c_int64),
        ("n_action_nodes", C.c_int64),
        ("n_transitions", C.c_int64),
        ("state_bytes", C.c_int64),
        ("action_doubles", C.c_int64),
    ]


class TreeBuffers(C.Structure):
    _fields_ = [
        ("total_n", C.c_void_p),
        ("s_labels", C.c_void_p),
        ("child_offsets", C.c_void_p),
        ("child_ids", C.c_void_p),
        ("n", C.c_void_p),
        ("q", C.c_void_p),
        ("a_labels", C.c_void_p),
        ("n_a_children", C.c_void_p),
        ("trans_offsets", C.c_void_p),
        ("trans_spnode", C.c_void_p),
        ("trans_r", C.c_void_p),
        ("trans_count", C.c_void_p),
        ("a_label_vectors", C.c_void_p),
    ]


# every symbol include/mctsb200.h declares (tests check the built library exports all of them)
EXPORTED_SYMBOLS = (
    "mctsb200_abi_version",
    "mctsb200_last_error",
    "mctsb200_ctx_create",
    "mctsb200_ctx_destroy",
    "mctsb200_ctx_create_multi",
    "mctsb200_ctx_device_count",
    "mctsb200_comm_unique_id",
    "mctsb200_comm_init",
    "mctsb200_root_sums_words",
    "mctsb200_root_sums_merge",
    "mctsb200_mdp_lookup",
    "mctsb200_mdp_info",
    "mctsb200_mdp_configure",
    "mctsb200_mdp_register",
    "mctsb200_mdp_register_continuous",
    "mctsb200_plugin_cache_stats",
    "mctsb200_plan",
    "mctsb200_plan_device",
    "mctsb200_plan_root_parallel",
    "mctsb200_run_episodes",
    "mctsb200_set_progress_callback",
    "mctsb200_set_timer",
    "mctsb200_root_parallel_decide",
    "mctsb200_root_stats",
    "mctsb200_root_stats_n",
    "mctsb200_tree_sizes_query",
    "mctsb200_tree_export",
    "mctsb200_clear_trees",
    "mctsb200_root_action_vectors",
    "mctsb200_tree_vis_stats",
    "mctsb200_device_log",
    "mctsb200_device_cos",
    "mctsb200_device_philox",
)


PROGRESS_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int64, C.c_int64)  # mctsb200_progress_fn
TIMER_FN = C.CFUNCTYPE(C.c_double, C.c_void_p)                         # mctsb200_timer_fn


def declare(lib: C.CDLL) -> C.CDLL:
    """Attach argtypes/restypes for every entry point of include/mctsb200.h."""
    vp, i32, i64, u32, u64, dbl, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double, C.c_size_t
    P = C.POINTER
    lib.mctsb200_abi_version.argtypes = []
    lib.mctsb200_abi_version.restype = i32
    lib.mctsb200_last_error.argtypes = []
    lib.mctsb200_last_error.restype = C.c_char_p
    lib.mctsb200_ctx_create.argtypes = [i32, vp, P(vp)]
    lib.mctsb200_ctx_destroy.argtypes = [vp]
    lib.mctsb200_ctx_create_multi.argtypes = [P(i32), i32, P(vp)]
    lib.mctsb200_ctx_device_count.argtypes = [vp, P(i32)]
    lib.mctsb200_comm_unique_id.argtypes = [vp, sz]
    lib.mctsb200_comm_init.argtypes = [vp, vp, sz, i32, i32]
    lib.mctsb200_root_sums_words.argtypes = [vp, vp, i32, P(i32)]
    lib.mctsb200_root_sums_merge.argtypes = [vp, i32, vp, vp]
    lib.mctsb200_mdp_lookup.argtypes = [C.c_char_p, P(i32)]
    lib.mctsb200_mdp_info.argtypes = [i32, P(i32), P(i64)]
    lib.mctsb200_mdp_configure.argtypes = [vp, i32, vp, sz]
    lib.mctsb200_mdp_register.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, i32, i32, i32, P(i32)]
    lib.mctsb200_mdp_register_continuous.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, i32, i32, P(i32)]
    lib.mctsb200_plugin_cache_stats.argtypes = [P(i64), P(i64)]
    lib.mctsb200_plan.argtypes = [vp, P(SolverCfg), i32, vp, i64, sz, i64, vp, vp, vp, P(PlanStats)]
    lib.mctsb200_plan_device.argtypes = [vp, P(SolverCfg), i32, vp, i64, sz, i64, vp, vp, vp, P(PlanStats)]
    lib.mctsb200_plan_root_parallel.argtypes = [vp, P(SolverCfg), i32, vp, sz, i64, i64, vp, vp, vp, P(PlanStats)]
    lib.mctsb200_run_episodes.argtypes = [vp, P(SolverCfg), i32, vp, i64, sz, i32, u64, i64, vp, vp, vp, P(PlanStats)]
    lib.mctsb200_set_progress_callback.argtypes = [vp, PROGRESS_FN, vp, i64]
    lib.mctsb200_set_timer.argtypes = [vp, TIMER_FN, vp]
    lib.mctsb200_root_parallel_decide.argtypes = [vp, vp, i32, dbl, P(i32), P(dbl), vp]
    lib.mctsb200_root_stats.argtypes = [vp, i64, P(i32), vp, vp, vp, P(i64)]
    lib.mctsb200_root_stats_n.argtypes = [vp, i64, i64, P(i32), vp, vp, vp, P(i64)]
    lib.mctsb200_tree_sizes_query.argtypes = [vp, i64, P(TreeSizes)]
    lib.mctsb200_tree_export.argtypes = [vp, i64, P(TreeBuffers)]
    lib.mctsb200_clear_trees.argtypes = [vp]
    lib.mctsb200_root_action_vectors.argtypes = [vp, i64, i64, vp]
    lib.mctsb200_tree_vis_stats.argtypes = [vp, i64, vp, vp, vp, i64, P(i64)]
    lib.mctsb200_device_log.argtypes = [vp, vp, i64, vp]
    lib.mctsb200_device_cos.argtypes = [vp, vp, i64, vp]
    lib.mctsb200_device_philox.argtypes = [vp, u64, u64, u32, u32, i32, vp]
    for name in EXPORTED_SYMBOLS:
        if name != "mctsb200_last_error":
            getattr(lib, name).restype = i32
    return lib
