/*
 * mctsb200.h -- C ABI of libmctsb200.so: the B200 (sm_100a) search loop behind MCTS.jl's
 * MCTSSolver/DPWSolver -> solve -> action/action_info API.
 *
 * What this boundary replaces in the reference (all paths relative to /root/reference):
 *   - the body of build_tree                      src/vanilla.jl:209-237   (vanilla UCT)
 *   - the search block of DPW action_info         src/dpw.jl:29-67         (double progressive widening)
 *   - everything those call downward: simulate (src/vanilla.jl:240-276, src/dpw.jl:80-154),
 *     insert_node! (src/vanilla.jl:292-307), insert_state_node!/insert_action_node!
 *     (src/dpw_types.jl:162-186), best_sanode_UCB (src/vanilla.jl:354-386, src/dpw.jl:179-199),
 *     best_sanode_Q / best_sanode (src/vanilla.jl:339-349, src/dpw.jl:163-173),
 *     estimate_value(::SolvedRolloutEstimator) (src/domain_knowledge.jl:65-73) with the POMDPTools
 *     RolloutSimulator loop, numeric init_Q/init_N/estimate_value (src/domain_knowledge.jl:10,82,91),
 *     RandomActionGenerator (src/action_gen.jl:11-13) and the model's gen/isterminal/actions.
 * The solver structs, solve(), action()/action_info() wrappers, default_action handling and tree
 * visualisation stay on the host (Julia glue: mcts.jl_b200/julia/MCTSB200.jl; Python mirror used by the
 * tests: mcts.jl_b200/solvers.py).  INTEGRATION.md shows the ccall stubs.
 *
 * Conventions
 *   - every entry point is extern "C", takes plain pointers and sizes, never throws, never aborts;
 *   - return value: 0 = MCTSB200_OK, negative = error; text via mctsb200_last_error() (thread local);
 *   - the caller owns every host buffer; the library owns all device memory inside a ctx;
 *   - a ctx made by mctsb200_ctx_create is bound to ONE CUDA device; a ctx made by mctsb200_ctx_create_multi owns one child
 *     ctx per device of this process and fans every call out to them (one host thread and one stream per device, roots sharded
 *     in contiguous blocks, Philox streams keyed by the global root index, so the result equals the one-device result bit
 *     for bit). With one process per GPU (torchrun, MPI) each rank keeps its own ctx, shards with `root_index_base`, and
 *     joins a communicator with mctsb200_comm_init for the one collective the path has (root-parallel search). A ctx is not
 *     re-entrant, exactly like a reference planner (src/vanilla.jl:129-135);
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails.
 *   - actions are returned as 0-based indices into actions(mdp) iteration order
 *     (SimpleGridWorld: up, down, left, right; MountainCar: -1.0, 0.0, 1.0); node ids in exported
 *     trees are 1-based in insertion order, as in MCTSTree/DPWTree.
 */
#ifndef MCTSB200_H
#define MCTSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCTSB200_ABI_VERSION 2

/* ---- status codes -------------------------------------------------------------------------- */
enum {
    MCTSB200_OK = 0,
    MCTSB200_ERR_INVALID_ARG = -1,
    MCTSB200_ERR_UNSUPPORTED = -2,   /* option not expressible on the device (host closures etc.)    */
    MCTSB200_ERR_TERMINAL_ROOT = -3, /* DPW called on a terminal state (src/dpw.jl:22-27)             */
    MCTSB200_ERR_CAPACITY = -4,      /* a per-tree table overflowed                                    */
    MCTSB200_ERR_CUDA = -5,          /* CUDA runtime error or no device                                */
    MCTSB200_ERR_STATE = -6          /* call sequence error (e.g. tree query before plan)              */
};

/* ---- built-in MDP plugins ------------------------------------------------------------------- */
enum { MCTSB200_MDP_GRIDWORLD = 0, MCTSB200_MDP_MOUNTAINCAR = 1, MCTSB200_MDP_GAUSSIAN_BOX = 2, MCTSB200_MDP_COUNT = 3 };

#define MCTSB200_GW_MAX_SPECIAL 32

/* SimpleGridWorld (POMDPModels.jl; restated in SURVEY.md 8c). State = two int64 (x, y), 1-based;
 * the terminal state is (-1,-1). */
typedef struct {
    int32_t nx, ny;         /* size, default (10,10); nx*ny + 1 <= 65535                          */
    double tprob;           /* default 0.7                                                        */
    double discount;        /* default 0.95                                                       */
    int32_t n_rewards;      /* rewards Dict entries                                               */
    int32_t n_terminate;    /* terminate_from Set entries (default: the reward cells)             */
    int32_t reward_x[MCTSB200_GW_MAX_SPECIAL], reward_y[MCTSB200_GW_MAX_SPECIAL];
    double reward_v[MCTSB200_GW_MAX_SPECIAL];
    int32_t term_x[MCTSB200_GW_MAX_SPECIAL], term_y[MCTSB200_GW_MAX_SPECIAL];
} mctsb200_gridworld_params;

/* MountainCar (POMDPModels.jl). State = two float64 (x, v). */
typedef struct {
    double discount; /* 0.99  */
    double cost;     /* -1.0  */
    double jackpot;  /* 100.0 */
} mctsb200_mountaincar_params;

/* A continuous-state, CONTINUOUS-ACTION model: the MDP of the reference's test/discussion_514.jl (states(m) = Box, actions(m) = Box,
 * transition(m, s, a) = MvNormal(s, Diagonal(var)), discount 0.9, reward 0) with the drift and the reward that make a search over it
 * worth comparing:  sp = s + drift * (a_1, ..., a_AD, 0, ...) + sqrt(var) * z,  z ~ N(0, I) (Box-Muller on the tree's Philox stream with
 * the deterministic log / cos / sin of mctsb200_detmath.h);  r = -cost * |sp|^2;  actions(m, s) = Box(a_lo, a_hi), sampled uniformly
 * by next_action (RandomActionGenerator, src/action_gen.jl:11-13: rand(rng, actions(mdp, s))). drift = 0, cost = 0 is the test's
 * model. State = MCTSB200_BOX_STATE_DOUBLES float64 (the test's SVector{3}, padded with one 0.0), action = MCTSB200_BOX_ACTION_DOUBLES
 * float64. No terminal states. Only DPW can plan for it (vanilla UCT enumerates actions(mdp, s), src/vanilla.jl:297). */
#define MCTSB200_BOX_STATE_DOUBLES 4
#define MCTSB200_BOX_ACTION_DOUBLES 2
typedef struct {
    double discount;                               /* 0.9 */
    double a_lo[MCTSB200_BOX_ACTION_DOUBLES], a_hi[MCTSB200_BOX_ACTION_DOUBLES]; /* [-5,-5], [5,5] */
    double var[3];                                 /* diagonal covariance, [0.1, 0.1, 0.1] */
    double drift;                                  /* 0.0 in the test */
    double cost;                                   /* 0.0 in the test */
} mctsb200_gaussian_box_params;

/* ---- solver configuration -------------------------------------------------------------------- */
enum { MCTSB200_KIND_UCT = 0, MCTSB200_KIND_DPW = 1 };
enum { MCTSB200_EST_ROLLOUT = 0, MCTSB200_EST_CONSTANT = 1, MCTSB200_EST_TABLE = 2 };
enum {
    MCTSB200_POLICY_RANDOM = 0,    /* RandomPolicy: rand(rng, actions(mdp, s))                      */
    MCTSB200_POLICY_CONSTANT = 1,  /* FunctionPolicy(s -> actions[param[0]])  (test/options.jl:29)  */
    MCTSB200_POLICY_GW_SEEK = 2,   /* notebook SeekTarget(GWPos(param[0], param[1]))                */
    MCTSB200_POLICY_MC_ENERGY = 3, /* MountainCar: push in the direction of motion (v<0 ? -1 : +1)  */
    MCTSB200_POLICY_TABLE = 4      /* FunctionPolicy(f) for any state-only f on a dense-state MDP, tabulated by the host:
                                      action index = policy_table[state_index] (src/domain_knowledge.jl:57)             */
};

/* run-time registered MDP plugins (mctsb200_mdp_register below) */
#define MCTSB200_MDP_PLUGIN_BASE 100
#define MCTSB200_POLICY_PLUGIN 100 /* rollout_policy ids >= this are passed to the plugin's policy_action            */
#define MCTSB200_PLUGIN_PARAM_BYTES 256

/* Flat mirror of MCTSSolver (src/vanilla.jl:28-62) and DPWSolver (src/dpw_types.jl:45-100).
 * Fields that hold host closures in the reference have no representation here. What runs BETWEEN launches is a host hook (callback,
 * show_progress: mctsb200_set_progress_callback; timer: mctsb200_set_timer); what looks at a state of a dense-state MDP only is a table
 * the host fills (init_q_table, init_n_table, value_table, policy_table, next_action_table); everything else (reset_callback,
 * closures over hashed states or tree nodes) the host glue must refuse (MCTSB200_ERR_UNSUPPORTED) -- there is no CPU fallback. */
typedef struct {
    int32_t struct_size;          /* = sizeof(mctsb200_solver_cfg), ABI guard                      */
    int32_t kind;                 /* MCTSB200_KIND_*                                               */
    int64_t n_iterations;         /* (a max_time search runs at most min(2^24 - 2, 2 * 10^9 / depth) of them.) Visit counts are 32-bit on the device (the reference's are Int64): a
                                     call whose bound -- visits already in kept trees + n_iterations * depth -- passes 2 * 10^9
                                     is refused with MCTSB200_ERR_CAPACITY (clear_tree! and start over), never wrapped         */
    int32_t depth;
    int32_t rollouts_per_leaf;    /* 1 = reference semantics; K>1 averages K rollouts per leaf     */
    double exploration_constant;  /* >= 0                                                          */
    double max_time;              /* seconds; +inf = off. Checked between iteration chunks.        */
    /* DPW only */
    double k_action, alpha_action, k_state, alpha_state;
    int32_t enable_action_pw, enable_state_pw, check_repeat_state, check_repeat_action;
    int32_t keep_tree;            /* DPW keep_tree / vanilla reuse_tree: the next plan call with the same ctx, solver kind,
                                     MDP and number of roots continues tree i for root i (src/vanilla.jl:213-217,
                                     src/dpw.jl:31-37) until mctsb200_clear_trees. A dense-state MDP with merged states has a
                                     bounded node table; any other tree (hashed states: MountainCar, run-time plugins; unmerged
                                     states) grows by up to n_iterations state nodes per call and gets room for
                                     keep_tree_calls calls, after which a tree reports MCTSB200_ERR_CAPACITY (clear and start
                                     over)                                                                              */
    int32_t lanes_per_tree;       /* 0 = auto; else 1,2,4,8,16,32 lanes cooperate on one tree      */
    /* node initialisers and leaf estimator */
    double init_Q;
    int64_t init_N;
    int32_t estimate_kind;        /* MCTSB200_EST_*                                                */
    int32_t rollout_policy;       /* MCTSB200_POLICY_*                                             */
    double estimate_const;
    int32_t rollout_policy_param[4];
    int32_t rollout_max_depth;    /* RolloutEstimator.max_depth; -1 = remaining depth              */
    int32_t keep_tree_calls;      /* plan calls a kept tree of an unbounded model has room for (0 = 8)             */
    double rollout_eps;
    uint64_t seed;                /* Philox key                                                    */
    /* Optional dense tables (host pointers, may be NULL; only for MDPs with a dense state index):
     * device-expressible stand-ins for Function-valued hooks. Index = state_index*n_actions + a,
     * resp. state_index. When non-NULL they override init_Q / init_N / estimate_const. */
    const double* init_q_table;
    const int64_t* init_n_table;
    const double* value_table;    /* used when estimate_kind == MCTSB200_EST_TABLE                 */
    const int32_t* policy_table;  /* [n_states] action indices, used when rollout_policy == MCTSB200_POLICY_TABLE */
    int32_t enable_tree_vis;      /* MCTSSolver.enable_tree_vis: record (action node => next state node) visit counts, the
                                     tree's _vis_stats (src/vanilla.jl:268-270, 328-334), for D3Tree(::MCTSTree); vanilla UCT on
                                     models with bounded branching (max_branch >= 1) only                                */
    int32_t reserved1;
    const int32_t* next_action_table; /* DPWSolver.next_action as a state-only function f(mdp, s, snode) (src/action_gen.jl:14,
                                     test/options.jl:50), tabulated by the host over a dense state space: the action index action
                                     widening proposes at a state, [n_states]; draws no random number. NULL = RandomActionGenerator
                                     (src/action_gen.jl:11-13). Host pointer; dense-state MDPs only */
} mctsb200_solver_cfg;

/* Work counters of one plan call (sums over all trees of the call). "Simulation" = one iteration of
 * src/vanilla.jl:229-235 / src/dpw.jl:48-56; "rollout step" = one iteration of the RolloutSimulator
 * loop. algorithmic_bytes applies SURVEY.md 8(d): UCT 144 B/level + 112 B/inserted state node;
 * DPW (144 + 24*C) B/level + 56 B/state node + 72 B/action node. */
typedef struct {
    int64_t n_trees;
    int64_t n_simulations;
    int64_t n_tree_levels;       /* selection steps (levels that reach best_sanode_UCB)            */
    int64_t n_dpw_children_seen; /* sum over DPW levels of C = children present at selection       */
    int64_t n_state_inserts;
    int64_t n_action_inserts;
    int64_t n_rollouts;
    int64_t n_rollout_steps;
    int64_t n_transition_samples; /* DPW: draws from transitions[sanode] (src/dpw.jl:140)          */
    int64_t n_select_exact;      /* selections that needed the exact div/sqrt score (rest: table fast path) */
    int64_t algorithmic_bytes;
    double kernel_ms;            /* device time of the search kernels (CUDA events)                */
    double total_ms;             /* host wall time of the call incl. copies                         */
    int64_t h2d_bytes, d2h_bytes;
    int32_t lanes_per_tree;      /* the value actually used                                         */
    int32_t n_kernel_launches;
    int32_t iterations_done;     /* < n_iterations only when max_time stopped the search           */
    int32_t kernel_variant;      /* 0 = tile kernel (search.cuh), 1 = shared-memory lane-per-tree UCT kernel (uct_fast.cuh) */
} mctsb200_plan_stats;

typedef struct mctsb200_ctx mctsb200_ctx;

#ifndef __CUDACC_RTC__ /* (this header is also part of the run-time compiled plugin programs: types only there) */
/* ---- lifecycle --------------------------------------------------------------------------------- */
int32_t mctsb200_abi_version(void);
const char* mctsb200_last_error(void);
/* device_id: CUDA ordinal. stream: a cudaStream_t to launch on (e.g. torch's current stream), or NULL
 * for a stream owned by the ctx. */
int32_t mctsb200_ctx_create(int32_t device_id, void* stream, mctsb200_ctx** out);
int32_t mctsb200_ctx_destroy(mctsb200_ctx* ctx);

/* ---- multi-GPU (the reference has no counterpart: it plans one state per call on one thread) ----------------- */
/* One ctx over n_dev devices of this process. Every entry point below accepts it: mctsb200_mdp_configure configures all devices;
 * mctsb200_plan / mctsb200_run_episodes shard the roots / episodes in contiguous blocks (global indices = root_index_base + i, so
 * outputs and trees equal the one-device call bit for bit; no data-path collective); mctsb200_plan_root_parallel splits the
 * ensemble's trees over the devices and merges the per-device partial sums with ONE ncclAllReduce (sum, int64, 5*|A|+1 words)
 * per decision over NVLink -- the only inter-GPU traffic of the library; tree queries address tree i of the whole call.
 * Not available on a multi-device ctx: mctsb200_plan_device and d_sums_out (device pointers belong to one device), the host
 * hooks between launches. NCCL (libnccl.so.2) is loaded at the first multi-device create; one device needs none. */
int32_t mctsb200_ctx_create_multi(const int32_t* device_ids, int32_t n_dev, mctsb200_ctx** out);
int32_t mctsb200_ctx_device_count(mctsb200_ctx* ctx, int32_t* n_dev_out);
/* One process per GPU: rank 0 makes an id (128 bytes = NCCL_UNIQUE_ID_BYTES), the host distributes it by whatever it has
 * (torch.distributed broadcast, MPI_Bcast, a file), and every rank attaches its single-device ctx. Afterwards
 * mctsb200_plan_root_parallel on that ctx is a collective call: it returns the sums merged over all ranks. */
int32_t mctsb200_comm_unique_id(void* id_out, size_t bytes);
int32_t mctsb200_comm_init(mctsb200_ctx* ctx, const void* unique_id, size_t bytes, int32_t n_ranks, int32_t rank);

/* ---- MDP plugins ------------------------------------------------------------------------------- */
int32_t mctsb200_mdp_lookup(const char* name, int32_t* mdp_id); /* "SimpleGridWorld", "MountainCar" */
int32_t mctsb200_mdp_info(int32_t mdp_id, int32_t* n_actions, int64_t* state_bytes);
int32_t mctsb200_mdp_configure(mctsb200_ctx* ctx, int32_t mdp_id, const void* pod_params, size_t bytes);

/* Run-time registration of an MDP plugin (SURVEY.md 8b "MDP plugin (device side)", 8f rank 3): the device-side
 * counterpart of defining POMDPs.gen / isterminal / actions / discount for a new model in Julia (call sites
 * src/vanilla.jl:247,258,297, src/dpw.jl:22,87,107,122) and of a custom rollout policy (src/domain_knowledge.jl:56-57).
 * `cuda_source` is CUDA C++ text defining `struct <plugin_struct>` with the members documented in
 * mcts.jl_b200/csrc/custom.cuh (Params beginning with `double discount`, is_terminal, gen, policy_action); states are
 * `state_doubles` float64 values (2 or 4: pad unused ones with 0.0), actions are indices 0..n_actions-1 (2, 3 or 4),
 * `max_branch` bounds the distinct successors of one (s, a): 0 = unbounded (always correct; what a host wrapper should
 * default to), 1 = deterministic transitions -- an explicit promise: the kernels then reuse the first sampled successor of an
 * action node instead of calling gen again, so a stochastic gen registered with 1 would get silently narrow trees. The tile kernel is compiled around the plugin with NVRTC for sm_100a; compiler errors are returned through
 * mctsb200_last_error(). Needs no GPU (the cubin is loaded at the first plan call). The id is process-wide; every ctx
 * configures its own parameters with mctsb200_mdp_configure(ctx, id, &Params, sizeof(Params)). Registering the same
 * name again with a different definition replaces it (same id). mctsb200_mdp_lookup / mctsb200_mdp_info know the id.
 * Plugin MDPs run on the tile kernel: hashed states, so no dense tables (init_q_table, value_table, policy_table); kept trees get
 * room for keep_tree_calls calls. */
int32_t mctsb200_mdp_register(const char* name, const char* cuda_source, const char* plugin_struct, int32_t state_doubles,
                              int32_t n_actions, int32_t max_branch, int32_t* mdp_id_out);
/* The same for a CONTINUOUS action space (actions(mdp) = a Box of `action_doubles` float64 coordinates, 1 to 4; the reference's
 * test/discussion_514.jl with a model of the caller's own): the struct defines is_terminal, next_action (the action DPW's action
 * widening proposes: next_action(gen, mdp, s, snode), src/dpw.jl:98 -- for the default RandomActionGenerator a draw from the Box,
 * src/action_gen.jl:11-13), gen(params, s, a, rng, sp, r) with `a` the action vector, and policy_action for rollout policies
 * >= MCTSB200_POLICY_PLUGIN (mcts.jl_b200/csrc/custom.cuh). Successors are unbounded (max_branch = 0). Such a model plans like
 * MCTSB200_MDP_GAUSSIAN_BOX: DPW with action widening only, actions come back through mctsb200_root_action_vectors and
 * tree_buffers.a_label_vectors, mctsb200_mdp_info reports n_actions = 0. The DPW program is compiled at registration. */
int32_t mctsb200_mdp_register_continuous(const char* name, const char* cuda_source, const char* plugin_struct, int32_t state_doubles,
                                         int32_t action_doubles, int32_t* mdp_id_out);
/* Compiled plugin programs are cached on disk, keyed by a hash of the generated program, the library's kernel sources, the compiler
 * options and the CUDA version ($MCTSB200_CACHE_DIR, default ~/.cache/mctsb200; MCTSB200_CACHE=0 disables): a later process
 * registers a known plugin without running the compiler. Counts for this process: programs read from the cache / compiled. */
int32_t mctsb200_plugin_cache_stats(int64_t* hits_out, int64_t* misses_out);

/* ---- planning: replaces build_tree (src/vanilla.jl:209-237) / the DPW loop (src/dpw.jl:29-67) ---- */
/* Plans for n_roots independent root states (one tree each). roots: n_roots*state_bytes host bytes.
 * root_index_base: global index of roots[0] -- Philox streams are keyed by global root index so a
 * sharded run returns the same trees as an unsharded one. Outputs (host): action_idx_out[n_roots]
 * = index of best_sanode_Q's action, best_q_out[n_roots] = its Q. status_out (may be NULL): per-root
 * status (MCTSB200_OK / MCTSB200_ERR_TERMINAL_ROOT / MCTSB200_ERR_CAPACITY). Trees stay resident in the ctx until the
 * next plan call or ctx_destroy. */
int32_t mctsb200_plan(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* roots,
                      int64_t n_roots, size_t state_bytes, int64_t root_index_base, int32_t* action_idx_out,
                      double* best_q_out, int32_t* status_out, mctsb200_plan_stats* stats_out);

/* Same, but roots/action/best_q/status are DEVICE pointers (e.g. torch tensors); nothing is copied and
 * the call returns after the kernels are enqueued and counters read back. */
int32_t mctsb200_plan_device(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, int32_t mdp_id,
                             const void* d_roots, int64_t n_roots, size_t state_bytes, int64_t root_index_base,
                             int32_t* d_action_idx_out, double* d_best_q_out, int32_t* d_status_out,
                             mctsb200_plan_stats* stats_out);

/* Root-parallel search of ONE root (BASELINE config 5): n_trees independent trees of the same root
 * (Philox tree index = tree_index_base + t), n_iterations each. Writes sum_n[a] = sum_t N_t(a) and
 * sum_nq[a] = sum_t N_t(a)*Q_t(a) (host, n_actions each), then the host calls mctsb200_root_parallel_decide.
 * The sums are accumulated in INTEGER arithmetic: every once-rounded double product N*Q is taken as a multiple of 2^-64
 * (rounded toward zero; |N*Q| < 2^40, else the sum is NaN) and added in 128-bit two's complement, and the total is rounded to
 * double once. They therefore do not depend on the order of the trees, on the number of devices or on the reduction order of
 * the collective: on a multi-device ctx, or on a ctx that joined a communicator (mctsb200_comm_init), the partial sums of
 * all devices / ranks are merged by ONE ncclAllReduce inside this call and every rank receives the same merged sums, bit-equal to
 * a single-device call with the same trees (n_trees may be 0 on a rank of a communicator). d_sums_out: optional DEVICE
 * buffer of 2*n_actions doubles that receives [sum_n | sum_nq] as returned. */
int32_t mctsb200_plan_root_parallel(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, int32_t mdp_id,
                                    const void* root_state, size_t state_bytes, int64_t n_trees,
                                    int64_t tree_index_base, double* sum_n_out, double* sum_nq_out,
                                    double* d_sums_out, mctsb200_plan_stats* stats_out);
/* For hosts that merge with their own transport (MPI, gloo): the integer words behind the sums of this ctx's last
 * root-parallel call -- [n[A] | four 32-bit limbs per action in int64 | inexact flag], 5*|A|+1 words; add them across ranks as
 * int64 (any order) and convert with mctsb200_root_sums_merge. Two-call pattern: words_out NULL returns the count. */
int32_t mctsb200_root_sums_words(mctsb200_ctx* ctx, int64_t* words_out, int32_t capacity, int32_t* n_words_out);
int32_t mctsb200_root_sums_merge(const int64_t* words, int32_t n_actions, double* sum_n_out, double* sum_nq_out);
/* Q_a = sum_nq/sum_n (a with sum_n == 0 keeps init_Q); action = first argmax, as best_sanode_Q. */
int32_t mctsb200_root_parallel_decide(const double* sum_n, const double* sum_nq, int32_t n_actions,
                                      double init_q, int32_t* action_idx_out, double* best_q_out, double* q_out);

/* ---- host hooks between launches -------------------------------------------------------------------- */
/* Coarse stand-ins for MCTSSolver.callback (src/vanilla.jl:26,231: called after every iteration) and solver.timer
 * (src/vanilla.jl:25,226-232; src/dpw.jl:44-52). Host functions cannot run inside a kernel, but they can run between
 * launches: with a progress function set, every plan call of this ctx runs in chunks of `every` iterations
 * (every = 1 reproduces the reference's per-iteration callback exactly, at one launch per iteration) and calls
 * progress(user, iterations_done, n_iterations) after each chunk; a non-zero return ends the search. The tree
 * queries below are valid inside the callback. timer(user) -> seconds replaces the monotonic clock that max_time is
 * measured with. NULL clears either hook. */
typedef int32_t (*mctsb200_progress_fn)(void* user, int64_t iterations_done, int64_t n_iterations);
typedef double (*mctsb200_timer_fn)(void* user);
int32_t mctsb200_set_progress_callback(mctsb200_ctx* ctx, mctsb200_progress_fn fn, void* user, int64_t every);
int32_t mctsb200_set_timer(mctsb200_ctx* ctx, mctsb200_timer_fn fn, void* user);

/* ---- episode driver -------------------------------------------------------------------------------- */
/* Replaces simulate(RolloutSimulator(max_steps = max_steps), mdp, planner, s0) -- the loop that
 * bench/non-terminal_gw.jl:18,31, bench/non-terminal_gw_dpw.jl and test/performance.jl:33-40 time -- for n_episodes
 * start states in lockstep, without leaving the device between environment steps: at step t every running episode
 * plans from its current state (one mctsb200_plan over the whole batch with the Philox key cfg->seed + t; with
 * keep_tree the trees follow their episodes), then the environment moves with gen(mdp, s, a, rng), rng = Philox stream
 * (env_seed, episode_index_base + i, t, 0). An episode ends in a terminal state or after max_steps steps.
 * Outputs (host, n_episodes each; may be NULL): the discounted return sum_t gamma^t r_t, the number of steps taken and
 * the last state. stats_out: work counters summed over all plan calls. */
int32_t mctsb200_run_episodes(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* initial_states,
                              int64_t n_episodes, size_t state_bytes, int32_t max_steps, uint64_t env_seed,
                              int64_t episode_index_base, double* return_out, int32_t* steps_out, void* final_states_out,
                              mctsb200_plan_stats* stats_out);

/* ---- tree queries (info[:tree], value(), ranked_actions) ------------------------------------------ */
/* Children of the root of tree root_i, in child order. Buffers hold >= n_actions entries (mctsb200_mdp_info). A continuous action
 * space has no such bound: there this entry point only answers the count query (all three arrays NULL -> n_children, total_n) and
 * refuses to write arrays (MCTSB200_ERR_UNSUPPORTED); use mctsb200_root_stats_n. */
int32_t mctsb200_root_stats(mctsb200_ctx* ctx, int64_t root_i, int32_t* n_children, int32_t* action_idx,
                            int64_t* n, double* q, int64_t* total_n);
/* The same with the arrays' capacity stated: at most `capacity` entries are written, n_children is the full count (two-call pattern:
 * NULL arrays first). For every kind of tree. */
int32_t mctsb200_root_stats_n(mctsb200_ctx* ctx, int64_t root_i, int64_t capacity, int32_t* n_children, int32_t* action_idx, int64_t* n,
                              double* q, int64_t* total_n);

typedef struct {
    int64_t n_state_nodes;
    int64_t n_action_nodes;
    int64_t n_transitions;   /* DPW: distinct (sanode, spnode) records; 0 for UCT */
    int64_t state_bytes;
    int64_t action_doubles;  /* 0: actions are indices into actions(mdp) (a_labels); > 0: actions are vectors (a_label_vectors) */
} mctsb200_tree_sizes;

/* Caller-allocated arrays sized from mctsb200_tree_sizes (two-call pattern). Field names follow
 * MCTSTree (src/vanilla.jl:64-94) and DPWTree (src/dpw_types.jl:123-159). Any pointer may be NULL. */
typedef struct {
    /* per state node */
    int64_t* total_n;        /* [n_state_nodes]                                                    */
    void* s_labels;          /* [n_state_nodes * state_bytes]                                      */
    int64_t* child_offsets;  /* [n_state_nodes + 1] CSR into child_ids                             */
    int64_t* child_ids;      /* [n_action_nodes] 1-based action-node ids                           */
    /* per action node, 1-based id = position + 1 */
    int64_t* n;              /* [n_action_nodes]                                                   */
    double* q;               /* [n_action_nodes]                                                   */
    int32_t* a_labels;       /* [n_action_nodes] action index                                      */
    /* DPW only */
    int64_t* n_a_children;   /* [n_action_nodes]                                                   */
    int64_t* trans_offsets;  /* [n_action_nodes + 1] CSR into the arrays below                     */
    int64_t* trans_spnode;   /* [n_transitions] 1-based state-node id                              */
    double* trans_r;         /* [n_transitions]                                                    */
    int64_t* trans_count;    /* [n_transitions] multiplicity in transitions[sanode]                */
    /* continuous action spaces (action_doubles > 0): the action-node labels are vectors; a_labels is then filled with -1 */
    double* a_label_vectors; /* [n_action_nodes * action_doubles]                                  */
} mctsb200_tree_buffers;

int32_t mctsb200_tree_sizes_query(mctsb200_ctx* ctx, int64_t root_i, mctsb200_tree_sizes* sizes);
int32_t mctsb200_tree_export(mctsb200_ctx* ctx, int64_t root_i, const mctsb200_tree_buffers* buffers);
int32_t mctsb200_clear_trees(mctsb200_ctx* ctx); /* clear_tree! (src/vanilla.jl:148, src/dpw.jl:6-8) */
/* tree._vis_stats of a vanilla tree planned with enable_tree_vis (src/vanilla.jl:328-334): for every (action node, next state node)
 * edge the search walked, how often -- what D3Tree(::MCTSTree) (src/visualization.jl:52-122) draws below an action node. Entries
 * sorted by (said, spid), ids 1-based. Two-call pattern: NULL arrays return the count in n_out. */
int32_t mctsb200_tree_vis_stats(mctsb200_ctx* ctx, int64_t root_i, int64_t* said, int64_t* spid, int64_t* count, int64_t capacity, int64_t* n_out);
/* Continuous action spaces: the action best_sanode chose for roots [first_root, first_root + n) of the last plan call, as vectors
 * (n * action_doubles float64 -- MCTSB200_BOX_ACTION_DOUBLES for the built-in model, what a plugin was registered with, also
 * reported by mctsb200_tree_sizes.action_doubles; the action_idx_out of mctsb200_plan then holds the chosen child's position among the root's
 * children, -1 if it has none). a_labels[sanode] of src/dpw.jl:65. */
int32_t mctsb200_root_action_vectors(mctsb200_ctx* ctx, int64_t first_root, int64_t n, double* out);

/* ---- numerical contract helpers (exposed so hosts/tests can check device == host bit-for-bit) ---- */
/* Evaluates det_log / det_cos (mctsb200_detmath.h) ON THE DEVICE for n host inputs. */
int32_t mctsb200_device_log(mctsb200_ctx* ctx, const double* x, int64_t n, double* out);
int32_t mctsb200_device_cos(mctsb200_ctx* ctx, const double* x, int64_t n, double* out);
/* First 4*n_blocks Philox4x32-10 words of stream (seed, tree, iteration, stream_id), from the device. */
int32_t mctsb200_device_philox(mctsb200_ctx* ctx, uint64_t seed, uint64_t tree, uint32_t iteration,
                               uint32_t stream_id, int32_t n_blocks, uint32_t* out);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif

/*
 * Random-number contract (shared by the library and the test oracle; the reference's single
 * MersenneTwister stream -- src/vanilla.jl:53,141, src/domain_knowledge.jl:53 -- is replaced by
 * counter-based Philox4x32-10 streams, so stochastic runs compare with Julia only statistically):
 *   key     = (seed lo32, seed hi32)
 *   counter = (block, iteration, tree lo32, (tree hi16 << 16) | stream_id)
 *   stream 0          : tree descent of that iteration (next_action, gen, transition sampling)
 *   stream 1 + k      : k-th rollout of that iteration (k = 0 for the reference's single rollout)
 *   words are consumed in order x0,x1,x2,x3 of block 0, then block 1, ...
 *   uniform_int(n)    = (word * n) >> 32           (rand(rng, 1:n) stand-in)
 *   uniform01()       = word * 2^-32               (rand(rng) stand-in)
 */
#endif /* MCTSB200_H */
