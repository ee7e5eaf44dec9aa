// capi.cu -- host side of libmctsb200.so: the C ABI of include/mctsb200.h.
//
// Owns device memory (node tables, lookup tables, counters), validates solver options the way the Julia
// glue needs them (unsupported -> error, never a CPU fallback), sizes the per-tree tables, picks the tile
// width, launches the search kernels (search.cuh) in iteration chunks (max_time), runs the root
// decision kernel and exports trees in the reference's MCTSTree / DPWTree vector layout.
#ifndef MCTSB200_HOST_EMU
#include <cuda_runtime.h>
#endif

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "search.cuh"
#include "dpw_fast.cuh"
#include "custom.cuh"
#ifndef MCTSB200_HOST_EMU
#include <dlfcn.h>
#include <nccl.h>

#include <map>

#include "plugin_rt.h"
#endif
#include <thread>

using namespace mctsb200;

namespace {

thread_local std::string g_err = "";

int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) return fail(MCTSB200_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));  \
    } while (0)

// device buffer that only grows
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Storage {
    bool valid = false;
    bool direct = false;  // rows indexed by the state's dense index (uct_fast.cuh); Row::key holds the insertion sequence number
    bool keep = false;    // the trees were built with reuse_tree / keep_tree and may be continued by the next plan call
    long long visits = 0; // upper bound of the visits any node of the kept trees has received so far
    double q_bound = 0;   // bound of |Q| in the kept trees (DevCfg::q_bound of the calls that built them)
    double q_bound_tile = 0;
    int kind = 0, mdp_id = 0;
    long long n_trees = 0;
    Geometry geo{};
    size_t row_bytes = 0, aux_bytes = 0, map_entry_bytes = 0;
    DevBuf hdr, rows, aux, trans, map, path;
    DevBuf plog;        // push-log arena of models with unbounded branching (search.cuh: kPushLog)
    DevBuf anodes;      // action-node table of continuous-action trees (search.cuh: dpw_box_iteration)
    DevBuf vis;         // _vis_stats edges of vanilla trees planned with enable_tree_vis
    int vis_branch = 0;
    DevBuf action_vec;  // the decisions' action vectors of the last plan call (continuous action spaces)
    int action_doubles = 0;  // ... and their width (0: the last plan call had a finite action set)
    DevBuf rows_fast;   // working table of uct_fast.cuh / dpw_fast.cuh (interleaved); `rows` receives the Row<M> table after the search
    DevBuf trans_fast;  // dpw_fast.cuh: per-slot transition multisets (interleaved)
    // the fast kernels' tables are converted to Row / Aux / TransRec only when a tree query needs them
    int32_t (*finalize_fn)(mctsb200_ctx*) = nullptr;
    long long fin_slots = 0;
    bool fin_order = false;
};

}  // namespace

struct mctsb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    bool gw_ok = false, mc_ok = false;
    mctsb200_gridworld_params gw_host{};
    GridWorldDev::Params gw_dev{};
    DevBuf gw_reward, gw_term, gw_fast;
    MountainCarDev::Params mc_dev{};
    bool box_ok = false;
    GaussianBoxDev::Params box_dev{};

    Storage st;
    DevBuf counters, log_tab, sqrtlog_tab, rsqrt_tab, nmin_state, init_q, init_n, value_tab, io_roots, io_action, io_bestq, io_status, sums;
    DevBuf order, order_bins, policy_tab, nmin_action_tab, next_action_tab;
    double gw_reward_max = 0.0;
    bool gw_fast_det = false;
    // host hooks run between launches (mctsb200_set_progress_callback / mctsb200_set_timer)
    mctsb200_progress_fn progress_fn = nullptr;
    void* progress_user = nullptr;
    long long progress_every = 0;
    mctsb200_timer_fn timer_fn = nullptr;
    void* timer_user = nullptr;
    DevBuf ep_states, ep_disc, ep_ret, ep_steps, ep_active, ep_count;
    int log_tab_n = 0;
    // multi-GPU (mctsb200_ctx_create_multi / mctsb200_comm_init): a group ctx owns one child ctx per device and fans every call
    // out to them, one host thread per device; a ctx that is a rank of a communicator merges its root-parallel partial sums
    // with ONE ncclAllReduce on its own stream
    std::vector<mctsb200_ctx*> kids;
    std::vector<long long> kid_lo;  // shard bounds of the last plan / episode call: child k holds items [kid_lo[k], kid_lo[k+1])
    void* comm = nullptr;           // ncclComm_t
    int comm_ranks = 1, comm_rank = 0;
    std::vector<long long> root_words;  // integer partial sums of the last root-parallel call (root_sums_kernel)
#ifndef MCTSB200_HOST_EMU
    // run-time registered MDP plugins (plugin_rt.cu): parameters configured for this ctx, by plugin index
    struct alignas(16) PluginParams { unsigned char bytes[kCustomParamBytes]; };
    std::map<int, PluginParams> plugin_params;
    int cur_plugin = -1;       // plugin index of the call in progress
    int cur_max_branch = 1;    // its max_branch (0 = unbounded)
#endif
};

namespace {

// ---- NCCL, loaded on first use (the library has no link-time dependency on it; a process that already holds a libnccl.so.2,
// e.g. torch's, shares that copy) -----------------------------------------------------------------------------------------
#ifndef MCTSB200_HOST_EMU
struct Nccl {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char* names[] = {std::getenv("MCTSB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !n[0]) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = "libnccl.so.2 not found (set MCTSB200_NCCL_LIB): multi-GPU root-parallel search needs NCCL";
            return false;
        }
#define MCTSB200_SYM(f)                                                    \
    f = reinterpret_cast<decltype(f)>(dlsym(handle, "nccl" #f));           \
    if (!f) {                                                              \
        error = "libnccl lacks nccl" #f;                                   \
        handle = nullptr;                                                  \
        return false;                                                      \
    }
        MCTSB200_SYM(GetUniqueId) MCTSB200_SYM(CommInitRank) MCTSB200_SYM(CommInitAll) MCTSB200_SYM(CommDestroy) MCTSB200_SYM(AllReduce)
        MCTSB200_SYM(GetErrorString)
#undef MCTSB200_SYM
        return true;
    }
};
Nccl g_nccl;
#define NCCL_TRY(expr)                                                                                         \
    do {                                                                                                       \
        ncclResult_t r__ = (expr);                                                                             \
        if (r__ != ncclSuccess) return fail(MCTSB200_ERR_CUDA, "%s: %s", #expr, g_nccl.GetErrorString(r__));  \
    } while (0)
#endif

// the int64 words of root_sums_kernel (search.cuh) -> [sum_n | sum_nq] doubles: limbs recombined in 128-bit two's complement, one
// rounding to double, exact scaling by 2^-64; NaN if a product could not be taken as a fixed-point value
void root_sums_to_double(const long long* w, int A, double* sum_n, double* sum_nq) {
    for (int a = 0; a < A; ++a) {
        sum_n[a] = (double)w[a];
        const long long* c = w + A + 4 * a;
        const __int128 v = (__int128)c[0] + (__int128)((unsigned __int128)(__int128)c[1] << 32) + (__int128)((unsigned __int128)(__int128)c[2] << 64) +
                           (__int128)((unsigned __int128)(__int128)c[3] << 96);
        sum_nq[a] = w[5 * A] ? NAN : std::ldexp((double)v, -64);
    }
}

template <class M>
struct MdpAccess;
template <>
struct MdpAccess<GridWorldDev> {
    static bool ready(const mctsb200_ctx* c) { return c->gw_ok; }
    static int mdp_id(const mctsb200_ctx*) { return MCTSB200_MDP_GRIDWORLD; }
    static long long max_branch(const mctsb200_ctx*) { return GridWorldDev::kMaxBranch; }
    static const GridWorldDev::Params& params(const mctsb200_ctx* c) { return c->gw_dev; }
    static int dense_count(const mctsb200_ctx* c) { return c->gw_dev.nx * c->gw_dev.ny + 1; }
    static int32_t check_root(const mctsb200_ctx* c, const GridWorldDev::AbiState& s) {
        const bool inb = s.x >= 1 && s.x <= c->gw_dev.nx && s.y >= 1 && s.y <= c->gw_dev.ny;
        const bool term = s.x == -1 && s.y == -1;
        return (inb || term) ? 0 : -1;
    }
    static bool policy_ok(int p) { return p == MCTSB200_POLICY_RANDOM || p == MCTSB200_POLICY_CONSTANT || p == MCTSB200_POLICY_GW_SEEK || p == MCTSB200_POLICY_TABLE; }
    static bool fast_ready(const mctsb200_ctx* c) { return c->gw_dev.fast_tt != nullptr; }
    static double reward_bound(const mctsb200_ctx* c) { return c->gw_reward_max; }  // max |reward(s, a)|
    static bool fast_deterministic(const mctsb200_ctx* c) { return c->gw_fast_det; }   // every FastEntry has one outcome
    // Deterministic transitions + a deterministic rollout policy + vanilla UCT (DPW draws random actions): trees planned
    // from equal root states are then identical whatever their Philox streams, so lanes that share a root run in lockstep.
    static bool lockstep(const mctsb200_ctx* c, const mctsb200_solver_cfg* cfg) {
        return c->gw_host.tprob == 1.0 && cfg->kind == MCTSB200_KIND_UCT && cfg->estimate_kind == MCTSB200_EST_ROLLOUT &&
               cfg->rollout_policy != MCTSB200_POLICY_RANDOM;
    }
    // a state-only deterministic rollout policy as a table over dense state indices (GridWorldDev::policy_action restated on the host)
    static bool tabulate_policy(const mctsb200_ctx* c, const mctsb200_solver_cfg* cfg, std::vector<unsigned char>& out) {
        const int nx = c->gw_dev.nx, n = c->gw_dev.ncells;
        if (cfg->rollout_policy == MCTSB200_POLICY_TABLE) {  // FunctionPolicy(f), tabulated by the caller
            out.assign((size_t)n + 1, 0);
            for (int cell = 0; cell <= n; ++cell) out[(size_t)cell] = (unsigned char)cfg->policy_table[cell];
            return true;
        }
        if (cfg->rollout_policy != MCTSB200_POLICY_GW_SEEK) return false;
        out.assign((size_t)n + 1, 0);
        for (int cell = 0; cell <= n; ++cell) {
            const int x = cell % nx + 1, y = cell / nx + 1, tx = cfg->rollout_policy_param[0], ty = cfg->rollout_policy_param[1];
            out[(size_t)cell] = (unsigned char)(tx > x ? 3 : tx < x ? 2 : ty > y ? 0 : 1);
        }
        return true;
    }
};
template <>
struct MdpAccess<MountainCarDev> {
    static bool ready(const mctsb200_ctx* c) { return c->mc_ok; }
    static int mdp_id(const mctsb200_ctx*) { return MCTSB200_MDP_MOUNTAINCAR; }
    static long long max_branch(const mctsb200_ctx*) { return MountainCarDev::kMaxBranch; }
    static const MountainCarDev::Params& params(const mctsb200_ctx* c) { return c->mc_dev; }
    static int dense_count(const mctsb200_ctx*) { return 0; }
    static int32_t check_root(const mctsb200_ctx*, const MountainCarDev::AbiState& s) {
        return (std::isfinite(s.x) && std::isfinite(s.v) && std::fabs(s.x) < 1e5) ? 0 : -1;
    }
    static bool policy_ok(int p) { return p == MCTSB200_POLICY_RANDOM || p == MCTSB200_POLICY_CONSTANT || p == MCTSB200_POLICY_MC_ENERGY; }
    static bool lockstep(const mctsb200_ctx*, const mctsb200_solver_cfg*) { return false; }
    static double reward_bound(const mctsb200_ctx* c) { return std::fmax(std::fabs(c->mc_dev.cost), std::fabs(c->mc_dev.jackpot)); }  // max |reward(s, a, sp)|
};
template <>
struct MdpAccess<GaussianBoxDev> {
    static bool ready(const mctsb200_ctx* c) { return c->box_ok; }
    static int mdp_id(const mctsb200_ctx*) { return MCTSB200_MDP_GAUSSIAN_BOX; }
    static long long max_branch(const mctsb200_ctx*) { return 1LL << 30; }  // continuous noise
    static const GaussianBoxDev::Params& params(const mctsb200_ctx* c) { return c->box_dev; }
    static int dense_count(const mctsb200_ctx*) { return 0; }
    static int32_t check_root(const mctsb200_ctx*, const GaussianBoxDev::AbiState& s) {
        for (int i = 0; i < GaussianBoxDev::kStateDoubles; ++i)
            if (!std::isfinite(s.v[i])) return -1;
        return 0;
    }
    static bool policy_ok(int p) { return p == MCTSB200_POLICY_RANDOM; }  // RandomPolicy: rand(rng, actions(mdp, s))
    static bool lockstep(const mctsb200_ctx*, const mctsb200_solver_cfg*) { return false; }
    static double reward_bound(const mctsb200_ctx*) { return NAN; }
};
#ifndef MCTSB200_HOST_EMU
// a plugin registered at run time (custom.cuh): hashed float64 states like MountainCar, parameters are the plugin's POD
template <int SD, int A>
struct MdpAccess<CustomDev<SD, A>> {
    using M = CustomDev<SD, A>;
    static bool ready(const mctsb200_ctx* c) { return c->plugin_params.count(c->cur_plugin) != 0; }
    static int mdp_id(const mctsb200_ctx* c) { return MCTSB200_MDP_PLUGIN_BASE + c->cur_plugin; }
    static long long max_branch(const mctsb200_ctx* c) { return c->cur_max_branch >= 1 ? c->cur_max_branch : (1LL << 30); }
    static const typename M::Params& params(const mctsb200_ctx* c) {
        return *reinterpret_cast<const typename M::Params*>(c->plugin_params.find(c->cur_plugin)->second.bytes);
    }
    static int dense_count(const mctsb200_ctx*) { return 0; }
    static int32_t check_root(const mctsb200_ctx*, const typename M::AbiState& s) {
        for (int i = 0; i < SD; ++i)
            if (!std::isfinite(s.v[i])) return -1;
        return 0;
    }
    static bool policy_ok(int p) { return p == MCTSB200_POLICY_RANDOM || p == MCTSB200_POLICY_CONSTANT || p >= MCTSB200_POLICY_PLUGIN; }
    static bool lockstep(const mctsb200_ctx*, const mctsb200_solver_cfg*) { return false; }
    static double reward_bound(const mctsb200_ctx*) { return NAN; }
};
// ... with a continuous action space (mctsb200_mdp_register_continuous)
template <int SD, int AD>
struct MdpAccess<CustomBoxDev<SD, AD>> {
    using M = CustomBoxDev<SD, AD>;
    static bool ready(const mctsb200_ctx* c) { return c->plugin_params.count(c->cur_plugin) != 0; }
    static int mdp_id(const mctsb200_ctx* c) { return MCTSB200_MDP_PLUGIN_BASE + c->cur_plugin; }
    static long long max_branch(const mctsb200_ctx*) { return 1LL << 30; }
    static const typename M::Params& params(const mctsb200_ctx* c) {
        return *reinterpret_cast<const typename M::Params*>(c->plugin_params.find(c->cur_plugin)->second.bytes);
    }
    static int dense_count(const mctsb200_ctx*) { return 0; }
    static int32_t check_root(const mctsb200_ctx*, const typename M::AbiState& s) {
        for (int i = 0; i < SD; ++i)
            if (!std::isfinite(s.v[i])) return -1;
        return 0;
    }
    static bool policy_ok(int p) { return p == MCTSB200_POLICY_RANDOM || p >= MCTSB200_POLICY_PLUGIN; }
    static bool lockstep(const mctsb200_ctx*, const mctsb200_solver_cfg*) { return false; }
    static double reward_bound(const mctsb200_ctx*) { return NAN; }
};
#endif

// For every (boundary class, action): the positions where the reference's sampling expression
// (GridWorldDev::sample_destination, evaluated here on the host with the same individually rounded operations)
// steps from one destination to the next as the 32-bit word w grows, and the destination on each plateau.
void gw_thresholds(double tprob, double p_other, GridWorldDev::Thresholds* out) {
    for (unsigned mask = 0; mask < 16; ++mask)
        for (int a = 0; a < 4; ++a) {
            unsigned long long first_gt[4];  // first w with destination(w) > i; 2^32 if none
            for (int i = 0; i < 4; ++i) {
                unsigned long long lo = 0, hi = 1ull << 32;
                while (lo < hi) {
                    const unsigned long long mid = lo + (hi - lo) / 2;
                    if (GridWorldDev::sample_destination(tprob, p_other, mask, a, (uint32_t)mid) > i) hi = mid;
                    else lo = mid + 1;
                }
                first_gt[i] = lo;
            }
            GridWorldDev::Thresholds& T = out[mask * 4 + a];
            for (int i = 0; i < 4; ++i) T.t[i] = 0xffffffffu;
            T.pad[0] = T.pad[1] = T.pad[2] = 0;
            int n_steps = 0;
            uint32_t codes = (uint32_t)GridWorldDev::sample_destination(tprob, p_other, mask, a, 0u);
            unsigned long long prev = 0;
            for (int i = 0; i < 4; ++i) {
                const unsigned long long v = first_gt[i];
                if (v == 0 || v >= (1ull << 32) || v == prev) continue;  // no step here (or same step as before)
                T.t[n_steps] = (uint32_t)(v - 1);                       // w > v-1  <=>  w >= v
                ++n_steps;
                codes |= (uint32_t)GridWorldDev::sample_destination(tprob, p_other, mask, a, (uint32_t)v) << (4 * n_steps);
                prev = v;
            }
            T.codes = codes;
        }
}

// min N in [0, bound] with  len <= k * N^alpha  (the reference's widening test, src/dpw.jl:95,121,
// evaluated with the host's pow exactly as written there); INT_MAX if none.
int nmin_for(double k, double alpha, int len, long long bound) {
    auto pred = [&](long long N) { return (double)len <= k * std::pow((double)N, alpha); };
    if (pred(0)) return 0;
    if (!pred(bound)) return 0x7fffffff;
    long long lo = 0, hi = bound;  // pred(lo) false, pred(hi) true; pow is monotone in N for alpha >= 0
    while (hi - lo > 1) {
        const long long mid = lo + (hi - lo) / 2;
        if (pred(mid)) hi = mid;
        else lo = mid;
    }
    while (hi > 1 && pred(hi - 1)) --hi;  // guard against a non-monotone last ulp
    return (int)hi;
}

int next_pow2(long long v) {
    long long p = 1;
    while (p < v) p <<= 1;
    return (int)p;
}

// A max_time search may ask for "unbounded" iterations -- the reference's own tests pass typemax(Int)
// (test/runtests.jl:91-118). Iteration stamps and visit counters are 32-bit here, so such a call runs at most
// min(2^24 - 2, 2e9 / depth) iterations: hours of search for one tree.
const mctsb200_solver_cfg* effective_cfg(const mctsb200_solver_cfg* cfg, mctsb200_solver_cfg& tmp) {
    if (!cfg || cfg->struct_size != (int32_t)sizeof(mctsb200_solver_cfg) || !std::isfinite(cfg->max_time)) return cfg;
    const long long cap = std::min<long long>((1LL << 24) - 2, 2000000000LL / std::max(std::min(cfg->depth, 100000), 1));
    if (cfg->n_iterations <= cap) return cfg;
    tmp = *cfg;
    tmp.n_iterations = cap;
    return &tmp;
}

int32_t validate_cfg(const mctsb200_solver_cfg* cfg, int n_actions, bool dense, bool (*policy_ok)(int), bool box_actions = false) {
    if (!cfg) return fail(MCTSB200_ERR_INVALID_ARG, "cfg is NULL");
    if (cfg->struct_size != (int32_t)sizeof(mctsb200_solver_cfg))
        return fail(MCTSB200_ERR_INVALID_ARG, "solver_cfg.struct_size %d != %zu (ABI mismatch)", cfg->struct_size, sizeof(mctsb200_solver_cfg));
    if (cfg->kind != MCTSB200_KIND_UCT && cfg->kind != MCTSB200_KIND_DPW) return fail(MCTSB200_ERR_INVALID_ARG, "unknown solver kind %d", cfg->kind);
    if (cfg->n_iterations < 0 || cfg->n_iterations > 2000000000LL) return fail(MCTSB200_ERR_INVALID_ARG, "n_iterations out of range");
    if (cfg->depth < 0 || cfg->depth > 100000) return fail(MCTSB200_ERR_INVALID_ARG, "depth out of range");
    if ((double)cfg->n_iterations * (double)std::max(cfg->depth, 1) > 2.0e9)
        return fail(MCTSB200_ERR_CAPACITY, "n_iterations*depth exceeds the 32-bit visit counters");
    if (!(cfg->exploration_constant >= 0.0) || !std::isfinite(cfg->exploration_constant))
        return fail(MCTSB200_ERR_INVALID_ARG, "exploration_constant must be finite and >= 0");
    if (cfg->rollouts_per_leaf < 1 || cfg->rollouts_per_leaf > 1024) return fail(MCTSB200_ERR_INVALID_ARG, "rollouts_per_leaf must be in [1,1024]");
    const int g = cfg->lanes_per_tree;
    if (!(g == 0 || g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32)) return fail(MCTSB200_ERR_INVALID_ARG, "lanes_per_tree must be 0,1,2,4,8,16,32");
    if (cfg->init_N < 0 || cfg->init_N > 1000000000LL) return fail(MCTSB200_ERR_INVALID_ARG, "init_N out of range");
    if (std::isnan(cfg->max_time) || cfg->max_time < 0) return fail(MCTSB200_ERR_INVALID_ARG, "max_time must be >= 0");
    if (cfg->estimate_kind < MCTSB200_EST_ROLLOUT || cfg->estimate_kind > MCTSB200_EST_TABLE) return fail(MCTSB200_ERR_INVALID_ARG, "unknown estimate_kind");
    if (cfg->estimate_kind == MCTSB200_EST_TABLE && (!dense || !cfg->value_table))
        return fail(MCTSB200_ERR_UNSUPPORTED, "table-valued estimate_value needs a dense-state MDP and a value_table");
    if ((cfg->init_q_table || cfg->init_n_table) && !dense) return fail(MCTSB200_ERR_UNSUPPORTED, "init tables need a dense-state MDP");
    if (cfg->estimate_kind == MCTSB200_EST_ROLLOUT) {
        if (!policy_ok(cfg->rollout_policy)) return fail(MCTSB200_ERR_UNSUPPORTED, "rollout policy %d is not available for this MDP", cfg->rollout_policy);
        if (cfg->rollout_policy == MCTSB200_POLICY_CONSTANT && (cfg->rollout_policy_param[0] < 0 || cfg->rollout_policy_param[0] >= n_actions))
            return fail(MCTSB200_ERR_INVALID_ARG, "constant rollout action out of range");
        if (cfg->rollout_policy == MCTSB200_POLICY_TABLE && (!dense || !cfg->policy_table))
            return fail(MCTSB200_ERR_UNSUPPORTED, "a tabulated rollout policy needs a dense-state MDP and a policy_table");
        if (cfg->rollout_max_depth < -1) return fail(MCTSB200_ERR_INVALID_ARG, "rollout max_depth must be >= -1");
        if (std::isnan(cfg->rollout_eps)) return fail(MCTSB200_ERR_INVALID_ARG, "rollout eps is NaN");
    }
    if (cfg->kind == MCTSB200_KIND_DPW) {
        const double v[4] = {cfg->k_action, cfg->alpha_action, cfg->k_state, cfg->alpha_state};
        for (double x : v)
            if (!(x >= 0.0) || !std::isfinite(x)) return fail(MCTSB200_ERR_UNSUPPORTED, "DPW k/alpha must be finite and >= 0");
        if (cfg->enable_action_pw && !cfg->check_repeat_action && !box_actions)  // (action nodes of a continuous space have their own table)
            return fail(MCTSB200_ERR_UNSUPPORTED, "action widening with check_repeat_action=false (duplicate action nodes) is not supported on the device");
    }
    return MCTSB200_OK;
}

// Sizes and (re)allocates the node tables of a plan call. With reuse_tree / keep_tree (src/vanilla.jl:213-217,
// src/dpw.jl:31-37) and compatible tables from the previous call, the tables are kept as they are (*reused = true).
template <class M>
int32_t setup_storage(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, long long n_trees, bool direct, bool* reused) {
    constexpr int A = M::kActions;
    const bool dpw = cfg->kind == MCTSB200_KIND_DPW;
    const long long n_iter = cfg->n_iterations;
    const int dense = MdpAccess<M>::dense_count(ctx);
    const bool merge_states = !dpw || cfg->check_repeat_state;
    long long NS = n_iter + 1;
    if (M::kDense && merge_states) NS = std::min<long long>(NS, dense);
    // A kept tree grows call after call (src/dpw.jl:31-37, src/vanilla.jl:213-217). A dense state space with merged states bounds
    // it; any other tree gets room for `keep_tree_calls` calls (one new state node per iteration at most) and reports
    // MCTSB200_ERR_CAPACITY when that is used up.
    const bool bounded = M::kDense && merge_states;
    long long calls = 1;
    if (cfg->keep_tree) {
        if (bounded) {
            NS = dense;
        } else {
            if (cfg->keep_tree_calls < 0 || cfg->keep_tree_calls > 100000) return fail(MCTSB200_ERR_INVALID_ARG, "keep_tree_calls out of range");
            calls = cfg->keep_tree_calls ? cfg->keep_tree_calls : 8;
            NS = calls * (n_iter + 1);  // per call: one new state node per iteration at most, and a root the tree may not hold yet
            if (NS > 0x7fffffffLL / 4) return fail(MCTSB200_ERR_CAPACITY, "kept trees would exceed 2^29 state nodes per tree");
        }
    }
    if (direct) NS = dense;
    const bool maintain_lookup = !dpw || cfg->keep_tree || cfg->check_repeat_state;
    if (M::kDense && maintain_lookup && NS > 65535) return fail(MCTSB200_ERR_CAPACITY, "dense state map limited to 65535 nodes per tree");
    long long NT = 0;
    if (dpw) {
        const long long by_iter = std::max<long long>(1, calls * n_iter * std::max(cfg->depth, 1));
        const long long by_pairs = NS * A * std::min<long long>(MdpAccess<M>::max_branch(ctx), NS);
        NT = cfg->check_repeat_state ? ((cfg->keep_tree && bounded) ? by_pairs : std::min(by_iter, by_pairs)) : std::max<long long>(1, calls * n_iter);
    }
    long long MAP = 0;
    size_t map_entry = 0;
    if (direct) {
        MAP = NS;  // insertion sequence number per state (uct_fast.cuh)
        map_entry = sizeof(uint16_t);
    } else if (M::kDense) {
        MAP = (dense + 63) / 64 * 64;
        map_entry = sizeof(uint16_t);
    } else {
        MAP = next_pow2(2 * NS);
        map_entry = sizeof(HashSlot);
    }
    // Unbounded branching (a plugin registered with max_branch = 0): every action node keeps the reference's transitions vector,
    // one entry per push, in chunks that double. A node with p pushes has allocated < 4p + 8 entries in all.
    long long NP = 0;
    if (dpw && MdpAccess<M>::max_branch(ctx) >= (1LL << 30)) {
        const long long pushes = std::max<long long>(1, calls * n_iter * std::max(cfg->depth, 1));  // at most one push per tree level
        NP = 4 * pushes + 8 * std::min<long long>(NS * A, pushes);
        if (NP > 0x7fffffffLL) return fail(MCTSB200_ERR_CAPACITY, "push log would exceed 2^31 entries per tree");
    }
    // continuous action spaces: one action node and one child-list entry per tree level at most, beside the transition pushes
    long long NA = 0;
    if constexpr (M::kActionDoubles > 0) {
        const long long pushes = std::max<long long>(1, calls * n_iter * std::max(cfg->depth, 1));
        NA = pushes + 1;
        NP = 4 * (2 * pushes) + 8 * std::min<long long>(NS + NA, 2 * pushes);
        if (NA > 0x7fffffffLL || NP > 0x7fffffffLL) return fail(MCTSB200_ERR_CAPACITY, "action-node table would exceed 2^31 entries per tree");
    }
    Geometry g;
    g.NS = (int)NS;
    g.NT = (int)std::max<long long>(NT, 1);
    g.MAP = (int)MAP;
    g.depth = std::max(cfg->depth, 1);
    g.NP = (int)NP;
    g.NA = (int)NA;

    const size_t b_hdr = sizeof(TreeHeader) * (size_t)n_trees;
    const size_t b_rows = sizeof(Row<M>) * (size_t)NS * (size_t)n_trees;
    // (vanilla UCT on a hashed-state single-successor model keeps its successor cache in the same table: search.cuh, kSuccCache)
    const bool uct_succ_cache = !dpw && !M::kDense && MdpAccess<M>::max_branch(ctx) == 1;
    const size_t b_aux = (dpw || uct_succ_cache) ? sizeof(Aux<M>) * (size_t)NS * (size_t)n_trees : 0;
    const size_t b_trans = dpw ? sizeof(TransRec) * (size_t)g.NT * (size_t)n_trees : 0;
    const size_t b_map = map_entry * (size_t)MAP * (size_t)n_trees;
    const size_t b_path = direct ? 0 : sizeof(PathEntry) * (size_t)g.depth * (size_t)n_trees;
    const size_t b_plog = sizeof(int) * (size_t)NP * (size_t)n_trees;
    const size_t b_anodes = sizeof(BoxANode) * (size_t)NA * (size_t)n_trees;
    const size_t total = b_hdr + b_rows + b_aux + b_trans + b_map + b_path + b_plog + b_anodes;

    Storage& st = ctx->st;
    *reused = false;
    if (cfg->keep_tree && st.valid && st.keep && st.kind == cfg->kind && st.mdp_id == MdpAccess<M>::mdp_id(ctx) && st.n_trees == n_trees) {
        if (st.direct != direct || st.geo.NS != g.NS || st.geo.NT != g.NT || st.geo.MAP != g.MAP || st.geo.NP != g.NP || st.geo.NA != g.NA)
            return fail(MCTSB200_ERR_STATE, "the kept trees cannot be continued with this configuration (different kernel layout); call clear_tree first");
        CUDA_TRY(st.path.ensure(b_path));
        st.geo.depth = g.depth;
        *reused = true;
        return MCTSB200_OK;
    }
    const size_t have = st.hdr.cap + st.rows.cap + st.aux.cap + st.trans.cap + st.map.cap + st.path.cap + st.plog.cap + st.anodes.cap;
    const bool grows = b_hdr > st.hdr.cap || b_rows > st.rows.cap || b_aux > st.aux.cap || b_trans > st.trans.cap || b_map > st.map.cap ||
                       b_path > st.path.cap || b_plog > st.plog.cap || b_anodes > st.anodes.cap;
    if (grows) {  // (cudaMemGetInfo is slow: only ask when something has to be allocated)
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        if (total > free_b + have)
            return fail(MCTSB200_ERR_CAPACITY, "node tables need %.2f GB but only %.2f GB of HBM are free; plan fewer roots per call", total / 1e9,
                        (free_b + have) / 1e9);
    }
    st.valid = false;
    CUDA_TRY(st.hdr.ensure(b_hdr));
    CUDA_TRY(st.rows.ensure(b_rows));
    CUDA_TRY(st.aux.ensure(b_aux));
    CUDA_TRY(st.trans.ensure(b_trans));
    CUDA_TRY(st.map.ensure(b_map));
    CUDA_TRY(st.path.ensure(b_path));
    CUDA_TRY(st.plog.ensure(b_plog));
    CUDA_TRY(st.anodes.ensure(b_anodes));
    st.kind = cfg->kind;
    st.direct = direct;
    st.keep = cfg->keep_tree != 0;
    st.visits = 0;
    st.mdp_id = MdpAccess<M>::mdp_id(ctx);
    st.n_trees = n_trees;
    st.geo = g;
    st.row_bytes = sizeof(Row<M>);
    st.aux_bytes = sizeof(Aux<M>);
    st.map_entry_bytes = map_entry;
    CUDA_TRY(cudaMemsetAsync(st.hdr.p, 0, b_hdr, ctx->stream));
    if (b_map) CUDA_TRY(cudaMemsetAsync(st.map.p, 0, b_map, ctx->stream));
    // (the fast kernel's own table is allocated and zero-filled in run_search, once the slot count is known)
    return MCTSB200_OK;
}

constexpr long long kUcbTabMax = 1LL << 22;  // entries of the log / sqrt-log / 1/sqrt tables (visit counts beyond: computed)

template <class M>
int32_t build_dev_cfg(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, DevCfg& d, long long prior_visits) {
    constexpr int A = M::kActions;
    std::memset(&d, 0, sizeof d);
    d.n_iterations = cfg->n_iterations;
    d.c = cfg->exploration_constant;
    d.init_Q = cfg->init_Q;
    d.estimate_const = cfg->estimate_const;
    d.rollout_eps = cfg->rollout_eps;
    {
        const double g = MdpAccess<M>::params(ctx).discount;  // (M::discount is a device function)
        const int steps = std::max(cfg->rollout_max_depth, cfg->depth) + 1;
        d.disc_always_positive = (cfg->rollout_eps == 0.0 && g > 0.0 && g <= 1.0 && std::pow(g, (double)steps) > 1e-300) ? 1 : 0;
        // |Q| <= max(|init_Q|, max|r| * sum_{t < H} g^t): every Q is init_Q or a running mean of discounted returns over at most
        // H = depth + rollout steps rewards (a little slack for the rounding of the running mean)
        const double H = (double)std::max(cfg->depth, 1) + (double)std::max(cfg->rollout_max_depth == -1 ? cfg->depth : cfg->rollout_max_depth, 0);
        const double geo = (g >= 0.0 && g < 1.0) ? std::min(H, 1.0 / (1.0 - g)) : (g == 1.0 ? H : NAN);
        d.q_bound = std::fmax(std::fabs(cfg->init_Q), MdpAccess<M>::reward_bound(ctx) * geo) * 1.0009765625;
        if (!std::isfinite(cfg->init_Q) || !std::isfinite(MdpAccess<M>::reward_bound(ctx) * geo)) d.q_bound = NAN;
        d.q_bound_tile = (cfg->estimate_kind == MCTSB200_EST_ROLLOUT && !cfg->init_q_table) ? d.q_bound : NAN;
    }
    d.seed = cfg->seed;
    d.depth = cfg->depth;
    d.K = cfg->rollouts_per_leaf;
    d.init_N = (int)cfg->init_N;
    d.estimate_kind = cfg->estimate_kind;
    d.policy = cfg->rollout_policy;
    for (int i = 0; i < 4; ++i) d.pparam[i] = cfg->rollout_policy_param[i];
    d.rollout_max_depth = cfg->rollout_max_depth;
    d.enable_action_pw = cfg->enable_action_pw;
    d.enable_state_pw = cfg->enable_state_pw;
    d.check_repeat_state = cfg->check_repeat_state;
    d.check_repeat_action = cfg->check_repeat_action;
    d.maintain_lookup = (cfg->keep_tree || cfg->check_repeat_state) ? 1 : 0;

    long long max_init_n = cfg->init_N;
    const int dense = MdpAccess<M>::dense_count(ctx);
    if constexpr (M::kFast) {
        std::vector<unsigned char> pol;
        if (cfg->estimate_kind == MCTSB200_EST_ROLLOUT && cfg->rollout_policy == MCTSB200_POLICY_TABLE)
            for (int i = 0; i < MdpAccess<M>::dense_count(ctx); ++i)
                if (cfg->policy_table[i] < 0 || cfg->policy_table[i] >= A) return fail(MCTSB200_ERR_INVALID_ARG, "policy_table entry %d out of range", i);
        if (cfg->estimate_kind == MCTSB200_EST_ROLLOUT && MdpAccess<M>::tabulate_policy(ctx, cfg, pol)) {
            CUDA_TRY(ctx->policy_tab.ensure(pol.size()));
            CUDA_TRY(cudaMemcpy(ctx->policy_tab.p, pol.data(), pol.size(), cudaMemcpyHostToDevice));
            d.policy_table = (const unsigned char*)ctx->policy_tab.p;
        }
    }
    if (cfg->next_action_table && cfg->kind == MCTSB200_KIND_DPW && cfg->enable_action_pw) {
        if (!M::kDense || dense <= 0) return fail(MCTSB200_ERR_UNSUPPORTED, "a tabulated next_action needs a dense-state MDP");
        std::vector<unsigned char> tab((size_t)dense);
        for (int i = 0; i < dense; ++i) {
            if (cfg->next_action_table[i] < 0 || cfg->next_action_table[i] >= A) return fail(MCTSB200_ERR_INVALID_ARG, "next_action_table entry %d out of range", i);
            tab[(size_t)i] = (unsigned char)cfg->next_action_table[i];
        }
        CUDA_TRY(ctx->next_action_tab.ensure(tab.size()));
        CUDA_TRY(cudaMemcpy(ctx->next_action_tab.p, tab.data(), tab.size(), cudaMemcpyHostToDevice));
        d.next_action_table = (const unsigned char*)ctx->next_action_tab.p;
    }
    if (cfg->init_q_table) {
        const size_t bytes = sizeof(double) * (size_t)dense * A;
        CUDA_TRY(ctx->init_q.ensure(bytes));
        CUDA_TRY(cudaMemcpyAsync(ctx->init_q.p, cfg->init_q_table, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d.init_q_table = (const double*)ctx->init_q.p;
    }
    if (cfg->init_n_table) {
        std::vector<int> tmp((size_t)dense * A);
        max_init_n = 0;
        for (size_t i = 0; i < tmp.size(); ++i) {
            const int64_t v = cfg->init_n_table[i];
            if (v < 0 || v > 1000000000LL) return fail(MCTSB200_ERR_INVALID_ARG, "init_n_table entry out of range");
            tmp[i] = (int)v;
            max_init_n = std::max<long long>(max_init_n, v);
        }
        CUDA_TRY(ctx->init_n.ensure(tmp.size() * sizeof(int)));
        CUDA_TRY(cudaMemcpy(ctx->init_n.p, tmp.data(), tmp.size() * sizeof(int), cudaMemcpyHostToDevice));
        d.init_n_table = (const int*)ctx->init_n.p;
    }
    if (cfg->value_table && cfg->estimate_kind == MCTSB200_EST_TABLE) {
        const size_t bytes = sizeof(double) * (size_t)dense;
        CUDA_TRY(ctx->value_tab.ensure(bytes));
        CUDA_TRY(cudaMemcpyAsync(ctx->value_tab.p, cfg->value_table, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d.value_table = (const double*)ctx->value_tab.p;
    }

    // visit-count bound: every backup adds one visit; at most depth backups per simulation
    const long long bound_raw = max_init_n * A + prior_visits + cfg->n_iterations * (long long)std::max(cfg->depth, 1) + 1;
    const long long bound = std::min<long long>(2000000000LL, bound_raw);
    if (bound_raw > 2000000000LL) return fail(MCTSB200_ERR_CAPACITY, "visit counts may exceed 32 bits (clear_tree to start over)");

    // log table: det_log(N), N < log_tab_n (grown lazily, shared by all configurations)
    // (a max_time search has no useful bound: a smaller table, visit counts beyond it take the exact path)
    int want = (int)std::min<long long>(bound + 1, std::isfinite(cfg->max_time) ? (1LL << 20) : kUcbTabMax);
    if (want > ctx->log_tab_n) {
        want = (int)std::min<long long>(std::max<long long>(want, 2LL * ctx->log_tab_n), kUcbTabMax);  // kept trees: grow in doublings
        std::vector<double> tab((size_t)want), stab((size_t)want), rtab((size_t)want);
        for (int i = 0; i < want; ++i) {
            tab[(size_t)i] = det_log((double)i);
            stab[(size_t)i] = i >= 1 ? std::sqrt(tab[(size_t)i]) : 0.0;
            rtab[(size_t)i] = i >= 1 ? 1.0 / std::sqrt((double)i) : INFINITY;
        }
        CUDA_TRY(ctx->log_tab.ensure(sizeof(double) * (size_t)want));
        CUDA_TRY(ctx->sqrtlog_tab.ensure(sizeof(double) * (size_t)want));
        CUDA_TRY(ctx->rsqrt_tab.ensure(sizeof(double) * (size_t)want));
        CUDA_TRY(cudaMemcpy(ctx->log_tab.p, tab.data(), sizeof(double) * (size_t)want, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(ctx->sqrtlog_tab.p, stab.data(), sizeof(double) * (size_t)want, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(ctx->rsqrt_tab.p, rtab.data(), sizeof(double) * (size_t)want, cudaMemcpyHostToDevice));
        ctx->log_tab_n = want;
    }
    d.log_tab = (const double*)ctx->log_tab.p;
    d.sqrtlog_tab = (const double*)ctx->sqrtlog_tab.p;
    d.rsqrt_tab = (const double*)ctx->rsqrt_tab.p;
    d.log_tab_n = ctx->log_tab_n;
    if (const char* ev = std::getenv("MCTSB200_UCB_TABLE_ENTRIES"))  // test hook: a short table exercises the beyond-the-table path
        if (std::atoi(ev) >= 2048) d.log_tab_n = std::min(d.log_tab_n, std::atoi(ev));
    d.rsqrt_tab_n = d.log_tab_n;

    if (cfg->kind == MCTSB200_KIND_DPW) {
        for (int len = 0; len < 16; ++len) d.nmin_action[len] = len <= A ? nmin_for(cfg->k_action, cfg->alpha_action, len, bound) : 0x7fffffff;
        if constexpr (M::kActionDoubles > 0) {
            // a continuous action space: a state node may own as many children as the widening rule allows (one more per visit at most)
            std::vector<int> na;
            // (a kept tree brings the children of earlier calls -- at most one per earlier visit; the table ends at the first length
            // no visit count below the bound allows)
            const long long max_len = std::min<long long>(prior_visits + cfg->n_iterations * (long long)std::max(cfg->depth, 1) + 1, (1 << 20) - 2);
            for (long long len = 0; len <= max_len; ++len) {
                const int v = nmin_for(cfg->k_action, cfg->alpha_action, (int)len, bound);
                na.push_back(v);
                if (v == 0x7fffffff) break;
            }
            CUDA_TRY(ctx->nmin_action_tab.ensure(na.size() * sizeof(int)));
            CUDA_TRY(cudaMemcpy(ctx->nmin_action_tab.p, na.data(), na.size() * sizeof(int), cudaMemcpyHostToDevice));
            d.nmin_action_tab = (const int*)ctx->nmin_action_tab.p;
            d.nmin_action_n = (int)na.size();
        }
        std::vector<int> ns;
        // n_a_children never exceeds the model's branching (merged states) nor the pushes to its action node: one per visit, and an
        // unmerged tree visits a node at most once per iteration. A kept tree brings the pushes of earlier calls (prior_visits: found
        // by the randomised test -- a table sized for one call made kept unmerged trees stop widening). The loop ends at the first
        // length no visit count below the bound allows.
        const long long need_len = cfg->check_repeat_state
                                       ? std::min<long long>(MdpAccess<M>::max_branch(ctx), prior_visits + cfg->n_iterations * (long long)std::max(cfg->depth, 1) + 1)
                                       : prior_visits + cfg->n_iterations + 1;
        const long long max_len = std::min<long long>(need_len, (1 << 20) - 2);
        for (long long len = 0; len <= max_len; ++len) {
            const int v = nmin_for(cfg->k_state, cfg->alpha_state, (int)len, bound);
            ns.push_back(v);
            if (v == 0x7fffffff) break;
        }
        if (ns.back() != 0x7fffffff && (long long)ns.size() - 1 < need_len)
            return fail(MCTSB200_ERR_UNSUPPORTED, "state-widening table would exceed 2^20 entries");
        CUDA_TRY(ctx->nmin_state.ensure(ns.size() * sizeof(int)));
        CUDA_TRY(cudaMemcpy(ctx->nmin_state.p, ns.data(), ns.size() * sizeof(int), cudaMemcpyHostToDevice));
        d.nmin_state = (const int*)ctx->nmin_state.p;
        d.nmin_state_n = (int)ns.size();
    }
    return MCTSB200_OK;
}

// debugging / tuning switches (environment): MCTSB200_PATH_SMEM=0 keeps the tile kernel's path stacks in HBM,
// MCTSB200_FAST=0 disables the shared-memory lane-per-tree kernels, MCTSB200_LANE_ORDER=0/1 switches the grouping of trees
// with equal root states into the same warps (default: on exactly when equal roots give identical trees, see lockstep())
bool env_enabled(const char* name, bool dflt = true) {
    const char* v = std::getenv(name);
    if (!v || !v[0]) return dflt;
    return v[0] != '0';
}

bool use_path_smem(int depth, int G) {
#ifdef MCTSB200_HOST_EMU
    return false;
#endif
    return env_enabled("MCTSB200_PATH_SMEM") && path_smem_bytes(depth, G) <= kPathSmemLimit;
}

// Tile-kernel occupancy and launch: the ahead-of-time instantiations (inst_*.cu) or, for a plugin registered at run time,
// the kernels NVRTC compiled around it (plugin_rt.cu).
#ifndef MCTSB200_HOST_EMU
const void* plugin_kernel_or_fail(mctsb200_ctx* ctx, int kind, int kernel_id) {
    std::string err;
    const void* k = mctsb200_rt::plugin_kernel(ctx->cur_plugin, kind, kernel_id, err);
    if (!k) fail(MCTSB200_ERR_CUDA, "%s", err.c_str());
    return k;
}
int tile_kernel_id(int G) { return G == 1 ? 0 : G == 2 ? 1 : G == 4 ? 2 : G == 8 ? 3 : G == 16 ? 4 : 5; }
#endif

template <class M, int KIND>
int tile_occupancy(mctsb200_ctx* ctx, int G, size_t smem) {
#ifndef MCTSB200_HOST_EMU
    if constexpr (M::kCustom) {
        const void* k = plugin_kernel_or_fail(ctx, KIND, tile_kernel_id(G));
        if (!k) return 0;
        int blocks = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k, kBlockThreads, smem) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        return blocks;
    } else
#endif
    {
        (void)ctx;
        return search_occupancy<M, KIND>(G, smem);
    }
}

template <class M, int KIND>
int32_t tile_launch(mctsb200_ctx* ctx, const KernelArgs<M>& args, int G) {
#ifndef MCTSB200_HOST_EMU
    if constexpr (M::kCustom) {
        const void* k = plugin_kernel_or_fail(ctx, KIND, tile_kernel_id(G));
        if (!k) return MCTSB200_ERR_CUDA;
        const long long threads = args.n_slots * G;
        const unsigned grid = (unsigned)((threads + kBlockThreads - 1) / kBlockThreads);
        const size_t smem = args.path_smem ? path_smem_bytes(args.geo.depth, G, args.path_tiles) : 0;
        void* params[] = {(void*)&args};
        CUDA_TRY(cudaLaunchKernel(k, dim3(grid), dim3(kBlockThreads), params, smem, ctx->stream));
        return MCTSB200_OK;
    } else
#endif
    {
        launch_search<M, KIND>(args, G, ctx->stream);
        return MCTSB200_OK;
    }
}

template <class M, int KIND>
int pick_lanes(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, long long n_trees) {
#ifdef MCTSB200_HOST_EMU
    return 1;  // the host emulator runs one lane at a time
#endif
    if (cfg->lanes_per_tree) return cfg->lanes_per_tree;
    // One rollout per leaf and room for a warp per tree: a single lane runs the tree and the other 31 stay idle
    // (run_search spreads the trees, see order_spread_kernel) -- the same one-tree-per-warp isolation as a 32-lane tile
    // without its shuffles and redundant lanes (MountainCar DPW, 1 024 roots: 13.5 vs 10.5 Msim/s).
    if (cfg->rollouts_per_leaf == 1) {
        const size_t smem1 = use_path_smem(std::max(cfg->depth, 1), 1) ? path_smem_bytes(std::max(cfg->depth, 1), 1) : 0;
        if (n_trees * 32 <= (long long)tile_occupancy<M, KIND>(ctx, 1, smem1) * kBlockThreads * ctx->sm_count) return 1;
    }
    // Widest tile that still lets every tree be resident at once (one wave). Lanes beyond |A| and K do no useful work,
    // but a wider tile also means fewer TREES per warp, and trees sharing a warp wait for each other in every phase
    // (measured on MountainCar DPW, 1 024 roots: 6.3 / 8.4 / 10.3 / 12.8 Msim/s for 4 / 8 / 16 / 32 lanes per tree).
    int best = 1;
    const int depth = std::max(cfg->depth, 1);
    for (int G = 2; G <= 32; G *= 2) {
        const size_t smem = use_path_smem(depth, G) ? path_smem_bytes(depth, G) : 0;
        const long long resident = (long long)tile_occupancy<M, KIND>(ctx, G, smem) * kBlockThreads * ctx->sm_count;
        if (n_trees * G <= resident) best = G;
    }
    return best;
}

void fill_stats(mctsb200_plan_stats* s, const mctsb200_solver_cfg* cfg, const unsigned long long* c, int n_actions) {
    s->n_simulations = (int64_t)c[CNT_SIMULATIONS];
    s->n_tree_levels = (int64_t)c[CNT_LEVELS];
    s->n_dpw_children_seen = (int64_t)c[CNT_CHILDREN_SEEN];
    s->n_state_inserts = (int64_t)c[CNT_STATE_INSERTS];
    s->n_action_inserts = (int64_t)c[CNT_ACTION_INSERTS];
    s->n_rollouts = (int64_t)c[CNT_ROLLOUTS];
    s->n_rollout_steps = (int64_t)c[CNT_ROLLOUT_STEPS];
    s->n_transition_samples = (int64_t)c[CNT_TRANS_SAMPLES];
    s->n_select_exact = (int64_t)c[CNT_SELECT_EXACT];
    if (cfg->kind == MCTSB200_KIND_UCT)
        s->algorithmic_bytes = (80 + 16 * (int64_t)n_actions) * s->n_tree_levels + (48 + 16 * (int64_t)n_actions) * s->n_state_inserts;
    else
        s->algorithmic_bytes = 144 * s->n_tree_levels + 24 * s->n_dpw_children_seen + 56 * s->n_state_inserts + 72 * s->n_action_inserts;
}

constexpr size_t kFastSmemLimit = 220 * 1024;

// The shared-memory lane-per-tree kernel (uct_fast.cuh) applies to vanilla UCT with one rollout per leaf on a
// dense-state MDP whose tables fit in shared memory. Returns the policy variant or -1.
template <class M, int KIND>
int fast_variant(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, long long prior_visits) {
    if constexpr (!M::kFast) {
        return -1;
    } else {
        if (!env_enabled("MCTSB200_FAST") || !MdpAccess<M>::fast_ready(ctx)) return -1;
        if (cfg->enable_tree_vis) return -1;  // (the tile kernel records _vis_stats)
        if (cfg->estimate_kind != MCTSB200_EST_ROLLOUT || cfg->rollouts_per_leaf != 1 || cfg->lanes_per_tree > 1) return -1;
        if (cfg->init_q_table || cfg->init_n_table) return -1;
        if (cfg->exploration_constant == 0.0) return -1;  // argmax-Q selection: the tile kernel has that path
        // (iteration, level) tags of the children: 24 bits of iteration for UCT, 16 for DPW
        if (cfg->depth > 254 || cfg->n_iterations >= (KIND == MCTSB200_KIND_UCT ? (1LL << 24) - 1 : (1LL << 16) - 1)) return -1;
        if (KIND == MCTSB200_KIND_DPW && !(cfg->check_repeat_state && cfg->check_repeat_action)) return -1;  // unmerged nodes: tile kernel
        if (KIND == MCTSB200_KIND_DPW && cfg->next_action_table) return -1;  // a tabulated next_action: tile kernel
        (void)prior_visits;  // (visit counts beyond the UCB tables take the exact path inside the kernels)
        if (fast_smem_bytes<M>(MdpAccess<M>::params(ctx), std::max(cfg->depth, 1), 64) > kFastSmemLimit) return -1;
        switch (cfg->rollout_policy) {
            case MCTSB200_POLICY_RANDOM: return FAST_POLICY_RANDOM;
            case MCTSB200_POLICY_CONSTANT: return FAST_POLICY_CONSTANT;
            case MCTSB200_POLICY_GW_SEEK: return FAST_POLICY_TABLE;
            case MCTSB200_POLICY_TABLE: return FAST_POLICY_TABLE;
            default: return -1;
        }
    }
}

// threads per block of the fast kernel: one block per SM, as many lanes as the batch needs and the per-level records
// in shared memory allow
template <class M>
int fast_block_threads(mctsb200_ctx* ctx, long long slots_per_sm, int depth) {
    int block = (int)std::min<long long>(kFastMaxThreads, std::max<long long>(32, (slots_per_sm + 31) / 32 * 32));
    if constexpr (M::kFast)
        while (block > 32 && fast_smem_bytes<M>(MdpAccess<M>::params(ctx), depth, block) > kFastSmemLimit) block -= 32;
    return block;
}

// Converts the fast kernels' interleaved tables of the last plan call into the Row / Aux / TransRec tables.
template <class M, int KIND>
int32_t finalize_pending(mctsb200_ctx* ctx) {
    Storage& st = ctx->st;
    st.finalize_fn = nullptr;
    if constexpr (M::kFast) {
        const long long n_slots = st.fin_slots;
        const int* order = st.fin_order ? (const int*)ctx->order.p : nullptr;
        unsigned long long* max_visits = (unsigned long long*)ctx->counters.p + CNT_MAX_VISITS;
        if (KIND == MCTSB200_KIND_UCT) {
            const long long n_rows = (n_slots + 31) / 32 * 32 * st.geo.NS;
            auto kernel = finalize_rows_kernel<M>;
            MCTSB200_LAUNCH(kernel, (unsigned)((n_rows + 255) / 256), 256, 0, ctx->stream, (const FastRow*)st.rows_fast.p, order, (const uint16_t*)st.map.p,
                            (Row<M>*)st.rows.p, st.geo.NS, n_slots, max_visits);
        } else {
            auto kernel = dpw_finalize_kernel<M>;
            MCTSB200_LAUNCH(kernel, (unsigned)((n_slots + 127) / 128), 128, 0, ctx->stream, (const FastRow*)st.rows_fast.p,
                            (const DpwTransRow*)st.trans_fast.p, order, (const uint16_t*)st.map.p, MdpAccess<M>::params(ctx), (TreeHeader*)st.hdr.p,
                            (Row<M>*)st.rows.p, (Aux<M>*)st.aux.p, (TransRec*)st.trans.p, st.geo.NS, st.geo.NT, n_slots, max_visits);
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // (the tree queries read with plain cudaMemcpy)
    }
    return MCTSB200_OK;
}

// Runs the search for trees whose roots are already on the device. d_* outputs may be null.
template <class M, int KIND>
int32_t run_search(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, const typename M::AbiState* d_roots, long long n_trees,
                   long long root_stride, long long tree_index_base, int* d_action, double* d_bestq, int* d_status,
                   mctsb200_plan_stats* stats, unsigned* status_flags = nullptr) {
    Storage& st = ctx->st;
    // visits already in the trees if this call continues kept ones (reuse_tree / keep_tree)
    const bool may_reuse = cfg->keep_tree && st.valid && st.keep && st.kind == cfg->kind && st.mdp_id == MdpAccess<M>::mdp_id(ctx) && st.n_trees == n_trees;
    const long long prior_visits = may_reuse ? st.visits : 0;
    DevCfg dcfg;
    int32_t rc = build_dev_cfg<M>(ctx, cfg, dcfg, prior_visits);
    if (rc) return rc;
    const int fast = fast_variant<M, KIND>(ctx, cfg, prior_visits);
    if (may_reuse) {
        dcfg.q_bound = (std::isnan(st.q_bound) || std::isnan(dcfg.q_bound)) ? NAN : std::max(dcfg.q_bound, st.q_bound);
        dcfg.q_bound_tile = (std::isnan(st.q_bound_tile) || std::isnan(dcfg.q_bound_tile)) ? NAN : std::max(dcfg.q_bound_tile, st.q_bound_tile);
    }
    bool reused = false;
    rc = setup_storage<M>(ctx, cfg, n_trees, fast >= 0, &reused);
    if (rc) return rc;
    st.vis_branch = 0;
    if (KIND == MCTSB200_KIND_UCT && cfg->enable_tree_vis) {
        const long long mb = MdpAccess<M>::max_branch(ctx);
        if (mb < 1 || mb > 64) return fail(MCTSB200_ERR_UNSUPPORTED, "enable_tree_vis needs a model with bounded branching (max_branch 1..64)");
        st.vis_branch = (int)mb;
        const size_t bytes = sizeof(VisRec) * (size_t)st.geo.NS * M::kActions * (size_t)mb * (size_t)n_trees;
        const bool fresh = !reused || bytes > st.vis.cap;
        CUDA_TRY(st.vis.ensure(bytes));
        if (fresh) CUDA_TRY(cudaMemsetAsync(st.vis.p, 0, bytes, ctx->stream));
    }
    if (reused) {  // same trees, new search: iteration counters and per-root status start over
        MCTSB200_LAUNCH(begin_call_kernel, (unsigned)((n_trees + 255) / 256), 256, 0, ctx->stream, (TreeHeader*)st.hdr.p, n_trees);
        CUDA_TRY(cudaGetLastError());
    }
    st.valid = false;
    CUDA_TRY(ctx->counters.ensure(sizeof(unsigned long long) * CNT_COUNT));
    CUDA_TRY(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * CNT_COUNT, ctx->stream));

    KernelArgs<M> args;
    std::memset(&args, 0, sizeof args);
    args.mdp = MdpAccess<M>::params(ctx);
    args.cfg = dcfg;
    args.geo = st.geo;
    args.hdr = (TreeHeader*)st.hdr.p;
    args.rows = (Row<M>*)st.rows.p;
    args.aux = (KIND == MCTSB200_KIND_DPW || (!M::kDense && MdpAccess<M>::max_branch(ctx) == 1 && env_enabled("MCTSB200_SUCC_CACHE"))) ? (Aux<M>*)st.aux.p : nullptr;
    args.trans = KIND == MCTSB200_KIND_DPW ? (TransRec*)st.trans.p : nullptr;
    args.plog = (KIND == MCTSB200_KIND_DPW && st.geo.NP > 0) ? (int*)st.plog.p : nullptr;
    st.action_doubles = M::kActionDoubles;
    if constexpr (M::kActionDoubles > 0) {
        args.anodes = (BoxANode*)st.anodes.p;
        CUDA_TRY(st.action_vec.ensure(sizeof(double) * M::kActionDoubles * (size_t)n_trees));
        args.action_vec_out = (double*)st.action_vec.p;
    }
    args.map = st.map.p;
    args.path = (PathEntry*)st.path.p;
    args.roots = d_roots;
    args.counters = (unsigned long long*)ctx->counters.p;
    args.n_trees = n_trees;
    args.tree_index_base = tree_index_base;
    args.root_stride = root_stride;
    args.vis = st.vis_branch ? (VisRec*)st.vis.p : nullptr;
    args.vis_branch = st.vis_branch;
    // (measured on B200, profiles/r2b_*: 4 % SLOWER on config 2, 1 % on config 3 -- the rows that miss L2 are not what the
    // levels wait for; kept as a switch for the record)
    args.prefetch_l2 = env_enabled("MCTSB200_PREFETCH_L2", false) ? std::max(1, std::atoi(std::getenv("MCTSB200_PREFETCH_L2"))) : 0;

    const int G = fast >= 0 ? 1 : pick_lanes<M, KIND>(ctx, cfg, n_trees);
    const int n_iter = (int)cfg->n_iterations;
    int launches = 0, done = 0;
    args.path_smem = fast < 0 && use_path_smem(st.geo.depth, G) ? 1 : 0;
    args.n_slots = n_trees;
    int fast_block = 0;  // block size the lane order was dealt for
    if constexpr (M::kDense) {
        // (kept trees live in their lane's slot of the interleaved table: the slot <-> tree map must not change between calls)
        if (root_stride == 1 && n_trees > 32 && n_trees < (1LL << 30) && !cfg->keep_tree &&
            env_enabled("MCTSB200_LANE_ORDER", fast >= 0 && MdpAccess<M>::lockstep(ctx, cfg))) {
            const int n_bins = MdpAccess<M>::dense_count(ctx);
            const int tpw = 32 / G;  // tiles per warp
            // upper bound of the padded slot count (the exact one is only known on the device): every non-empty
            // group wastes at most tpw-1 slots
            const long long bound = n_trees + std::min<long long>(n_bins, n_trees) * (tpw - 1);
            long long n_warps = (bound + tpw - 1) / tpw;
            // thread-block shape of the launch that will consume the order
            long long wpb = kBlockThreads / 32;
            if (fast >= 0) {
                const long long per_sm = (n_warps * tpw + ctx->sm_count - 1) / ctx->sm_count;
                fast_block = fast_block_threads<M>(ctx, per_sm, st.geo.depth);
                wpb = fast_block / 32;
            }
            const long long n_blocks = (n_warps + wpb - 1) / wpb;
            n_warps = n_blocks * wpb;
            const unsigned grid = (unsigned)((n_trees + 255) / 256);
            CUDA_TRY(ctx->order.ensure(sizeof(int) * (size_t)(n_warps * tpw)));
            CUDA_TRY(ctx->order_bins.ensure(sizeof(int) * (size_t)n_bins));
            CUDA_TRY(cudaMemsetAsync(ctx->order_bins.p, 0, sizeof(int) * (size_t)n_bins, ctx->stream));
            CUDA_TRY(cudaMemsetAsync(ctx->order.p, 0xff, sizeof(int) * (size_t)(n_warps * tpw), ctx->stream));
            auto hist = order_hist_kernel<M>;
            auto scatter = order_scatter_kernel<M>;
            MCTSB200_LAUNCH(hist, grid, 256, 0, ctx->stream, args.mdp, d_roots, n_trees, (int*)ctx->order_bins.p);
            MCTSB200_LAUNCH(order_scan_kernel, 1, 1024, 0, ctx->stream, (int*)ctx->order_bins.p, n_bins, tpw);
            MCTSB200_LAUNCH(scatter, grid, 256, 0, ctx->stream, args.mdp, d_roots, n_trees, (int*)ctx->order_bins.p, (int*)ctx->order.p, tpw, n_blocks, wpb);
            CUDA_TRY(cudaGetLastError());
            args.order = (const int*)ctx->order.p;
            args.n_slots = n_warps * tpw;
            launches += 3;
        }
    }
    // (root_stride 0 = an ensemble of one root, mctsb200_plan_root_parallel: the lanes do not run in lockstep either)
    if ((fast >= 0 || G == 1) && !args.order && !cfg->keep_tree && n_trees < (1LL << 30)) {
        // lanes that do not run in lockstep: spread the trees over more warps while one wave of blocks still holds them all
        int spread = 1;
        const char* ev = std::getenv("MCTSB200_SPREAD");
        if (ev && ev[0]) {
            spread = std::atoi(ev);
        } else if (fast < 0) {
            // tile kernel with one lane per tree: a warp per tree when one wave holds them all (pick_lanes chose G = 1 for that)
            const size_t smem1 = args.path_smem ? path_smem_bytes(st.geo.depth, 1) : 0;
            if (cfg->rollouts_per_leaf == 1 && n_trees * 32 <= (long long)tile_occupancy<M, KIND>(ctx, 1, smem1) * kBlockThreads * ctx->sm_count) spread = 32;
        } else if (!MdpAccess<M>::lockstep(ctx, cfg)) {
            // (measured on B200: pays up to about 8 warps per SM, config 3 +2.5 %, 4096 DPW roots +24 %; beyond that the wider
            // footprint in L1/L2 eats the gain). A warp's iteration lasts as long as the longest descent plus the longest rollout
            // plus the longest backup among its lanes; with one or two trees per warp that sum shrinks to the tree's own -- for
            // vanilla UCT down to one tree per warp (1 184 trees: 14.5 ms against 18.4 ms at four per warp), not for DPW
            // (profiles/r2_experiments.md)
            const int max_spread = KIND == MCTSB200_KIND_UCT ? 32 : 8;
            for (int s2 = 2; s2 <= max_spread; s2 *= 2)
                if ((n_trees * s2 + 31) / 32 <= (long long)ctx->sm_count * 8) spread = s2;
        }
        if (spread == 2 || spread == 4 || spread == 8 || spread == 16 || spread == 32) {
            const int active = 32 / spread;
            const long long n_warps = (n_trees + active - 1) / active;
            args.n_slots = n_warps * 32;
            CUDA_TRY(ctx->order.ensure(sizeof(int) * (size_t)args.n_slots));
            MCTSB200_LAUNCH(order_spread_kernel, (unsigned)((args.n_slots + 255) / 256), 256, 0, ctx->stream, (int*)ctx->order.p, args.n_slots, active, n_trees);
            CUDA_TRY(cudaGetLastError());
            args.order = (const int*)ctx->order.p;
            launches += 1;
            if (fast < 0) {
                // only the first `active` tiles of a warp hold a tree: their path stacks are all the block needs in shared memory
                // (depth 50 with one tree per warp: 3 KB instead of 100 KB, which did not fit and left the stacks in HBM)
                args.path_tiles = active;
                args.path_smem = (env_enabled("MCTSB200_PATH_SMEM") && path_smem_bytes(st.geo.depth, G, active) <= kPathSmemLimit) ? 1 : 0;
#ifdef MCTSB200_HOST_EMU
                args.path_smem = 0;
#endif
            }
        }
    }
    if (fast >= 0) {
        // the fast kernel's working table: 64 B per (slot, state), slots padded to whole warps, zero = "not in the tree"
        const long long slots32 = (args.n_slots + 31) / 32 * 32;
        const size_t bytes = (size_t)slots32 * (size_t)st.geo.NS * 64;
        if (bytes > st.rows_fast.cap) {
            size_t free_b = 0, total_b = 0;
            CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            if (bytes > free_b + st.rows_fast.cap)
                return fail(MCTSB200_ERR_CAPACITY, "node tables need %.2f GB more HBM than is free; plan fewer roots per call", (bytes - st.rows_fast.cap) / 1e9);
        }
        if (!reused) {
            CUDA_TRY(st.rows_fast.ensure(bytes));
            CUDA_TRY(cudaMemsetAsync(st.rows_fast.p, 0, bytes, ctx->stream));
        }
        args.rows_fast = st.rows_fast.p;
        if (KIND == MCTSB200_KIND_DPW) {
            if (!reused) {
                CUDA_TRY(st.trans_fast.ensure(2 * bytes));
                CUDA_TRY(cudaMemsetAsync(st.trans_fast.p, 0, 2 * bytes, ctx->stream));
            }
            args.trans_fast = st.trans_fast.p;
        }
    }
    // one search launch over iterations [args.iter_begin, args.iter_end)
    auto launch = [&]() -> int32_t {
        if constexpr (M::kFast) {
            if (fast >= 0) {
                const long long per_sm = (args.n_slots + ctx->sm_count - 1) / ctx->sm_count;
                const int block = fast_block ? fast_block : fast_block_threads<M>(ctx, per_sm, args.geo.depth);
                const int grid = (int)((args.n_slots + block - 1) / block);
                const size_t smem = fast_smem_bytes<M>(args.mdp, args.geo.depth, block);
                if (KIND == MCTSB200_KIND_UCT) {
                    // successors-ahead descent: only where the extra lookups land in idle issue slots
                    const bool ahead = env_enabled("MCTSB200_AHEAD", MdpAccess<M>::lockstep(ctx, cfg) || args.n_slots <= 128LL * ctx->sm_count);
                    // single-outcome transitions: no random words for them at all
                    const bool det = ahead && MdpAccess<M>::fast_deterministic(ctx) && env_enabled("MCTSB200_DET", true);
                    return launch_uct_fast<M>(args, fast, det ? FAST_MODE_DET : ahead ? FAST_MODE_AHEAD : FAST_MODE_PLAIN, grid, block, smem, ctx->stream);
                }
                return launch_dpw_fast<M>(args, fast, grid, block, smem, ctx->stream);
            }
        }
        return tile_launch<M, KIND>(ctx, args, G);
    };
    float kernel_ms = 0.f;
    // seconds on the clock max_time is measured with: the caller's timer (solver.timer) or the monotonic clock
    auto now_s = [&]() -> double {
        if (ctx->timer_fn) return ctx->timer_fn(ctx->timer_user);
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    const bool timed = std::isfinite(cfg->max_time);
    const bool hooked = ctx->progress_fn != nullptr && ctx->progress_every >= 1;
    if (!timed && !hooked) {
        args.iter_begin = 0;
        args.iter_end = n_iter;
        CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
        if (launch()) return M::kCustom ? MCTSB200_ERR_CUDA : fail(MCTSB200_ERR_CUDA, "launching the search kernel failed");
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        launches += 1;
        done = n_iter;
    } else {
        // max_time (src/vanilla.jl:232, src/dpw.jl:52): the timer is read after each chunk of iterations; at least one
        // iteration always runs. Chunks grow until one takes ~1/16 of the budget. A progress callback
        // (src/vanilla.jl:231) fixes the chunk at `every` iterations.
        const double start_s = now_s();
        int chunk = hooked ? (int)std::min<long long>(ctx->progress_every, 1 << 20) : 1;
        CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
        while (done < n_iter || (n_iter == 0 && launches == 0)) {
            args.iter_begin = done;
            args.iter_end = std::min(n_iter, done + chunk);
            const double c0 = now_s();
            if (launch()) return M::kCustom ? MCTSB200_ERR_CUDA : fail(MCTSB200_ERR_CUDA, "launching the search kernel failed");
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            ++launches;
            done = args.iter_end;
            if (hooked) {
                // the tree queries work inside the callback (the fast kernels' tables are converted on demand)
                st.valid = true;
                if constexpr (M::kFast) {
                    st.fin_slots = args.n_slots;
                    st.fin_order = args.order != nullptr;
                    st.finalize_fn = fast >= 0 ? &finalize_pending<M, KIND> : nullptr;
                }
                const int32_t stop = ctx->progress_fn(ctx->progress_user, done, n_iter);
                st.valid = false;
                st.finalize_fn = nullptr;
                if (stop) break;
            }
            if (n_iter == 0) break;
            if (!timed) continue;
            const double now = now_s();
            const double elapsed = now - start_s;
            if (elapsed >= cfg->max_time) break;
            if (hooked) continue;
            const double chunk_s = now - c0;
            if (chunk_s < cfg->max_time / 16.0) chunk = std::min(chunk * 2, 1 << 20);
            const double per_iter = chunk_s / (double)(args.iter_end - args.iter_begin);
            const double left = cfg->max_time - elapsed;
            if (per_iter > 0 && chunk * per_iter > left) chunk = std::max(1, (int)(left / per_iter));
        }
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    }
    // kept trees need the largest visit count (a by-product of the conversion) for the next call's table bound
    const bool lazy = fast >= 0 && !cfg->keep_tree && env_enabled("MCTSB200_LAZY_EXPORT");
    st.finalize_fn = nullptr;
    if constexpr (M::kFast) if (fast >= 0) {
        st.fin_slots = args.n_slots;
        st.fin_order = args.order != nullptr;
        if (lazy) {
            auto kernel = decide_fast_kernel<M, KIND>;
            MCTSB200_LAUNCH(kernel, (unsigned)((args.n_slots + 127) / 128), 128, 0, ctx->stream, (const TreeHeader*)st.hdr.p, (const FastRow*)st.rows_fast.p,
                            args.order, st.geo.NS, args.n_slots, d_action, d_bestq, d_status, (unsigned long long*)ctx->counters.p + CNT_STATUS_FLAGS);
            CUDA_TRY(cudaGetLastError());
            ++launches;
            st.finalize_fn = &finalize_pending<M, KIND>;
        } else {
            rc = finalize_pending<M, KIND>(ctx);
            if (rc) return rc;
            ++launches;
        }
    }
    if constexpr (M::kActionDoubles > 0) {
        const unsigned grid = (unsigned)((n_trees + 127) / 128);
        auto kernel = decide_box_kernel<M>;
        MCTSB200_LAUNCH(kernel, grid, 128, 0, ctx->stream, (const TreeHeader*)st.hdr.p, (const Aux<M>*)st.aux.p, (const BoxANode*)st.anodes.p,
                        (const int*)st.plog.p, st.geo, n_trees, d_action, d_bestq, d_status, args.action_vec_out,
                        (unsigned long long*)ctx->counters.p + CNT_STATUS_FLAGS);
        CUDA_TRY(cudaGetLastError());
        ++launches;
    } else if (!lazy) {
        const unsigned grid = (unsigned)((n_trees + 127) / 128);
        auto kernel = decide_kernel<M>;
        MCTSB200_LAUNCH(kernel, grid, 128, 0, ctx->stream, (const TreeHeader*)st.hdr.p, (const Row<M>*)st.rows.p, st.geo.NS, n_trees, d_action, d_bestq,
                        d_status, (unsigned long long*)ctx->counters.p + CNT_STATUS_FLAGS);
        CUDA_TRY(cudaGetLastError());
        ++launches;
    }
    unsigned long long h_cnt[CNT_COUNT];
    CUDA_TRY(cudaMemcpyAsync(h_cnt, ctx->counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaEventElapsedTime(&kernel_ms, ctx->ev0, ctx->ev1));
    st.valid = true;
    // bound of the visit counts now in the trees: exact after a fast-kernel call, else one visit per backup
    st.q_bound = dcfg.q_bound;
    st.q_bound_tile = dcfg.q_bound_tile;
    st.visits = fast >= 0 ? (long long)h_cnt[CNT_MAX_VISITS] : prior_visits + (long long)done * std::max(cfg->depth, 1);
    if (status_flags) *status_flags = (unsigned)h_cnt[CNT_STATUS_FLAGS];
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->n_trees = n_trees;
        fill_stats(stats, cfg, h_cnt, M::kActions);
        stats->kernel_ms = kernel_ms;
        stats->lanes_per_tree = G;
        stats->n_kernel_launches = launches;
        stats->iterations_done = done;
        stats->kernel_variant = fast >= 0 ? 1 : 0;
    }
    return MCTSB200_OK;
}

// what a continuous action space allows: DPW with action widening (vanilla UCT and DPW without widening enumerate actions(mdp, s),
// src/vanilla.jl:297, src/dpw.jl:107), one rollout per leaf or a constant estimate, fresh trees
int32_t check_box_cfg(const mctsb200_solver_cfg* cfg) {
    if (cfg->kind != MCTSB200_KIND_DPW) return fail(MCTSB200_ERR_UNSUPPORTED, "a continuous action space needs DPWSolver (vanilla UCT enumerates actions(mdp, s))");
    if (!cfg->enable_action_pw) return fail(MCTSB200_ERR_UNSUPPORTED, "a continuous action space needs enable_action_pw (its actions cannot be enumerated)");
    if (cfg->rollouts_per_leaf != 1) return fail(MCTSB200_ERR_UNSUPPORTED, "continuous action spaces run one rollout per leaf");
    return MCTSB200_OK;
}

template <class M>
int32_t plan_dispatch(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, const void* roots, bool roots_on_device, long long n_roots,
                      size_t state_bytes, long long base, int* action_out, double* bestq_out, int* status_out, bool out_on_device,
                      mctsb200_plan_stats* stats) {
    using Abi = typename M::AbiState;
    const auto t0 = std::chrono::steady_clock::now();
    if (!MdpAccess<M>::ready(ctx)) return fail(MCTSB200_ERR_STATE, "mdp_configure was not called for this MDP");
    mctsb200_solver_cfg cfg_capped;
    cfg = effective_cfg(cfg, cfg_capped);
    int32_t rc = validate_cfg(cfg, M::kActions, M::kDense, &MdpAccess<M>::policy_ok, M::kActionDoubles > 0);
    if (rc) return rc;
    if (state_bytes != sizeof(Abi)) return fail(MCTSB200_ERR_INVALID_ARG, "state_bytes %zu != %zu", state_bytes, sizeof(Abi));
    if (n_roots <= 0 || !roots) return fail(MCTSB200_ERR_INVALID_ARG, "need at least one root state");
    if (base < 0) return fail(MCTSB200_ERR_INVALID_ARG, "root_index_base must be >= 0");
    CUDA_TRY(cudaSetDevice(ctx->device));

    const Abi* d_roots = nullptr;
    int64_t h2d = 0, d2h = 0;
    if (roots_on_device) {
        d_roots = (const Abi*)roots;
    } else {
        const Abi* h = (const Abi*)roots;
        for (long long i = 0; i < n_roots; ++i)
            if (MdpAccess<M>::check_root(ctx, h[i])) return fail(MCTSB200_ERR_INVALID_ARG, "root state %lld is not a valid state of this MDP", i);
        CUDA_TRY(ctx->io_roots.ensure(sizeof(Abi) * (size_t)n_roots));
        CUDA_TRY(cudaMemcpyAsync(ctx->io_roots.p, roots, sizeof(Abi) * (size_t)n_roots, cudaMemcpyHostToDevice, ctx->stream));
        h2d += (int64_t)(sizeof(Abi) * (size_t)n_roots);
        d_roots = (const Abi*)ctx->io_roots.p;
    }
    int* d_action = action_out;
    double* d_bestq = bestq_out;
    int* d_status = status_out;
    if (!out_on_device) {
        CUDA_TRY(ctx->io_action.ensure(sizeof(int) * (size_t)n_roots));
        CUDA_TRY(ctx->io_bestq.ensure(sizeof(double) * (size_t)n_roots));
        CUDA_TRY(ctx->io_status.ensure(sizeof(int) * (size_t)n_roots));
        d_action = (int*)ctx->io_action.p;
        d_bestq = (double*)ctx->io_bestq.p;
        d_status = (int*)ctx->io_status.p;
    }
    unsigned flags = 0;  // OR of the trees' status, reduced on the device: errors surface even without a status buffer
    if constexpr (M::kActionDoubles > 0) {
        if (int32_t e = check_box_cfg(cfg)) return e;
        rc = run_search<M, MCTSB200_KIND_DPW>(ctx, cfg, d_roots, n_roots, 1, base, d_action, d_bestq, d_status, stats, &flags);
    } else {
        if (cfg->kind == MCTSB200_KIND_UCT) rc = run_search<M, MCTSB200_KIND_UCT>(ctx, cfg, d_roots, n_roots, 1, base, d_action, d_bestq, d_status, stats, &flags);
        else rc = run_search<M, MCTSB200_KIND_DPW>(ctx, cfg, d_roots, n_roots, 1, base, d_action, d_bestq, d_status, stats, &flags);
    }
    if (rc) return rc;
    int32_t worst = (flags & 2u) ? MCTSB200_ERR_CAPACITY : (flags & 1u) ? MCTSB200_ERR_TERMINAL_ROOT : MCTSB200_OK;
    if (!out_on_device) {
        if (action_out) CUDA_TRY(cudaMemcpyAsync(action_out, d_action, sizeof(int) * (size_t)n_roots, cudaMemcpyDeviceToHost, ctx->stream));
        if (bestq_out) CUDA_TRY(cudaMemcpyAsync(bestq_out, d_bestq, sizeof(double) * (size_t)n_roots, cudaMemcpyDeviceToHost, ctx->stream));
        if (status_out) CUDA_TRY(cudaMemcpyAsync(status_out, d_status, sizeof(int) * (size_t)n_roots, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        d2h += (int64_t)((action_out ? sizeof(int) : 0) + (bestq_out ? sizeof(double) : 0) + (status_out ? sizeof(int) : 0)) * n_roots;
    }
    if (stats) {
        stats->h2d_bytes = h2d;
        stats->d2h_bytes = d2h;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (worst == MCTSB200_ERR_TERMINAL_ROOT && !status_out)
        return fail(worst, "MCTS cannot handle terminal states: a root state is terminal (src/dpw.jl:22-27)");
    if (worst == MCTSB200_ERR_CAPACITY) return fail(worst, "a per-tree table overflowed");
    return MCTSB200_OK;
}

template <class M>
int32_t root_parallel_dispatch(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, const void* root_state, size_t state_bytes,
                               long long n_trees, long long base, double* sum_n_out, double* sum_nq_out, double* d_sums_out,
                               mctsb200_plan_stats* stats) {
    using Abi = typename M::AbiState;
    const auto t0 = std::chrono::steady_clock::now();
    if constexpr (M::kActionDoubles > 0) {
        return fail(MCTSB200_ERR_UNSUPPORTED, "root-parallel sums are defined per action of a finite action set");
    } else {
    if (!MdpAccess<M>::ready(ctx)) return fail(MCTSB200_ERR_STATE, "mdp_configure was not called for this MDP");
    mctsb200_solver_cfg cfg_capped;
    cfg = effective_cfg(cfg, cfg_capped);
    int32_t rc = validate_cfg(cfg, M::kActions, M::kDense, &MdpAccess<M>::policy_ok);
    if (rc) return rc;
    // (a rank of a communicator may hold no tree of a small ensemble: it still takes part in the allreduce)
    const bool in_comm = ctx->comm != nullptr && ctx->comm_ranks > 1;
    if (state_bytes != sizeof(Abi) || !root_state || n_trees < 0 || (n_trees == 0 && !in_comm)) return fail(MCTSB200_ERR_INVALID_ARG, "bad root_parallel arguments");
    if (MdpAccess<M>::check_root(ctx, *(const Abi*)root_state)) return fail(MCTSB200_ERR_INVALID_ARG, "root state is not a valid state of this MDP");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(ctx->io_roots.ensure(sizeof(Abi)));
    CUDA_TRY(cudaMemcpyAsync(ctx->io_roots.p, root_state, sizeof(Abi), cudaMemcpyHostToDevice, ctx->stream));
    unsigned flags = 0;
    if (n_trees == 0) {
        if (stats) std::memset(stats, 0, sizeof *stats);
    } else if (cfg->kind == MCTSB200_KIND_UCT)
        rc = run_search<M, MCTSB200_KIND_UCT>(ctx, cfg, (const Abi*)ctx->io_roots.p, n_trees, 0, base, nullptr, nullptr, nullptr, stats, &flags);
    else
        rc = run_search<M, MCTSB200_KIND_DPW>(ctx, cfg, (const Abi*)ctx->io_roots.p, n_trees, 0, base, nullptr, nullptr, nullptr, stats, &flags);
    if (rc) return rc;
    if (flags & 2u) return fail(MCTSB200_ERR_CAPACITY, "a per-tree table overflowed");
    if (flags & 1u) return fail(MCTSB200_ERR_TERMINAL_ROOT, "MCTS cannot handle terminal states: the root state is terminal (src/dpw.jl:22-27)");
    constexpr int W = kRootSumsWords(M::kActions);
    CUDA_TRY(ctx->sums.ensure(sizeof(long long) * W));
    unsigned long long* d_words = (unsigned long long*)ctx->sums.p;
    CUDA_TRY(cudaMemsetAsync(d_words, 0, sizeof(long long) * W, ctx->stream));
    int launches = 0;
    if (n_trees > 0) {
        if (ctx->st.finalize_fn && (rc = ctx->st.finalize_fn(ctx))) return rc;
        auto kernel = root_sums_kernel<M>;
        MCTSB200_LAUNCH(kernel, (unsigned)((n_trees + 255) / 256), 256, 0, ctx->stream, (const TreeHeader*)ctx->st.hdr.p, (const Row<M>*)ctx->st.rows.p,
                        ctx->st.geo.NS, n_trees, d_words);
        CUDA_TRY(cudaGetLastError());
        ++launches;
    }
#ifndef MCTSB200_HOST_EMU
    if (ctx->comm && ctx->comm_ranks > 1) {
        // the only inter-GPU traffic of a decision: one allreduce of 5*|A|+1 int64 words over NVLink, in place, on the search's stream
        NCCL_TRY(g_nccl.AllReduce(d_words, d_words, W, ncclInt64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
        ++launches;
    }
#endif
    long long h[W];
    CUDA_TRY(cudaMemcpyAsync(h, d_words, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->root_words.assign(h, h + W);  // (mctsb200_root_sums_words: hosts that merge with their own transport)
    double sums[2 * M::kActions];
    root_sums_to_double(h, M::kActions, sums, sums + M::kActions);
    for (int i = 0; i < M::kActions; ++i) {
        if (sum_n_out) sum_n_out[i] = sums[i];
        if (sum_nq_out) sum_nq_out[i] = sums[M::kActions + i];
    }
    if (d_sums_out) CUDA_TRY(cudaMemcpy(d_sums_out, sums, sizeof sums, cudaMemcpyHostToDevice));
    if (stats) {
        stats->n_kernel_launches += launches;
        stats->h2d_bytes = sizeof(Abi);
        stats->d2h_bytes = sizeof h;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return MCTSB200_OK;
    }
}

// ---- episode driver ------------------------------------------------------------------------------
template <class M>
int32_t episodes_dispatch(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, const void* initial_states, long long n, size_t state_bytes,
                          int max_steps, unsigned long long env_seed, long long base, double* return_out, int32_t* steps_out,
                          void* final_states_out, mctsb200_plan_stats* stats) {
    using Abi = typename M::AbiState;
    const auto t0 = std::chrono::steady_clock::now();
    if (!MdpAccess<M>::ready(ctx)) return fail(MCTSB200_ERR_STATE, "mdp_configure was not called for this MDP");
    mctsb200_solver_cfg cfg_capped;
    cfg = effective_cfg(cfg, cfg_capped);
    int32_t rc = validate_cfg(cfg, M::kActions, M::kDense, &MdpAccess<M>::policy_ok, M::kActionDoubles > 0);
    if (rc) return rc;
    if constexpr (M::kActionDoubles > 0) {
        if (int32_t e = check_box_cfg(cfg)) return e;
    }
    if (state_bytes != sizeof(Abi) || !initial_states || n <= 0 || max_steps < 0 || base < 0) return fail(MCTSB200_ERR_INVALID_ARG, "bad run_episodes arguments");
    const Abi* h = (const Abi*)initial_states;
    for (long long i = 0; i < n; ++i)
        if (MdpAccess<M>::check_root(ctx, h[i])) return fail(MCTSB200_ERR_INVALID_ARG, "start state %lld is not a valid state of this MDP", i);
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(ctx->ep_states.ensure(sizeof(Abi) * (size_t)n));
    CUDA_TRY(ctx->ep_disc.ensure(sizeof(double) * (size_t)n));
    CUDA_TRY(ctx->ep_ret.ensure(sizeof(double) * (size_t)n));
    CUDA_TRY(ctx->ep_steps.ensure(sizeof(int) * (size_t)n));
    CUDA_TRY(ctx->ep_active.ensure(sizeof(int) * (size_t)n));
    CUDA_TRY(ctx->ep_count.ensure(sizeof(int) * ((size_t)max_steps + 1)));
    CUDA_TRY(ctx->io_action.ensure(sizeof(int) * (size_t)n));
    CUDA_TRY(ctx->io_status.ensure(sizeof(int) * (size_t)n));
    Abi* d_s = (Abi*)ctx->ep_states.p;
    int* d_cnt = (int*)ctx->ep_count.p;  // d_cnt[t] = episodes still running before step t
    CUDA_TRY(cudaMemcpyAsync(d_s, initial_states, sizeof(Abi) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(int) * ((size_t)max_steps + 1), ctx->stream));
    const unsigned grid = (unsigned)((n + 255) / 256);
    const typename M::Params mdp = MdpAccess<M>::params(ctx);
#ifndef MCTSB200_HOST_EMU
    if constexpr (M::kCustom) {
        const void* k = plugin_kernel_or_fail(ctx, cfg->kind, mctsb200_rt::K_EPISODE_INIT);
        if (!k) return MCTSB200_ERR_CUDA;
        const Abi* a_s = d_s;
        long long a_n = n;
        double *a_disc = (double*)ctx->ep_disc.p, *a_ret = (double*)ctx->ep_ret.p;
        int *a_steps = (int*)ctx->ep_steps.p, *a_active = (int*)ctx->ep_active.p, *a_cnt = d_cnt;
        void* params[] = {(void*)&mdp, &a_s, &a_n, &a_disc, &a_ret, &a_steps, &a_active, &a_cnt};
        CUDA_TRY(cudaLaunchKernel(k, dim3(grid), dim3(256), params, 0, ctx->stream));
    } else
#endif
    {
        auto k = episode_init_kernel<M>;
        MCTSB200_LAUNCH(k, grid, 256, 0, ctx->stream, mdp, (const Abi*)d_s, n, (double*)ctx->ep_disc.p, (double*)ctx->ep_ret.p, (int*)ctx->ep_steps.p,
                        (int*)ctx->ep_active.p, d_cnt);
        CUDA_TRY(cudaGetLastError());
    }
    mctsb200_plan_stats tot;
    std::memset(&tot, 0, sizeof tot);
    int launches = 1;
    for (int t = 0; t < max_steps; ++t) {
        int running = 0;
        CUDA_TRY(cudaMemcpyAsync(&running, d_cnt + t, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (running == 0) break;
        mctsb200_solver_cfg c = *cfg;
        c.seed = cfg->seed + (uint64_t)t;
        mctsb200_plan_stats st;
        unsigned flags = 0;
        if constexpr (M::kActionDoubles == 0) {
            if (cfg->kind == MCTSB200_KIND_UCT)
                rc = run_search<M, MCTSB200_KIND_UCT>(ctx, &c, d_s, n, 1, base, (int*)ctx->io_action.p, nullptr, (int*)ctx->io_status.p, &st, &flags);
        }
        if (cfg->kind == MCTSB200_KIND_DPW)
            rc = run_search<M, MCTSB200_KIND_DPW>(ctx, &c, d_s, n, 1, base, (int*)ctx->io_action.p, nullptr, (int*)ctx->io_status.p, &st, &flags);
        if (rc) return rc;
        // (a finished episode's root is terminal: expected here; an overflowed table is not)
        if (flags & 2u) return fail(MCTSB200_ERR_CAPACITY, "a per-tree table overflowed at episode step %d", t);
        tot.n_simulations += st.n_simulations; tot.n_tree_levels += st.n_tree_levels; tot.n_dpw_children_seen += st.n_dpw_children_seen;
        tot.n_state_inserts += st.n_state_inserts; tot.n_action_inserts += st.n_action_inserts; tot.n_rollouts += st.n_rollouts;
        tot.n_rollout_steps += st.n_rollout_steps; tot.n_transition_samples += st.n_transition_samples; tot.n_select_exact += st.n_select_exact;
        tot.algorithmic_bytes += st.algorithmic_bytes; tot.kernel_ms += st.kernel_ms;
        tot.lanes_per_tree = st.lanes_per_tree; tot.kernel_variant = st.kernel_variant; tot.iterations_done = st.iterations_done;
        launches += st.n_kernel_launches + 1;
#ifndef MCTSB200_HOST_EMU
        if constexpr (M::kCustom) {
            const void* k = plugin_kernel_or_fail(ctx, cfg->kind, mctsb200_rt::K_EPISODE_STEP);
            if (!k) return MCTSB200_ERR_CUDA;
            Abi* a_s = d_s;
            const int* a_act = (const int*)ctx->io_action.p;
            const double* a_vec = (const double*)ctx->st.action_vec.p;  // (continuous action spaces: the decisions of this step's search)
            long long a_n = n, a_base = base;
            unsigned long long a_seed = env_seed;
            int a_t = t;
            double *a_disc = (double*)ctx->ep_disc.p, *a_ret = (double*)ctx->ep_ret.p;
            int *a_steps = (int*)ctx->ep_steps.p, *a_active = (int*)ctx->ep_active.p, *a_cnt = d_cnt + t + 1;
            void* params[] = {(void*)&mdp, &a_s, &a_act, &a_vec, &a_n, &a_seed, &a_base, &a_t, &a_disc, &a_ret, &a_steps, &a_active, &a_cnt};
            CUDA_TRY(cudaLaunchKernel(k, dim3(grid), dim3(256), params, 0, ctx->stream));
        } else
#endif
        {
            auto k = episode_step_kernel<M>;
            MCTSB200_LAUNCH(k, grid, 256, 0, ctx->stream, mdp, d_s, (const int*)ctx->io_action.p, (const double*)ctx->st.action_vec.p, n, env_seed, base, t, (double*)ctx->ep_disc.p,
                            (double*)ctx->ep_ret.p, (int*)ctx->ep_steps.p, (int*)ctx->ep_active.p, d_cnt + t + 1);
            CUDA_TRY(cudaGetLastError());
        }
    }
    if (return_out) CUDA_TRY(cudaMemcpyAsync(return_out, ctx->ep_ret.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (steps_out) CUDA_TRY(cudaMemcpyAsync(steps_out, ctx->ep_steps.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (final_states_out) CUDA_TRY(cudaMemcpyAsync(final_states_out, d_s, sizeof(Abi) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) {
        tot.n_trees = n;
        tot.n_kernel_launches = launches;
        tot.h2d_bytes = (int64_t)(sizeof(Abi) * (size_t)n);
        tot.d2h_bytes = (int64_t)((return_out ? 8 : 0) + (steps_out ? 4 : 0) + (final_states_out ? sizeof(Abi) : 0)) * n;
        tot.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *stats = tot;
    }
    return MCTSB200_OK;
}

// ---- tree export --------------------------------------------------------------------------------
template <class M>
struct HostTree {
    TreeHeader h;
    std::vector<Row<M>> rows;
    std::vector<Aux<M>> aux;
    std::vector<TransRec> trans;
};

template <class M>
int32_t fetch_tree(mctsb200_ctx* ctx, long long i, HostTree<M>& t) {
    Storage& st = ctx->st;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpy(&t.h, (const TreeHeader*)st.hdr.p + i, sizeof(TreeHeader), cudaMemcpyDeviceToHost));
    t.rows.resize((size_t)t.h.n_state);
    std::vector<int> cell_of;  // direct layout: dense state index of node id (insertion order)
    if (st.direct) {
        // rows are stored by dense state index; bring them into insertion order and restore the key
        if constexpr (M::kDense) {
            std::vector<Row<M>> all((size_t)st.geo.NS);
            CUDA_TRY(cudaMemcpy(all.data(), (const Row<M>*)st.rows.p + i * st.geo.NS, sizeof(Row<M>) * all.size(), cudaMemcpyDeviceToHost));
            cell_of.assign(t.rows.size(), -1);
            for (size_t cell = 0; cell < all.size(); ++cell) {
                const size_t seq = (size_t)all[cell].key;
                if (seq == 0) continue;  // the state is not in the tree
                if (seq > t.rows.size()) return fail(MCTSB200_ERR_STATE, "corrupt direct-indexed node table");
                t.rows[seq - 1] = all[cell];
                t.rows[seq - 1].key = (typename M::Key)cell;
                cell_of[seq - 1] = (int)cell;
            }
        }
    } else if (t.h.n_state)
        CUDA_TRY(cudaMemcpy(t.rows.data(), (const Row<M>*)st.rows.p + i * st.geo.NS, sizeof(Row<M>) * (size_t)t.h.n_state, cudaMemcpyDeviceToHost));
    if (st.kind == MCTSB200_KIND_DPW) {
        t.aux.resize((size_t)t.h.n_state);
        t.trans.resize((size_t)t.h.n_trans);
        if (st.direct) {
            std::vector<Aux<M>> all((size_t)st.geo.NS);
            CUDA_TRY(cudaMemcpy(all.data(), (const Aux<M>*)st.aux.p + i * st.geo.NS, sizeof(Aux<M>) * all.size(), cudaMemcpyDeviceToHost));
            for (size_t nid = 0; nid < t.aux.size(); ++nid)
                if (cell_of[nid] >= 0) t.aux[nid] = all[(size_t)cell_of[nid]];
        } else if (t.h.n_state)
            CUDA_TRY(cudaMemcpy(t.aux.data(), (const Aux<M>*)st.aux.p + i * st.geo.NS, sizeof(Aux<M>) * (size_t)t.h.n_state, cudaMemcpyDeviceToHost));
        if (t.h.n_trans) CUDA_TRY(cudaMemcpy(t.trans.data(), (const TransRec*)st.trans.p + i * st.geo.NT, sizeof(TransRec) * (size_t)t.h.n_trans, cudaMemcpyDeviceToHost));
    }
    return MCTSB200_OK;
}

// A continuous-action tree in the reference's DPWTree vectors: children lists from the arena, action nodes in id order with
// their vector labels, transitions[sanode] as distinct records with multiplicities (oldest first).
template <class M>
int32_t box_tree_export(mctsb200_ctx* ctx, long long i, const mctsb200_tree_buffers* b) {
    Storage& st = ctx->st;
    HostTree<M> t;
    int32_t rc = fetch_tree<M>(ctx, i, t);
    if (rc) return rc;
    const int ns = t.h.n_state, na = t.h.n_action;
    std::vector<BoxANode> an((size_t)na);
    std::vector<int> plog((size_t)t.h.n_plog);
    if (na) CUDA_TRY(cudaMemcpy(an.data(), (const BoxANode*)st.anodes.p + i * st.geo.NA, sizeof(BoxANode) * (size_t)na, cudaMemcpyDeviceToHost));
    if (t.h.n_plog) CUDA_TRY(cudaMemcpy(plog.data(), (const int*)st.plog.p + i * st.geo.NP, sizeof(int) * plog.size(), cudaMemcpyDeviceToHost));
    const auto& P = MdpAccess<M>::params(ctx);
    using Abi = typename M::AbiState;
    int64_t child_pos = 0;
    if (b->child_offsets) b->child_offsets[0] = 0;
    for (int s = 0; s < ns; ++s) {
        if (b->total_n) b->total_n[s] = t.rows[(size_t)s].total_n;
        if (b->s_labels) {
            const Abi lab = M::store(P, M::state_of(P, t.rows[(size_t)s].key));
            std::memcpy((char*)b->s_labels + sizeof(Abi) * (size_t)s, &lab, sizeof lab);
        }
        const Aux<M>& ax = t.aux[(size_t)s];
        for (int j = 0; j < ax.nac[0]; ++j) {
            if (b->child_ids) b->child_ids[child_pos] = (int64_t)plog[(size_t)(ax.pbase[0] + j)] + 1;
            ++child_pos;
        }
        if (b->child_offsets) b->child_offsets[s + 1] = child_pos;
    }
    int64_t pos = 0;
    if (b->trans_offsets) b->trans_offsets[0] = 0;
    std::vector<int> chain;
    for (int id = 0; id < na; ++id) {
        const BoxANode& c = an[(size_t)id];
        if (b->n) b->n[id] = c.n;
        if (b->q) b->q[id] = c.q;
        if (b->a_labels) b->a_labels[id] = -1;
        if (b->a_label_vectors)
            for (int k = 0; k < M::kActionDoubles; ++k) b->a_label_vectors[(size_t)id * M::kActionDoubles + k] = c.a[k];
        if (b->n_a_children) b->n_a_children[id] = c.nac;
        chain.clear();
        for (int rec = c.thead; rec >= 0; rec = t.trans[(size_t)rec].next) chain.push_back(rec);
        for (auto it = chain.rbegin(); it != chain.rend(); ++it) {
            const TransRec& tr = t.trans[(size_t)*it];
            if (b->trans_spnode) b->trans_spnode[pos] = tr.sp + 1;
            if (b->trans_r) b->trans_r[pos] = tr.r;
            if (b->trans_count) b->trans_count[pos] = tr.count;
            ++pos;
        }
        if (b->trans_offsets) b->trans_offsets[id + 1] = pos;
    }
    return MCTSB200_OK;
}

template <class M>
int32_t vis_stats_impl(mctsb200_ctx* ctx, long long i, int64_t* said, int64_t* spid, int64_t* count, int64_t capacity, int64_t* n_out) {
    Storage& st = ctx->st;
    if (st.kind != MCTSB200_KIND_UCT || st.vis_branch == 0) return fail(MCTSB200_ERR_STATE, "the last plan call did not record _vis_stats (MCTSSolver.enable_tree_vis)");
    HostTree<M> t;
    int32_t rc = fetch_tree<M>(ctx, i, t);
    if (rc) return rc;
    const int A = M::kActions, B = st.vis_branch;
    std::vector<VisRec> v((size_t)st.geo.NS * A * B);
    CUDA_TRY(cudaMemcpy(v.data(), (const VisRec*)st.vis.p + (size_t)i * v.size(), sizeof(VisRec) * v.size(), cudaMemcpyDeviceToHost));
    std::vector<std::array<int64_t, 3>> out;
    for (int node = 0; node < t.h.n_state; ++node)
        for (int slot = 0; slot < A; ++slot)
            for (int b = 0; b < B; ++b) {
                const VisRec& e = v[((size_t)node * A + slot) * B + b];
                if (e.count > 0) out.push_back({(int64_t)node * A + slot + 1, (int64_t)e.sp + 1, (int64_t)e.count});
            }
    std::sort(out.begin(), out.end());
    *n_out = (int64_t)out.size();
    if (said && spid && count) {
        if (capacity < *n_out) return fail(MCTSB200_ERR_INVALID_ARG, "need room for %lld entries", (long long)*n_out);
        for (size_t k = 0; k < out.size(); ++k) {
            said[k] = out[k][0];
            spid[k] = out[k][1];
            count[k] = out[k][2];
        }
    }
    return MCTSB200_OK;
}

template <class M>
int32_t tree_sizes_impl(mctsb200_ctx* ctx, long long i, mctsb200_tree_sizes* out) {
    TreeHeader h;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpy(&h, (const TreeHeader*)ctx->st.hdr.p + i, sizeof h, cudaMemcpyDeviceToHost));
    out->n_state_nodes = h.n_state;
    out->n_action_nodes = h.n_action;
    out->n_transitions = ctx->st.kind == MCTSB200_KIND_DPW ? h.n_trans : 0;
    out->state_bytes = (int64_t)sizeof(typename M::AbiState);
    out->action_doubles = M::kActionDoubles;
    return MCTSB200_OK;
}

template <class M>
int32_t tree_export_impl(mctsb200_ctx* ctx, long long i, const mctsb200_tree_buffers* b) {
    if constexpr (M::kActionDoubles > 0) return box_tree_export<M>(ctx, i, b);
    constexpr int A = M::kActions;
    HostTree<M> t;
    int32_t rc = fetch_tree<M>(ctx, i, t);
    if (rc) return rc;
    const auto& P = MdpAccess<M>::params(ctx);
    const bool dpw = ctx->st.kind == MCTSB200_KIND_DPW;
    const int ns = t.h.n_state;
    using Abi = typename M::AbiState;
    int64_t child_pos = 0;
    // reference action-node id (1-based) of child `slot` of state node `s`
    auto said = [&](int s, int slot) -> int64_t { return dpw ? (int64_t)t.aux[(size_t)s].seq[slot] : (int64_t)s * A + slot + 1; };
    if (b->child_offsets) b->child_offsets[0] = 0;
    for (int s = 0; s < ns; ++s) {
        const Row<M>& rw = t.rows[(size_t)s];
        if (b->total_n) b->total_n[s] = rw.total_n;
        if (b->s_labels) {
            const Abi lab = M::store(P, M::state_of(P, rw.key));
            std::memcpy((char*)b->s_labels + sizeof(Abi) * (size_t)s, &lab, sizeof lab);
        }
        for (int slot = 0; slot < rw.nchild(); ++slot) {
            const int64_t id = said(s, slot);
            if (b->child_ids) b->child_ids[child_pos] = id;
            ++child_pos;
            if (b->n) b->n[id - 1] = rw.n[slot];
            if (b->q) b->q[id - 1] = rw.q[slot];
            if (b->a_labels) b->a_labels[id - 1] = rw.alabel(slot);
            if (dpw && b->n_a_children) b->n_a_children[id - 1] = t.aux[(size_t)s].nac[slot];
        }
        if (b->child_offsets) b->child_offsets[s + 1] = child_pos;
    }
    if (dpw && (b->trans_offsets || b->trans_spnode || b->trans_r || b->trans_count)) {
        // CSR over action nodes in id order; records oldest-first (first-occurrence order)
        const int na = t.h.n_action;
        std::vector<int> owner_s((size_t)na, -1), owner_slot((size_t)na, 0);
        for (int s = 0; s < ns; ++s)
            for (int slot = 0; slot < t.rows[(size_t)s].nchild(); ++slot) {
                const int64_t id = said(s, slot);
                owner_s[(size_t)(id - 1)] = s;
                owner_slot[(size_t)(id - 1)] = slot;
            }
        int64_t pos = 0;
        if (b->trans_offsets) b->trans_offsets[0] = 0;
        std::vector<int> chain;
        for (int id = 0; id < na; ++id) {
            chain.clear();
            if (owner_s[(size_t)id] >= 0)
                for (int rec = t.aux[(size_t)owner_s[(size_t)id]].thead[owner_slot[(size_t)id]]; rec >= 0; rec = t.trans[(size_t)rec].next) chain.push_back(rec);
            for (auto it = chain.rbegin(); it != chain.rend(); ++it) {
                const TransRec& tr = t.trans[(size_t)*it];
                if (b->trans_spnode) b->trans_spnode[pos] = tr.sp + 1;
                if (b->trans_r) b->trans_r[pos] = tr.r;
                if (b->trans_count) b->trans_count[pos] = tr.count;
                ++pos;
            }
            if (b->trans_offsets) b->trans_offsets[id + 1] = pos;
        }
    }
    return MCTSB200_OK;
}

template <class M>
// capacity: entries the caller's arrays hold (-1: the legacy entry point, "at least n_actions" -- which a continuous action space
// cannot promise)
int32_t root_stats_impl(mctsb200_ctx* ctx, long long i, long long capacity, int32_t* n_children, int32_t* action_idx, int64_t* n, double* q,
                        int64_t* total_n) {
    TreeHeader h;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpy(&h, (const TreeHeader*)ctx->st.hdr.p + i, sizeof h, cudaMemcpyDeviceToHost));
    if (h.n_state == 0) {
        if (n_children) *n_children = 0;
        if (total_n) *total_n = 0;
        return MCTSB200_OK;
    }
    if constexpr (M::kActionDoubles > 0) {  // children listed in the arena; action_idx = position among the root's children
        Row<M> rw0;
        Aux<M> ax;
        CUDA_TRY(cudaMemcpy(&rw0, (const Row<M>*)ctx->st.rows.p + i * ctx->st.geo.NS + h.root, sizeof rw0, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(&ax, (const Aux<M>*)ctx->st.aux.p + i * ctx->st.geo.NS + h.root, sizeof ax, cudaMemcpyDeviceToHost));
        if (n_children) *n_children = ax.nac[0];
        if (total_n) *total_n = rw0.total_n;
        if (!action_idx && !n && !q) return MCTSB200_OK;  // the count query of the two-call pattern
        if (capacity < 0)
            return fail(MCTSB200_ERR_UNSUPPORTED, "a continuous action space puts no bound on a root's children (%d here): use mctsb200_root_stats_n", ax.nac[0]);
        const int n_out = (int)std::min<long long>(capacity, ax.nac[0]);
        std::vector<int> kids((size_t)n_out);
        if (n_out) CUDA_TRY(cudaMemcpy(kids.data(), (const int*)ctx->st.plog.p + i * ctx->st.geo.NP + ax.pbase[0], sizeof(int) * kids.size(), cudaMemcpyDeviceToHost));
        for (int j = 0; j < n_out; ++j) {
            BoxANode c;
            CUDA_TRY(cudaMemcpy(&c, (const BoxANode*)ctx->st.anodes.p + i * ctx->st.geo.NA + kids[(size_t)j], sizeof c, cudaMemcpyDeviceToHost));
            if (action_idx) action_idx[j] = j;
            if (n) n[j] = c.n;
            if (q) q[j] = c.q;
        }
        return MCTSB200_OK;
    } else {
        Row<M> rw;
        CUDA_TRY(cudaMemcpy(&rw, (const Row<M>*)ctx->st.rows.p + i * ctx->st.geo.NS + h.root, sizeof rw, cudaMemcpyDeviceToHost));
        if (n_children) *n_children = rw.nchild();
        if (total_n) *total_n = rw.total_n;
        const int n_out = capacity < 0 ? rw.nchild() : (int)std::min<long long>(capacity, rw.nchild());
        for (int j = 0; j < n_out; ++j) {
            if (action_idx) action_idx[j] = rw.alabel(j);
            if (n) n[j] = rw.n[j];
            if (q) q[j] = rw.q[j];
        }
        return MCTSB200_OK;
    }
}

__global__ void probe_log_kernel(const double* x, long long n, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = det_log(x[i]);
}
__global__ void probe_cos_kernel(const double* x, long long n, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = det_cos(x[i]);
}
__global__ void probe_philox_kernel(uint64_t seed, uint64_t tree, uint32_t iteration, uint32_t stream_id, int n_words, uint32_t* out) {
    if (threadIdx.x || blockIdx.x) return;
    Stream s;
    s.init(seed, tree, iteration, stream_id);
    for (int i = 0; i < n_words; ++i) out[i] = s.next();
}

int32_t check_tree_query(mctsb200_ctx* ctx, long long i) {
    if (!ctx) return fail(MCTSB200_ERR_INVALID_ARG, "ctx is NULL");
    if (!ctx->st.valid) return fail(MCTSB200_ERR_STATE, "no tree: call mctsb200_plan first");
    if (i < 0 || i >= ctx->st.n_trees) return fail(MCTSB200_ERR_INVALID_ARG, "root index %lld out of range [0,%lld)", i, ctx->st.n_trees);
    CUDA_TRY(cudaSetDevice(ctx->device));                       // (a child of a multi-device ctx: its stream belongs to its own device)
    if (ctx->st.finalize_fn) return ctx->st.finalize_fn(ctx);  // first tree query after a plan call: write the export tables
    return MCTSB200_OK;
}


// ---- multi-GPU fan-out (mctsb200_ctx_create_multi) ------------------------------------------------------------------------
// contiguous block of shard k of n items over w shards: sizes differ by at most one, earlier shards take the remainder
void shard_bounds(long long n, int w, std::vector<long long>& lo) {
    lo.assign((size_t)w + 1, 0);
    const long long base = n / w, rem = n % w;
    for (int k = 0; k < w; ++k) lo[(size_t)k + 1] = lo[(size_t)k] + base + (k < rem ? 1 : 0);
}

void add_stats(mctsb200_plan_stats& tot, const mctsb200_plan_stats& s) {
    tot.n_trees += s.n_trees; tot.n_simulations += s.n_simulations; tot.n_tree_levels += s.n_tree_levels;
    tot.n_dpw_children_seen += s.n_dpw_children_seen; tot.n_state_inserts += s.n_state_inserts; tot.n_action_inserts += s.n_action_inserts;
    tot.n_rollouts += s.n_rollouts; tot.n_rollout_steps += s.n_rollout_steps; tot.n_transition_samples += s.n_transition_samples;
    tot.n_select_exact += s.n_select_exact; tot.algorithmic_bytes += s.algorithmic_bytes;
    tot.kernel_ms = std::max(tot.kernel_ms, s.kernel_ms);  // the devices run side by side
    tot.h2d_bytes += s.h2d_bytes; tot.d2h_bytes += s.d2h_bytes; tot.n_kernel_launches += s.n_kernel_launches;
    if (s.n_trees) {
        tot.lanes_per_tree = s.lanes_per_tree;
        tot.kernel_variant = s.kernel_variant;
        tot.iterations_done = tot.iterations_done ? std::min(tot.iterations_done, s.iterations_done) : s.iterations_done;
    }
}

// fn(k, child) on one host thread per child; the first failing child's status and message are returned
template <class F>
int32_t fan_out(mctsb200_ctx* g, F&& fn) {
    const size_t n = g->kids.size();
    std::vector<int32_t> rc(n, MCTSB200_OK);
    std::vector<std::string> err(n);
    auto body = [&](size_t k) {
        rc[k] = fn((int)k, g->kids[k]);
        if (rc[k]) err[k] = g_err;  // (thread local)
    };
#ifdef MCTSB200_HOST_EMU
    for (size_t k = 0; k < n; ++k) body(k);  // (the emulator's "shared memory" is one static buffer: one kernel at a time)
#else
    std::vector<std::thread> th;
    for (size_t k = 1; k < n; ++k) th.emplace_back(body, k);
    body(0);
    for (auto& t : th) t.join();
#endif
    for (size_t k = 0; k < n; ++k)
        if (rc[k]) {
            g_err = "device " + std::to_string(g->kids[k]->device) + ": " + err[k];
            return rc[k];
        }
    return MCTSB200_OK;
}

// which child holds item i of the last sharded call
int32_t group_locate(mctsb200_ctx* g, long long i, mctsb200_ctx** kid, long long* local) {
    for (size_t k = 0; k + 1 < g->kid_lo.size(); ++k)
        if (i >= g->kid_lo[k] && i < g->kid_lo[k + 1]) {
            *kid = g->kids[k];
            *local = i - g->kid_lo[k];
            return MCTSB200_OK;
        }
    return g->kid_lo.empty() ? fail(MCTSB200_ERR_STATE, "no tree: call mctsb200_plan first") : fail(MCTSB200_ERR_INVALID_ARG, "root index %lld out of range", i);
}

int32_t group_plan(mctsb200_ctx* g, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* roots, int64_t n_roots, size_t state_bytes,
                   int64_t base, int32_t* action_out, double* bestq_out, int32_t* status_out, mctsb200_plan_stats* stats_out) {
    const auto t0 = std::chrono::steady_clock::now();
    if (n_roots <= 0 || !roots) return fail(MCTSB200_ERR_INVALID_ARG, "need at least one root state");
    shard_bounds(n_roots, (int)g->kids.size(), g->kid_lo);
    std::vector<mctsb200_plan_stats> st(g->kids.size());
    for (auto& x : st) std::memset(&x, 0, sizeof x);
    const int32_t rc = fan_out(g, [&](int k, mctsb200_ctx* c) -> int32_t {
        const long long lo = g->kid_lo[(size_t)k], hi = g->kid_lo[(size_t)k + 1];
        if (hi == lo) return MCTSB200_OK;
        // Philox streams are keyed by the GLOBAL root index: the sharded call returns what one device would
        return mctsb200_plan(c, cfg, mdp_id, (const char*)roots + (size_t)lo * state_bytes, hi - lo, state_bytes, base + lo, action_out ? action_out + lo : nullptr,
                             bestq_out ? bestq_out + lo : nullptr, status_out ? status_out + lo : nullptr, &st[(size_t)k]);
    });
    if (stats_out) {
        std::memset(stats_out, 0, sizeof *stats_out);
        for (const auto& x : st) add_stats(*stats_out, x);
        stats_out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return rc;
}

int32_t group_run_episodes(mctsb200_ctx* g, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* initial_states, int64_t n, size_t state_bytes,
                           int32_t max_steps, uint64_t env_seed, int64_t base, double* return_out, int32_t* steps_out, void* final_out,
                           mctsb200_plan_stats* stats_out) {
    const auto t0 = std::chrono::steady_clock::now();
    if (n <= 0 || !initial_states) return fail(MCTSB200_ERR_INVALID_ARG, "bad run_episodes arguments");
    shard_bounds(n, (int)g->kids.size(), g->kid_lo);
    std::vector<mctsb200_plan_stats> st(g->kids.size());
    for (auto& x : st) std::memset(&x, 0, sizeof x);
    const int32_t rc = fan_out(g, [&](int k, mctsb200_ctx* c) -> int32_t {
        const long long lo = g->kid_lo[(size_t)k], hi = g->kid_lo[(size_t)k + 1];
        if (hi == lo) return MCTSB200_OK;
        return mctsb200_run_episodes(c, cfg, mdp_id, (const char*)initial_states + (size_t)lo * state_bytes, hi - lo, state_bytes, max_steps, env_seed, base + lo,
                                     return_out ? return_out + lo : nullptr, steps_out ? steps_out + lo : nullptr,
                                     final_out ? (char*)final_out + (size_t)lo * state_bytes : nullptr, &st[(size_t)k]);
    });
    if (stats_out) {
        std::memset(stats_out, 0, sizeof *stats_out);
        for (const auto& x : st) add_stats(*stats_out, x);
        stats_out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return rc;
}

int32_t group_root_parallel(mctsb200_ctx* g, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* root_state, size_t state_bytes, int64_t n_trees,
                            int64_t base, double* sum_n_out, double* sum_nq_out, double* d_sums_out, mctsb200_plan_stats* stats_out) {
    const auto t0 = std::chrono::steady_clock::now();
    if (d_sums_out) return fail(MCTSB200_ERR_UNSUPPORTED, "d_sums_out belongs to one device: a multi-device ctx merges the sums itself");
    if (n_trees <= 0) return fail(MCTSB200_ERR_INVALID_ARG, "bad root_parallel arguments");
    int32_t n_actions = 0;
    if (int32_t e = mctsb200_mdp_info(mdp_id, &n_actions, nullptr)) return e;
    shard_bounds(n_trees, (int)g->kids.size(), g->kid_lo);
    const bool nccl = g->kids[0]->comm != nullptr;  // else (several ctxs on one device): the host adds the integer words
    std::vector<mctsb200_plan_stats> st(g->kids.size());
    for (auto& x : st) std::memset(&x, 0, sizeof x);
    std::vector<std::vector<double>> sums(g->kids.size(), std::vector<double>(2 * (size_t)n_actions, 0.0));
    const int32_t rc = fan_out(g, [&](int k, mctsb200_ctx* c) -> int32_t {
        const long long lo = g->kid_lo[(size_t)k], hi = g->kid_lo[(size_t)k + 1];
        c->root_words.clear();
        if (hi == lo && !nccl) return MCTSB200_OK;
        // with a communicator every child allreduces in place on its own stream (ONE ncclAllReduce per decision) and returns the
        // merged sums
        return mctsb200_plan_root_parallel(c, cfg, mdp_id, root_state, state_bytes, hi - lo, base + lo, sums[(size_t)k].data(),
                                           sums[(size_t)k].data() + n_actions, nullptr, &st[(size_t)k]);
    });
    if (rc) return rc;
    std::vector<double> merged = sums[0];
    if (!nccl && g->kids.size() > 1) {
        std::vector<long long> w((size_t)(5 * n_actions + 1), 0);
        for (mctsb200_ctx* c : g->kids)
            for (size_t i = 0; i < c->root_words.size() && i < w.size(); ++i) w[i] += c->root_words[i];
        root_sums_to_double(w.data(), n_actions, merged.data(), merged.data() + n_actions);
    }
    for (int a = 0; a < n_actions; ++a) {
        if (sum_n_out) sum_n_out[a] = merged[(size_t)a];
        if (sum_nq_out) sum_nq_out[a] = merged[(size_t)(n_actions + a)];
    }
    if (stats_out) {
        std::memset(stats_out, 0, sizeof *stats_out);
        for (const auto& x : st) add_stats(*stats_out, x);
        stats_out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return MCTSB200_OK;
}

}  // namespace

// =====================================================================================================
// extern "C" surface
// =====================================================================================================
// Runs `return CALL(M);` with M = the trait struct of mdp_id. A plugin id (mctsb200_mdp_register) selects that plugin for the
// call in progress; the host code exists for the shapes listed here (state doubles x actions).
#ifndef MCTSB200_HOST_EMU
using Custom22 = CustomDev<2, 2>;
using Custom23 = CustomDev<2, 3>;
using Custom24 = CustomDev<2, 4>;
using Custom42 = CustomDev<4, 2>;
using Custom43 = CustomDev<4, 3>;
using Custom44 = CustomDev<4, 4>;
using CustomBox21 = CustomBoxDev<2, 1>;
using CustomBox22 = CustomBoxDev<2, 2>;
using CustomBox23 = CustomBoxDev<2, 3>;
using CustomBox24 = CustomBoxDev<2, 4>;
using CustomBox41 = CustomBoxDev<4, 1>;
using CustomBox42 = CustomBoxDev<4, 2>;
using CustomBox43 = CustomBoxDev<4, 3>;
using CustomBox44 = CustomBoxDev<4, 4>;
bool plugin_shape_ok(int sd, int a) { return (sd == 2 || sd == 4) && a >= 2 && a <= 4; }
bool plugin_box_shape_ok(int sd, int ad) { return (sd == 2 || sd == 4) && ad >= 1 && ad <= 4; }
#define MCTSB200_FOR_PLUGIN(ctx, mdp_id, CALL)                                                              \
    if ((mdp_id) >= MCTSB200_MDP_PLUGIN_BASE) {                                                             \
        const mctsb200_rt::PluginInfo* pi__ = mctsb200_rt::plugin_info((mdp_id) - MCTSB200_MDP_PLUGIN_BASE); \
        if (pi__) {                                                                                         \
            (ctx)->cur_plugin = (mdp_id) - MCTSB200_MDP_PLUGIN_BASE;                                        \
            (ctx)->cur_max_branch = pi__->max_branch;                                                       \
            switch (pi__->action_doubles * 100 + pi__->state_doubles * 10 + pi__->n_actions) {              \
                case 120: return CALL(CustomBox21);                                                         \
                case 220: return CALL(CustomBox22);                                                         \
                case 140: return CALL(CustomBox41);                                                         \
                case 240: return CALL(CustomBox42);                                                         \
                case 320: return CALL(CustomBox23);                                                         \
                case 420: return CALL(CustomBox24);                                                         \
                case 340: return CALL(CustomBox43);                                                         \
                case 440: return CALL(CustomBox44);                                                         \
                case 22: return CALL(Custom22);                                                             \
                case 23: return CALL(Custom23);                                                             \
                case 24: return CALL(Custom24);                                                             \
                case 42: return CALL(Custom42);                                                             \
                case 43: return CALL(Custom43);                                                             \
                case 44: return CALL(Custom44);                                                             \
            }                                                                                               \
        }                                                                                                   \
    }
#else
#define MCTSB200_FOR_PLUGIN(ctx, mdp_id, CALL)
#endif
#define MCTSB200_FOR_MDP(ctx, mdp_id, CALL)                                     \
    do {                                                                        \
        if ((mdp_id) == MCTSB200_MDP_GRIDWORLD) return CALL(GridWorldDev);      \
        if ((mdp_id) == MCTSB200_MDP_MOUNTAINCAR) return CALL(MountainCarDev);  \
        if ((mdp_id) == MCTSB200_MDP_GAUSSIAN_BOX) return CALL(GaussianBoxDev); \
        MCTSB200_FOR_PLUGIN(ctx, mdp_id, CALL)                                  \
        return fail(MCTSB200_ERR_INVALID_ARG, "unknown mdp_id %d", (int)(mdp_id)); \
    } while (0)

extern "C" {

int32_t mctsb200_abi_version(void) { return MCTSB200_ABI_VERSION; }
#ifdef MCTSB200_HOST_EMU
// only the tests' host-emulation build exports this marker: the product loader refuses such a library unless a test hands it over
// explicitly (mcts.jl_b200/_lib.py), so no environment variable can turn the product path into a CPU path
int32_t mctsb200_host_emulation(void) { return 1; }
#endif
const char* mctsb200_last_error(void) { return g_err.c_str(); }

int32_t mctsb200_ctx_create(int32_t device_id, void* stream, mctsb200_ctx** out) {
    if (!out) return fail(MCTSB200_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(MCTSB200_ERR_CUDA, "no CUDA device available (%s); libmctsb200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device_id < 0 || device_id >= n_dev) return fail(MCTSB200_ERR_INVALID_ARG, "device %d out of range [0,%d)", device_id, n_dev);
    CUDA_TRY(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10) return fail(MCTSB200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device_id, prop.major, prop.minor);
    mctsb200_ctx* c = new mctsb200_ctx();
    c->device = device_id;
    c->sm_count = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return fail(MCTSB200_ERR_CUDA, "cudaStreamCreate failed");
        }
        c->own_stream = true;
    }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    *out = c;
    return MCTSB200_OK;
}

int32_t mctsb200_ctx_destroy(mctsb200_ctx* c) {
    if (!c) return MCTSB200_OK;
    if (!c->kids.empty()) {  // a multi-device ctx owns its children (and they own their communicator ranks)
        for (mctsb200_ctx* k : c->kids) mctsb200_ctx_destroy(k);
        delete c;
        return MCTSB200_OK;
    }
    cudaSetDevice(c->device);
#ifndef MCTSB200_HOST_EMU
    if (c->comm) {
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy((ncclComm_t)c->comm);
        c->comm = nullptr;
    }
#endif
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->gw_reward, &c->gw_term, &c->st.hdr, &c->st.rows, &c->st.rows_fast, &c->st.trans_fast, &c->st.aux, &c->st.trans, &c->st.map, &c->st.path, &c->st.plog, &c->st.anodes, &c->st.action_vec, &c->st.vis, &c->nmin_action_tab, &c->counters,
                      &c->log_tab, &c->sqrtlog_tab, &c->rsqrt_tab, &c->nmin_state, &c->init_q, &c->init_n, &c->value_tab, &c->io_roots, &c->io_action, &c->io_bestq, &c->io_status,
                      &c->sums, &c->order, &c->order_bins, &c->policy_tab, &c->next_action_tab, &c->gw_fast,
                      &c->ep_states, &c->ep_disc, &c->ep_ret, &c->ep_steps, &c->ep_active, &c->ep_count};
    for (DevBuf* b : bufs) b->release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return MCTSB200_OK;
}

int32_t mctsb200_ctx_create_multi(const int32_t* device_ids, int32_t n_dev, mctsb200_ctx** out) {
    if (!out) return fail(MCTSB200_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!device_ids || n_dev < 1 || n_dev > 64) return fail(MCTSB200_ERR_INVALID_ARG, "need 1..64 device ids");
    bool distinct = true;
    for (int i = 0; i < n_dev; ++i)
        for (int j = 0; j < i; ++j) distinct &= device_ids[i] != device_ids[j];
    mctsb200_ctx* g = new mctsb200_ctx();
    g->device = device_ids[0];
    for (int i = 0; i < n_dev; ++i) {
        mctsb200_ctx* k = nullptr;
        const int32_t rc = mctsb200_ctx_create(device_ids[i], nullptr, &k);
        if (rc) {
            for (mctsb200_ctx* x : g->kids) mctsb200_ctx_destroy(x);
            delete g;
            return rc;
        }
        g->kids.push_back(k);
    }
#ifndef MCTSB200_HOST_EMU
    if (n_dev > 1 && distinct) {
        // one communicator over the devices of this process: root-parallel partial sums are merged with one ncclAllReduce over
        // NVLink per decision (several ctxs on ONE device, a debugging setup, are merged by the host instead)
        std::vector<ncclComm_t> comms((size_t)n_dev);
        std::vector<int> devs(device_ids, device_ids + n_dev);
        ncclResult_t r = ncclSuccess;
        if (!g_nccl.load() || (r = g_nccl.CommInitAll(comms.data(), n_dev, devs.data())) != ncclSuccess) {
            const std::string why = g_nccl.handle ? g_nccl.GetErrorString(r) : g_nccl.error;
            for (mctsb200_ctx* x : g->kids) mctsb200_ctx_destroy(x);
            delete g;
            return fail(MCTSB200_ERR_CUDA, "ncclCommInitAll over %d devices failed: %s", n_dev, why.c_str());
        }
        for (int i = 0; i < n_dev; ++i) {
            g->kids[(size_t)i]->comm = comms[(size_t)i];
            g->kids[(size_t)i]->comm_ranks = n_dev;
            g->kids[(size_t)i]->comm_rank = i;
        }
    }
#endif
    *out = g;
    return MCTSB200_OK;
}

int32_t mctsb200_ctx_device_count(mctsb200_ctx* ctx, int32_t* n_dev_out) {
    if (!ctx || !n_dev_out) return fail(MCTSB200_ERR_INVALID_ARG, "NULL argument");
    *n_dev_out = ctx->kids.empty() ? 1 : (int32_t)ctx->kids.size();
    return MCTSB200_OK;
}

int32_t mctsb200_comm_unique_id(void* id_out, size_t bytes) {
#ifdef MCTSB200_HOST_EMU
    (void)id_out; (void)bytes;
    return fail(MCTSB200_ERR_UNSUPPORTED, "NCCL is a device-side transport");
#else
    if (!id_out || bytes < sizeof(ncclUniqueId)) return fail(MCTSB200_ERR_INVALID_ARG, "id buffer must hold %zu bytes", sizeof(ncclUniqueId));
    if (!g_nccl.load()) return fail(MCTSB200_ERR_CUDA, "%s", g_nccl.error.c_str());
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    std::memset(id_out, 0, bytes);
    std::memcpy(id_out, &id, sizeof id);
    return MCTSB200_OK;
#endif
}

int32_t mctsb200_comm_init(mctsb200_ctx* ctx, const void* unique_id, size_t bytes, int32_t n_ranks, int32_t rank) {
#ifdef MCTSB200_HOST_EMU
    (void)ctx; (void)unique_id; (void)bytes; (void)n_ranks; (void)rank;
    return fail(MCTSB200_ERR_UNSUPPORTED, "NCCL is a device-side transport");
#else
    if (!ctx || !unique_id || bytes < sizeof(ncclUniqueId)) return fail(MCTSB200_ERR_INVALID_ARG, "bad comm_init arguments");
    if (!ctx->kids.empty()) return fail(MCTSB200_ERR_UNSUPPORTED, "a multi-device ctx already owns its communicator");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(MCTSB200_ERR_INVALID_ARG, "rank %d of %d", rank, n_ranks);
    if (ctx->comm) return fail(MCTSB200_ERR_STATE, "this ctx already is a rank of a communicator");
    if (!g_nccl.load()) return fail(MCTSB200_ERR_CUDA, "%s", g_nccl.error.c_str());
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof id);
    ncclComm_t comm = nullptr;
    NCCL_TRY(g_nccl.CommInitRank(&comm, n_ranks, id, rank));
    ctx->comm = comm;
    ctx->comm_ranks = n_ranks;
    ctx->comm_rank = rank;
    return MCTSB200_OK;
#endif
}

int32_t mctsb200_root_sums_words(mctsb200_ctx* ctx, int64_t* words_out, int32_t capacity, int32_t* n_words_out) {
    if (!ctx || !n_words_out) return fail(MCTSB200_ERR_INVALID_ARG, "NULL argument");
    if (!ctx->kids.empty()) return fail(MCTSB200_ERR_UNSUPPORTED, "a multi-device ctx merges the sums itself");
    if (ctx->root_words.empty()) return fail(MCTSB200_ERR_STATE, "no root-parallel call yet");
    *n_words_out = (int32_t)ctx->root_words.size();
    if (words_out) {
        if (capacity < *n_words_out) return fail(MCTSB200_ERR_INVALID_ARG, "need room for %d words", *n_words_out);
        for (size_t i = 0; i < ctx->root_words.size(); ++i) words_out[i] = ctx->root_words[i];
    }
    return MCTSB200_OK;
}

int32_t mctsb200_root_sums_merge(const int64_t* words, int32_t n_actions, double* sum_n_out, double* sum_nq_out) {
    if (!words || n_actions < 1 || n_actions > 64 || !sum_n_out || !sum_nq_out) return fail(MCTSB200_ERR_INVALID_ARG, "bad arguments");
    std::vector<long long> w(words, words + 5 * n_actions + 1);
    root_sums_to_double(w.data(), n_actions, sum_n_out, sum_nq_out);
    return MCTSB200_OK;
}

int32_t mctsb200_mdp_lookup(const char* name, int32_t* mdp_id) {
    if (!name || !mdp_id) return fail(MCTSB200_ERR_INVALID_ARG, "NULL argument");
    if (!std::strcmp(name, "SimpleGridWorld")) *mdp_id = MCTSB200_MDP_GRIDWORLD;
    else if (!std::strcmp(name, "MountainCar")) *mdp_id = MCTSB200_MDP_MOUNTAINCAR;
    else if (!std::strcmp(name, "GaussianBox")) *mdp_id = MCTSB200_MDP_GAUSSIAN_BOX;
    else {
#ifndef MCTSB200_HOST_EMU
        const int idx = mctsb200_rt::lookup_plugin(name);
        if (idx >= 0) {
            *mdp_id = MCTSB200_MDP_PLUGIN_BASE + idx;
            return MCTSB200_OK;
        }
#endif
        return fail(MCTSB200_ERR_UNSUPPORTED,
                    "no device plugin registered for MDP '%s' (built in: SimpleGridWorld, MountainCar; others: mctsb200_mdp_register)", name);
    }
    return MCTSB200_OK;
}

int32_t mctsb200_mdp_register(const char* name, const char* cuda_source, const char* plugin_struct, int32_t state_doubles, int32_t n_actions,
                              int32_t max_branch, int32_t* mdp_id_out) {
#ifdef MCTSB200_HOST_EMU
    return fail(MCTSB200_ERR_UNSUPPORTED, "plugins are compiled for the device only");
#else
    if (!name || !name[0] || !cuda_source || !plugin_struct || !plugin_struct[0] || !mdp_id_out) return fail(MCTSB200_ERR_INVALID_ARG, "NULL or empty argument");
    if (!std::strcmp(name, "SimpleGridWorld") || !std::strcmp(name, "MountainCar")) return fail(MCTSB200_ERR_INVALID_ARG, "'%s' is a built-in MDP", name);
    if (!plugin_shape_ok(state_doubles, n_actions))
        return fail(MCTSB200_ERR_UNSUPPORTED, "plugin shape (%d state doubles, %d actions) is not available: state_doubles 2 or 4, n_actions 2..4", state_doubles,
                    n_actions);
    if (max_branch < 0) return fail(MCTSB200_ERR_INVALID_ARG, "max_branch must be >= 0 (0 = unbounded)");
    mctsb200_rt::PluginInfo info;
    info.name = name;
    info.source = cuda_source;
    info.struct_name = plugin_struct;
    info.state_doubles = state_doubles;
    info.n_actions = n_actions;
    info.max_branch = max_branch;
    std::string err;
    const int idx = mctsb200_rt::register_plugin(info, err);
    if (idx < 0) {
        g_err = err;  // (compiler logs are longer than fail()'s buffer)
        return MCTSB200_ERR_INVALID_ARG;
    }
    *mdp_id_out = MCTSB200_MDP_PLUGIN_BASE + idx;
    return MCTSB200_OK;
#endif
}

int32_t mctsb200_mdp_register_continuous(const char* name, const char* cuda_source, const char* plugin_struct, int32_t state_doubles,
                                         int32_t action_doubles, int32_t* mdp_id_out) {
#ifdef MCTSB200_HOST_EMU
    return fail(MCTSB200_ERR_UNSUPPORTED, "plugins are compiled for the device only");
#else
    if (!name || !name[0] || !cuda_source || !plugin_struct || !plugin_struct[0] || !mdp_id_out) return fail(MCTSB200_ERR_INVALID_ARG, "NULL or empty argument");
    if (!std::strcmp(name, "SimpleGridWorld") || !std::strcmp(name, "MountainCar") || !std::strcmp(name, "GaussianBox"))
        return fail(MCTSB200_ERR_INVALID_ARG, "'%s' is a built-in MDP", name);
    if (!plugin_box_shape_ok(state_doubles, action_doubles))
        return fail(MCTSB200_ERR_UNSUPPORTED, "plugin shape (%d state doubles, %d action doubles) is not available: state_doubles 2 or 4, action_doubles 1 to 4",
                    state_doubles, action_doubles);
    mctsb200_rt::PluginInfo info;
    info.name = name;
    info.source = cuda_source;
    info.struct_name = plugin_struct;
    info.state_doubles = state_doubles;
    info.n_actions = 0;
    info.max_branch = 0;
    info.action_doubles = action_doubles;
    std::string err;
    const int idx = mctsb200_rt::register_plugin(info, err);
    if (idx < 0) {
        g_err = err;
        return MCTSB200_ERR_INVALID_ARG;
    }
    *mdp_id_out = MCTSB200_MDP_PLUGIN_BASE + idx;
    return MCTSB200_OK;
#endif
}

int32_t mctsb200_plugin_cache_stats(int64_t* hits_out, int64_t* misses_out) {
    long long h = 0, m = 0;
#ifndef MCTSB200_HOST_EMU
    mctsb200_rt::plugin_cache_stats(&h, &m);
#endif
    if (hits_out) *hits_out = h;
    if (misses_out) *misses_out = m;
    return MCTSB200_OK;
}

int32_t mctsb200_mdp_info(int32_t mdp_id, int32_t* n_actions, int64_t* state_bytes) {
    if (mdp_id == MCTSB200_MDP_GRIDWORLD) {
        if (n_actions) *n_actions = GridWorldDev::kActions;
        if (state_bytes) *state_bytes = sizeof(GridWorldDev::AbiState);
    } else if (mdp_id == MCTSB200_MDP_MOUNTAINCAR) {
        if (n_actions) *n_actions = MountainCarDev::kActions;
        if (state_bytes) *state_bytes = sizeof(MountainCarDev::AbiState);
    } else if (mdp_id == MCTSB200_MDP_GAUSSIAN_BOX) {
        if (n_actions) *n_actions = 0;  // a continuous action space: actions are vectors (mctsb200_root_action_vectors)
        if (state_bytes) *state_bytes = sizeof(GaussianBoxDev::AbiState);
    } else {
#ifndef MCTSB200_HOST_EMU
        if (const mctsb200_rt::PluginInfo* pi = mctsb200_rt::plugin_info(mdp_id - MCTSB200_MDP_PLUGIN_BASE)) {
            if (n_actions) *n_actions = pi->n_actions;
            if (state_bytes) *state_bytes = (int64_t)sizeof(double) * pi->state_doubles;
            return MCTSB200_OK;
        }
#endif
        return fail(MCTSB200_ERR_INVALID_ARG, "unknown mdp_id %d", mdp_id);
    }
    return MCTSB200_OK;
}

int32_t mctsb200_mdp_configure(mctsb200_ctx* ctx, int32_t mdp_id, const void* pod, size_t bytes) {
    if (!ctx || !pod) return fail(MCTSB200_ERR_INVALID_ARG, "NULL argument");
    if (!ctx->kids.empty()) {
        for (mctsb200_ctx* k : ctx->kids)
            if (int32_t rc = mctsb200_mdp_configure(k, mdp_id, pod, bytes)) return rc;
        ctx->kid_lo.clear();
        return MCTSB200_OK;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (mdp_id == MCTSB200_MDP_GRIDWORLD) {
        if (bytes != sizeof(mctsb200_gridworld_params)) return fail(MCTSB200_ERR_INVALID_ARG, "gridworld params: %zu bytes, expected %zu", bytes, sizeof(mctsb200_gridworld_params));
        const mctsb200_gridworld_params& p = *(const mctsb200_gridworld_params*)pod;
        if (p.nx < 1 || p.ny < 1 || (long long)p.nx * p.ny + 1 > 65535) return fail(MCTSB200_ERR_INVALID_ARG, "grid size out of range");
        if (!(p.tprob >= 0.0 && p.tprob <= 1.0)) return fail(MCTSB200_ERR_INVALID_ARG, "tprob must be in [0,1]");
        if (!(p.discount >= 0.0 && p.discount <= 1.0)) return fail(MCTSB200_ERR_INVALID_ARG, "discount must be in [0,1]");
        if (p.n_rewards < 0 || p.n_rewards > MCTSB200_GW_MAX_SPECIAL || p.n_terminate < 0 || p.n_terminate > MCTSB200_GW_MAX_SPECIAL)
            return fail(MCTSB200_ERR_INVALID_ARG, "too many reward / terminate_from cells (max %d)", MCTSB200_GW_MAX_SPECIAL);
        const int cells = p.nx * p.ny;
        std::vector<GridWorldDev::CellInfo> info((size_t)cells + 1);
        auto inb = [&](int x, int y) { return x >= 1 && x <= p.nx && y >= 1 && y <= p.ny; };
        for (int c = 0; c <= cells; ++c) {
            GridWorldDev::CellInfo ci;
            ci.reward = 0.0;
            ci.term_from = c == cells ? 1u : 0u;
            ci.row = 0;
            if (c < cells) {
                const int x = c % p.nx + 1, y = c / p.nx + 1;
                ci.row = 4u * ((y + 1 > p.ny ? 1u : 0u) | (y - 1 < 1 ? 2u : 0u) | (x - 1 < 1 ? 4u : 0u) | (x + 1 > p.nx ? 8u : 0u));
            }
            info[(size_t)c] = ci;
        }
        for (int i = 0; i < p.n_rewards; ++i) {
            // a Dict key outside the grid can only ever be matched by the terminal state (-1,-1)
            if (inb(p.reward_x[i], p.reward_y[i])) info[(size_t)((p.reward_x[i] - 1) + (p.reward_y[i] - 1) * p.nx)].reward = p.reward_v[i];
            else if (p.reward_x[i] == -1 && p.reward_y[i] == -1) info[(size_t)cells].reward = p.reward_v[i];
        }
        for (int i = 0; i < p.n_terminate; ++i)
            if (inb(p.term_x[i], p.term_y[i])) info[(size_t)((p.term_x[i] - 1) + (p.term_y[i] - 1) * p.nx)].term_from = 1u;
        const double p_other = (1.0 - p.tprob) / (double)(GridWorldDev::kActions - 1);
        std::vector<GridWorldDev::Thresholds> thr(16 * 4);
        gw_thresholds(p.tprob, p_other, thr.data());
        // per-(cell, action) 16-byte entries for uct_fast.cuh; only for grids whose successor deltas fit in 8 bits
        ctx->gw_dev.fast_tt = nullptr;
        for (int i = 0; i < 4; ++i) ctx->gw_dev.fast_rbits[i] = 0;
        if (cells <= 127) {
            std::vector<GridWorldDev::FastEntry> fast(((size_t)cells + 1) * 4);
            const int dlt[5] = {0, p.nx, -p.nx, -1, 1};
            bool ok = true, neg_zero = false;
            for (int c = 0; c <= cells; ++c) {
                const double rw = info[(size_t)c].reward;
                uint64_t bits;
                std::memcpy(&bits, &rw, 8);
                if (bits != 0) ctx->gw_dev.fast_rbits[c >> 5] |= 1u << (c & 31);
                if (bits == 0x8000000000000000ull) neg_zero = true;
                for (int a = 0; a < 4; ++a) {
                    GridWorldDev::FastEntry& e = fast[(size_t)c * 4 + a];
                    e.t[0] = e.t[1] = e.t[2] = 0xffffffffu;
                    e.d = (uint32_t)((cells - c) & 0xff);  // plateau 0 -> terminal state
                    if (info[(size_t)c].term_from) continue;  // Deterministic(terminal state)
                    const GridWorldDev::Thresholds& T = thr[info[(size_t)c].row + a];
                    if (T.t[3] != 0xffffffffu || T.t[0] > T.t[1] || T.t[1] > T.t[2]) ok = false;  // (fast_gen relies on ascending thresholds)
                    e.d = 0;
                    for (int i = 0; i < 3; ++i) e.t[i] = T.t[i];
                    for (int i = 0; i < 4; ++i) e.d |= (uint32_t)(dlt[(T.codes >> (4 * i)) & 15u] & 0xff) << (8 * i);
                }
            }
            // a reward of -0.0 could make a running sum -0.0, for which adding +0.0 is not the identity: always add then
            if (neg_zero)
                for (int i = 0; i < 4; ++i) ctx->gw_dev.fast_rbits[i] = 0xffffffffu;
            ctx->gw_fast_det = true;
            for (const auto& e : fast)
                if (e.t[0] != 0xffffffffu) ctx->gw_fast_det = false;
            if (ok) {
                CUDA_TRY(ctx->gw_fast.ensure(fast.size() * sizeof(GridWorldDev::FastEntry)));
                CUDA_TRY(cudaMemcpy(ctx->gw_fast.p, fast.data(), fast.size() * sizeof(GridWorldDev::FastEntry), cudaMemcpyHostToDevice));
                ctx->gw_dev.fast_tt = (const GridWorldDev::FastEntry*)ctx->gw_fast.p;
            }
        }
        CUDA_TRY(ctx->gw_reward.ensure(info.size() * sizeof(GridWorldDev::CellInfo)));
        CUDA_TRY(ctx->gw_term.ensure(thr.size() * sizeof(GridWorldDev::Thresholds)));
        CUDA_TRY(cudaMemcpy(ctx->gw_reward.p, info.data(), info.size() * sizeof(GridWorldDev::CellInfo), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(ctx->gw_term.p, thr.data(), thr.size() * sizeof(GridWorldDev::Thresholds), cudaMemcpyHostToDevice));
        ctx->gw_host = p;
        ctx->gw_reward_max = 0.0;
        for (const auto& ci : info) ctx->gw_reward_max = std::isfinite(ci.reward) ? std::max(ctx->gw_reward_max, std::fabs(ci.reward)) : INFINITY;
        ctx->gw_dev.nx = p.nx;
        ctx->gw_dev.ny = p.ny;
        ctx->gw_dev.ncells = cells;
        {
            const int d[8] = {0, p.nx, -p.nx, -1, 1, 0, 0, 0};
            for (int i = 0; i < 8; ++i) ctx->gw_dev.delta[i] = d[i];
        }
        ctx->gw_dev.tprob = p.tprob;
        ctx->gw_dev.p_other = p_other;
        ctx->gw_dev.discount = p.discount;
        ctx->gw_dev.cells = (const GridWorldDev::CellInfo*)ctx->gw_reward.p;
        ctx->gw_dev.thr = (const GridWorldDev::Thresholds*)ctx->gw_term.p;
        ctx->gw_ok = true;
        ctx->st.valid = false;
        return MCTSB200_OK;
    }
    if (mdp_id == MCTSB200_MDP_MOUNTAINCAR) {
        if (bytes != sizeof(mctsb200_mountaincar_params)) return fail(MCTSB200_ERR_INVALID_ARG, "mountaincar params: %zu bytes, expected %zu", bytes, sizeof(mctsb200_mountaincar_params));
        const mctsb200_mountaincar_params& p = *(const mctsb200_mountaincar_params*)pod;
        if (!(p.discount >= 0.0 && p.discount <= 1.0)) return fail(MCTSB200_ERR_INVALID_ARG, "discount must be in [0,1]");
        ctx->mc_dev.discount = p.discount;
        ctx->mc_dev.cost = p.cost;
        ctx->mc_dev.jackpot = p.jackpot;
        ctx->mc_ok = true;
        ctx->st.valid = false;
        return MCTSB200_OK;
    }
    if (mdp_id == MCTSB200_MDP_GAUSSIAN_BOX) {
        if (bytes != sizeof(mctsb200_gaussian_box_params)) return fail(MCTSB200_ERR_INVALID_ARG, "gaussian box params: %zu bytes, expected %zu", bytes, sizeof(mctsb200_gaussian_box_params));
        const mctsb200_gaussian_box_params& p = *(const mctsb200_gaussian_box_params*)pod;
        if (!(p.discount >= 0.0 && p.discount <= 1.0)) return fail(MCTSB200_ERR_INVALID_ARG, "discount must be in [0,1]");
        ctx->box_dev.discount = p.discount;
        for (int i = 0; i < MCTSB200_BOX_ACTION_DOUBLES; ++i) {
            if (!(p.a_lo[i] <= p.a_hi[i]) || !std::isfinite(p.a_lo[i]) || !std::isfinite(p.a_hi[i])) return fail(MCTSB200_ERR_INVALID_ARG, "action box bounds must be finite, lo <= hi");
            ctx->box_dev.a_lo[i] = p.a_lo[i];
            ctx->box_dev.a_hi[i] = p.a_hi[i];
        }
        for (int i = 0; i < 3; ++i) {
            if (!(p.var[i] >= 0.0) || !std::isfinite(p.var[i])) return fail(MCTSB200_ERR_INVALID_ARG, "variances must be finite and >= 0");
            ctx->box_dev.sd[i] = std::sqrt(p.var[i]);
        }
        if (!std::isfinite(p.drift) || !std::isfinite(p.cost)) return fail(MCTSB200_ERR_INVALID_ARG, "drift and cost must be finite");
        ctx->box_dev.drift = p.drift;
        ctx->box_dev.cost = p.cost;
        ctx->box_ok = true;
        ctx->st.valid = false;
        return MCTSB200_OK;
    }
#ifndef MCTSB200_HOST_EMU
    if (mctsb200_rt::plugin_info(mdp_id - MCTSB200_MDP_PLUGIN_BASE)) {
        // the plugin's own Params POD, `double discount` first (custom.cuh)
        if (!pod || bytes < sizeof(double) || bytes > (size_t)kCustomParamBytes)
            return fail(MCTSB200_ERR_INVALID_ARG, "plugin params: %zu bytes, expected 8..%d (Params begins with `double discount`)", bytes, kCustomParamBytes);
        double discount;
        std::memcpy(&discount, pod, sizeof discount);
        if (!(discount >= 0.0 && discount <= 1.0)) return fail(MCTSB200_ERR_INVALID_ARG, "discount must be in [0,1]");
        mctsb200_ctx::PluginParams pp;
        std::memset(pp.bytes, 0, sizeof pp.bytes);
        std::memcpy(pp.bytes, pod, bytes);
        ctx->plugin_params[mdp_id - MCTSB200_MDP_PLUGIN_BASE] = pp;
        ctx->st.valid = false;
        return MCTSB200_OK;
    }
#endif
    return fail(MCTSB200_ERR_INVALID_ARG, "unknown mdp_id %d", mdp_id);
}

int32_t mctsb200_plan(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* roots, int64_t n_roots, size_t state_bytes,
                      int64_t root_index_base, int32_t* action_idx_out, double* best_q_out, int32_t* status_out, mctsb200_plan_stats* stats_out) {
    if (!ctx) return fail(MCTSB200_ERR_INVALID_ARG, "ctx is NULL");
    if (!ctx->kids.empty()) return group_plan(ctx, cfg, mdp_id, roots, n_roots, state_bytes, root_index_base, action_idx_out, best_q_out, status_out, stats_out);
#define MCTSB200_CALL(M) plan_dispatch<M>(ctx, cfg, roots, false, n_roots, state_bytes, root_index_base, action_idx_out, best_q_out, status_out, false, stats_out)
    MCTSB200_FOR_MDP(ctx, mdp_id, MCTSB200_CALL);
#undef MCTSB200_CALL
}

int32_t mctsb200_plan_device(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* d_roots, int64_t n_roots,
                             size_t state_bytes, int64_t root_index_base, int32_t* d_action_idx_out, double* d_best_q_out, int32_t* d_status_out,
                             mctsb200_plan_stats* stats_out) {
    if (!ctx) return fail(MCTSB200_ERR_INVALID_ARG, "ctx is NULL");
    if (!ctx->kids.empty()) return fail(MCTSB200_ERR_UNSUPPORTED, "device pointers belong to one device: use mctsb200_plan (host buffers) on a multi-device ctx");
#define MCTSB200_CALL(M) plan_dispatch<M>(ctx, cfg, d_roots, true, n_roots, state_bytes, root_index_base, d_action_idx_out, d_best_q_out, d_status_out, true, stats_out)
    MCTSB200_FOR_MDP(ctx, mdp_id, MCTSB200_CALL);
#undef MCTSB200_CALL
}

int32_t mctsb200_plan_root_parallel(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* root_state, size_t state_bytes,
                                    int64_t n_trees, int64_t tree_index_base, double* sum_n_out, double* sum_nq_out, double* d_sums_out,
                                    mctsb200_plan_stats* stats_out) {
    if (!ctx) return fail(MCTSB200_ERR_INVALID_ARG, "ctx is NULL");
    if (!ctx->kids.empty()) return group_root_parallel(ctx, cfg, mdp_id, root_state, state_bytes, n_trees, tree_index_base, sum_n_out, sum_nq_out, d_sums_out, stats_out);
#define MCTSB200_CALL(M) root_parallel_dispatch<M>(ctx, cfg, root_state, state_bytes, n_trees, tree_index_base, sum_n_out, sum_nq_out, d_sums_out, stats_out)
    MCTSB200_FOR_MDP(ctx, mdp_id, MCTSB200_CALL);
#undef MCTSB200_CALL
}

int32_t mctsb200_set_progress_callback(mctsb200_ctx* ctx, mctsb200_progress_fn fn, void* user, int64_t every) {
    if (!ctx) return fail(MCTSB200_ERR_INVALID_ARG, "ctx is NULL");
    if (ctx->kids.size() > 1) return fail(MCTSB200_ERR_UNSUPPORTED, "host hooks between launches are per device: not available on a multi-device ctx");
    if (ctx->kids.size() == 1) return mctsb200_set_progress_callback(ctx->kids[0], fn, user, every);
    if (fn && every < 1) return fail(MCTSB200_ERR_INVALID_ARG, "a progress callback needs every >= 1 iterations");
    ctx->progress_fn = fn;
    ctx->progress_user = fn ? user : nullptr;
    ctx->progress_every = fn ? every : 0;
    return MCTSB200_OK;
}

int32_t mctsb200_set_timer(mctsb200_ctx* ctx, mctsb200_timer_fn fn, void* user) {
    if (!ctx) return fail(MCTSB200_ERR_INVALID_ARG, "ctx is NULL");
    if (ctx->kids.size() > 1) return fail(MCTSB200_ERR_UNSUPPORTED, "host hooks between launches are per device: not available on a multi-device ctx");
    if (ctx->kids.size() == 1) return mctsb200_set_timer(ctx->kids[0], fn, user);
    ctx->timer_fn = fn;
    ctx->timer_user = fn ? user : nullptr;
    return MCTSB200_OK;
}

int32_t mctsb200_run_episodes(mctsb200_ctx* ctx, const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* initial_states, int64_t n_episodes,
                              size_t state_bytes, int32_t max_steps, uint64_t env_seed, int64_t episode_index_base, double* return_out,
                              int32_t* steps_out, void* final_states_out, mctsb200_plan_stats* stats_out) {
    if (!ctx) return fail(MCTSB200_ERR_INVALID_ARG, "ctx is NULL");
    if (!ctx->kids.empty())
        return group_run_episodes(ctx, cfg, mdp_id, initial_states, n_episodes, state_bytes, max_steps, env_seed, episode_index_base, return_out, steps_out,
                                  final_states_out, stats_out);
#define MCTSB200_CALL(M) \
    episodes_dispatch<M>(ctx, cfg, initial_states, n_episodes, state_bytes, max_steps, env_seed, episode_index_base, return_out, steps_out, final_states_out, stats_out)
    MCTSB200_FOR_MDP(ctx, mdp_id, MCTSB200_CALL);
#undef MCTSB200_CALL
}

int32_t mctsb200_root_parallel_decide(const double* sum_n, const double* sum_nq, int32_t n_actions, double init_q, int32_t* action_idx_out,
                                      double* best_q_out, double* q_out) {
    if (!sum_n || !sum_nq || n_actions < 1) return fail(MCTSB200_ERR_INVALID_ARG, "bad arguments");
    double best = -INFINITY;
    int act = 0;
    for (int a = 0; a < n_actions; ++a) {
        const double q = sum_n[a] > 0.0 ? sum_nq[a] / sum_n[a] : init_q;
        if (q_out) q_out[a] = q;
        if (q > best) {
            best = q;
            act = a;
        }
    }
    if (action_idx_out) *action_idx_out = act;
    if (best_q_out) *best_q_out = best;
    return MCTSB200_OK;
}

static int32_t root_stats_any(mctsb200_ctx* ctx, int64_t root_i, long long capacity, int32_t* n_children, int32_t* action_idx, int64_t* n, double* q,
                             int64_t* total_n) {
    if (ctx && !ctx->kids.empty()) {
        mctsb200_ctx* kid = nullptr;
        long long local = 0;
        if (int32_t e = group_locate(ctx, root_i, &kid, &local)) return e;
        return root_stats_any(kid, local, capacity, n_children, action_idx, n, q, total_n);
    }
    int32_t rc = check_tree_query(ctx, root_i);
    if (rc) return rc;
#define MCTSB200_CALL(M) root_stats_impl<M>(ctx, root_i, capacity, n_children, action_idx, n, q, total_n)
    MCTSB200_FOR_MDP(ctx, ctx->st.mdp_id, MCTSB200_CALL);
#undef MCTSB200_CALL
}
int32_t mctsb200_root_stats(mctsb200_ctx* ctx, int64_t root_i, int32_t* n_children, int32_t* action_idx, int64_t* n, double* q, int64_t* total_n) {
    return root_stats_any(ctx, root_i, -1, n_children, action_idx, n, q, total_n);
}
int32_t mctsb200_root_stats_n(mctsb200_ctx* ctx, int64_t root_i, int64_t capacity, int32_t* n_children, int32_t* action_idx, int64_t* n, double* q,
                              int64_t* total_n) {
    if (capacity < 0) return fail(MCTSB200_ERR_INVALID_ARG, "capacity must be >= 0");
    return root_stats_any(ctx, root_i, capacity, n_children, action_idx, n, q, total_n);
}

int32_t mctsb200_tree_sizes_query(mctsb200_ctx* ctx, int64_t root_i, mctsb200_tree_sizes* sizes) {
    if (ctx && !ctx->kids.empty()) {
        mctsb200_ctx* kid = nullptr;
        long long local = 0;
        if (int32_t e = group_locate(ctx, root_i, &kid, &local)) return e;
        return mctsb200_tree_sizes_query(kid, local, sizes);
    }
    int32_t rc = check_tree_query(ctx, root_i);
    if (rc) return rc;
    if (!sizes) return fail(MCTSB200_ERR_INVALID_ARG, "sizes is NULL");
#define MCTSB200_CALL(M) tree_sizes_impl<M>(ctx, root_i, sizes)
    MCTSB200_FOR_MDP(ctx, ctx->st.mdp_id, MCTSB200_CALL);
#undef MCTSB200_CALL
}

int32_t mctsb200_tree_export(mctsb200_ctx* ctx, int64_t root_i, const mctsb200_tree_buffers* buffers) {
    if (ctx && !ctx->kids.empty()) {
        mctsb200_ctx* kid = nullptr;
        long long local = 0;
        if (int32_t e = group_locate(ctx, root_i, &kid, &local)) return e;
        return mctsb200_tree_export(kid, local, buffers);
    }
    int32_t rc = check_tree_query(ctx, root_i);
    if (rc) return rc;
    if (!buffers) return fail(MCTSB200_ERR_INVALID_ARG, "buffers is NULL");
#define MCTSB200_CALL(M) tree_export_impl<M>(ctx, root_i, buffers)
    MCTSB200_FOR_MDP(ctx, ctx->st.mdp_id, MCTSB200_CALL);
#undef MCTSB200_CALL
}

int32_t mctsb200_tree_vis_stats(mctsb200_ctx* ctx, int64_t root_i, int64_t* said, int64_t* spid, int64_t* count, int64_t capacity, int64_t* n_out) {
    if (!n_out) return fail(MCTSB200_ERR_INVALID_ARG, "n_out is NULL");
    if (ctx && !ctx->kids.empty()) {
        mctsb200_ctx* kid = nullptr;
        long long local = 0;
        if (int32_t e = group_locate(ctx, root_i, &kid, &local)) return e;
        return mctsb200_tree_vis_stats(kid, local, said, spid, count, capacity, n_out);
    }
    int32_t rc = check_tree_query(ctx, root_i);
    if (rc) return rc;
#define MCTSB200_CALL(M) vis_stats_impl<M>(ctx, root_i, said, spid, count, capacity, n_out)
    MCTSB200_FOR_MDP(ctx, ctx->st.mdp_id, MCTSB200_CALL);
#undef MCTSB200_CALL
}

int32_t mctsb200_root_action_vectors(mctsb200_ctx* ctx, int64_t first_root, int64_t n, double* out) {
    if (!ctx || !out || n < 0 || first_root < 0) return fail(MCTSB200_ERR_INVALID_ARG, "bad arguments");
    if (!ctx->kids.empty()) {  // roots are sharded in contiguous blocks: ask each device for its part
        for (int64_t i = 0; i < n;) {
            mctsb200_ctx* kid = nullptr;
            long long local = 0;
            if (int32_t e = group_locate(ctx, first_root + i, &kid, &local)) return e;
            const long long m = std::min<long long>(n - i, kid->st.n_trees - local);
            if (int32_t e = mctsb200_root_action_vectors(kid, local, m, out + i * kid->st.action_doubles)) return e;
            i += m;
        }
        return MCTSB200_OK;
    }
    if (!ctx->st.valid) return fail(MCTSB200_ERR_STATE, "no tree: call mctsb200_plan first");
    const int ad = ctx->st.action_doubles;
    if (ad <= 0) return fail(MCTSB200_ERR_STATE, "the last plan call was not over a continuous action space");
    if (first_root + n > ctx->st.n_trees) return fail(MCTSB200_ERR_INVALID_ARG, "roots [%lld, %lld) out of range", (long long)first_root, (long long)(first_root + n));
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n) CUDA_TRY(cudaMemcpy(out, (const double*)ctx->st.action_vec.p + first_root * ad, sizeof(double) * ad * (size_t)n, cudaMemcpyDeviceToHost));
    return MCTSB200_OK;
}

int32_t mctsb200_clear_trees(mctsb200_ctx* ctx) {
    if (!ctx) return fail(MCTSB200_ERR_INVALID_ARG, "ctx is NULL");
    for (mctsb200_ctx* k : ctx->kids) k->st.valid = false;
    ctx->kid_lo.clear();
    ctx->st.valid = false;
    return MCTSB200_OK;
}

int32_t mctsb200_device_log(mctsb200_ctx* ctx, const double* x, int64_t n, double* out) {
    if (ctx && !ctx->kids.empty()) return mctsb200_device_log(ctx->kids[0], x, n, out);
    if (!ctx || !x || !out || n < 0) return fail(MCTSB200_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return MCTSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevBuf in, o;
    CUDA_TRY(in.ensure(sizeof(double) * (size_t)n));
    CUDA_TRY(o.ensure(sizeof(double) * (size_t)n));
    CUDA_TRY(cudaMemcpyAsync(in.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    MCTSB200_LAUNCH(probe_log_kernel, (unsigned)((n + 255) / 256), 256, 0, ctx->stream, (const double*)in.p, (long long)n, (double*)o.p);
    CUDA_TRY(cudaMemcpyAsync(out, o.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    in.release();
    o.release();
    return MCTSB200_OK;
}

int32_t mctsb200_device_cos(mctsb200_ctx* ctx, const double* x, int64_t n, double* out) {
    if (ctx && !ctx->kids.empty()) return mctsb200_device_cos(ctx->kids[0], x, n, out);
    if (!ctx || !x || !out || n < 0) return fail(MCTSB200_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return MCTSB200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevBuf in, o;
    CUDA_TRY(in.ensure(sizeof(double) * (size_t)n));
    CUDA_TRY(o.ensure(sizeof(double) * (size_t)n));
    CUDA_TRY(cudaMemcpyAsync(in.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    MCTSB200_LAUNCH(probe_cos_kernel, (unsigned)((n + 255) / 256), 256, 0, ctx->stream, (const double*)in.p, (long long)n, (double*)o.p);
    CUDA_TRY(cudaMemcpyAsync(out, o.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    in.release();
    o.release();
    return MCTSB200_OK;
}

int32_t mctsb200_device_philox(mctsb200_ctx* ctx, uint64_t seed, uint64_t tree, uint32_t iteration, uint32_t stream_id, int32_t n_blocks, uint32_t* out) {
    if (ctx && !ctx->kids.empty()) return mctsb200_device_philox(ctx->kids[0], seed, tree, iteration, stream_id, n_blocks, out);
    if (!ctx || !out || n_blocks < 1 || n_blocks > 4096) return fail(MCTSB200_ERR_INVALID_ARG, "bad arguments");
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevBuf o;
    CUDA_TRY(o.ensure(sizeof(uint32_t) * 4 * (size_t)n_blocks));
    MCTSB200_LAUNCH(probe_philox_kernel, 1, 32, 0, ctx->stream, seed, tree, iteration, stream_id, 4 * n_blocks, (uint32_t*)o.p);
    CUDA_TRY(cudaMemcpyAsync(out, o.p, sizeof(uint32_t) * 4 * (size_t)n_blocks, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    o.release();
    return MCTSB200_OK;
}

}  // extern "C"
