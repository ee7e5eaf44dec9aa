// search.cuh -- the search loop on the device: UCT (src/vanilla.jl:229-276, 292-307, 354-386),
// DPW (src/dpw.jl:48-56, 80-154, 179-199; src/dpw_types.jl:162-186) and leaf evaluation
// (src/domain_knowledge.jl:65-73 + the POMDPTools RolloutSimulator loop).  Paths are relative to
// /root/reference; the code is a from-scratch design, not a translation:
//
//   * one TILE of G lanes (G = 1,2,4,...,32; G = 32 with 32-thread blocks is "one CTA per tree")
//     owns one tree for the whole launch; every lane of the tile carries the same control state
//     (node, depth, state, Philox stream), so shared decisions cost no communication;
//   * selection: lane j scores child j, the tile arg-maxes with warp shuffles (ties -> lowest index,
//     exactly the reference's strict '>' scan);
//   * the reference's recursion is an explicit path stack; backup runs deepest-first after the
//     descent, which reproduces the stale-statistics behaviour of re-entered nodes;
//   * leaf rollouts: K = rollouts_per_leaf independent Philox streams, one lane per rollout, summed
//     in stream order (so K>1 is still bit-reproducible); K = 1 is the reference's estimator;
//   * no transcendental is evaluated per child: log(total_n) is one table load per level, and the
//     widening tests k*N^alpha are integer compares against host-built threshold tables.
#pragma once
#ifndef MCTSB200_HOST_EMU
#include <math_constants.h>
#endif

#include "common.cuh"
#include "models.cuh"

namespace mctsb200 {

#ifndef MCTSB200_BOX_SCAN_U
#define MCTSB200_BOX_SCAN_U 8  // children a lane has in flight in the scans of continuous-action trees
#endif
constexpr int kBlockThreads = 128;
// Two ideas that looked free and measured slower on B200 (config 4, profiles/r2e_*: 465 ms without either, 477 ms with the widening
// thresholds hoisted into registers, 520 ms with 1024-entry UCB table heads in shared memory -- the 16 KB per block come out of the
// L1 that serves the node tables at an 86 % hit rate); kept behind switches for the record.
#ifndef MCTSB200_TILE_TAB
#define MCTSB200_TILE_TAB 0
#endif
#ifndef MCTSB200_HOIST_NMIN
#define MCTSB200_HOIST_NMIN 0
#endif
constexpr int kTileTab = MCTSB200_TILE_TAB;  // entries of c*sqrt(log N) and 1/sqrt(n) a block of the tile kernel stages in shared memory (0 = none)

template <int G>
struct Tile {
    static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "tile width");
    unsigned mask;
    int lane;
    __device__ __forceinline__ Tile() {
        lane = threadIdx.x & (G - 1);
        if constexpr (G == 32) mask = 0xffffffffu;
        else mask = ((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
    }
    __device__ __forceinline__ void sync() const {
        if constexpr (G > 1) __syncwarp(mask);
    }
    __device__ __forceinline__ double bcast(double v, int src) const {
        if constexpr (G > 1) return __shfl_sync(mask, v, src, G);
        else return v;
    }
    __device__ __forceinline__ int min_int(int v) const {
        if constexpr (G > 1) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(mask, v, o, G));
        }
        return v;
    }
    __device__ __forceinline__ double max_double(double v) const {
        if constexpr (G > 1) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, G));
        }
        return v;
    }
    __device__ __forceinline__ double sum_double(double v) const {  // identical in every lane (xor butterfly)
        if constexpr (G > 1) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v = v + __shfl_xor_sync(mask, v, o, G);
        }
        return v;
    }
    __device__ __forceinline__ bool any(bool p) const {
        if constexpr (G > 1) return __any_sync(mask, p) != 0;
        else return p;
    }
    // arg-max of (score, idx) carrying the winner's (q, n, t) along
    __device__ __forceinline__ void argmax_payload(double& score, int& idx, double& q, int& n, double& t) const {
        if constexpr (G > 1) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                const double os = __shfl_xor_sync(mask, score, o, G);
                const int oi = __shfl_xor_sync(mask, idx, o, G);
                const double oq = __shfl_xor_sync(mask, q, o, G);
                const int on = __shfl_xor_sync(mask, n, o, G);
                const double ot = __shfl_xor_sync(mask, t, o, G);
                if (os > score || (os == score && oi < idx)) {
                    score = os;
                    idx = oi;
                    q = oq;
                    n = on;
                    t = ot;
                }
            }
        }
    }
    // arg-max of (score, idx): larger score wins, equal scores -> smaller idx
    __device__ __forceinline__ void argmax(double& score, int& idx) const {
        if constexpr (G > 1) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                const double os = __shfl_xor_sync(mask, score, o, G);
                const int oi = __shfl_xor_sync(mask, idx, o, G);
                if (os > score || (os == score && oi < idx)) {
                    score = os;
                    idx = oi;
                }
            }
        }
    }
};

// x / n for a visit count n >= 1. An exactly zero numerator (every backup level inside a zero-reward region) would
// send CUDA's IEEE division to its ~150-instruction slow path; 0 / n is the numerator itself, sign included.
__device__ __forceinline__ double div_by_count(double x, double n) { return x == 0.0 ? x : DM_DIV(x, n); }
// the same without a branch (two of these interleave in the backup below): the division always runs, on 1.0 when x is zero
__device__ __forceinline__ double div_by_count_nb(double x, double n) {
    const double d = DM_DIV(x == 0.0 ? 1.0 : x, n);
    return x == 0.0 ? x : d;
}

// A whole POD of the node tables into registers with 256-bit loads that are issued HERE: written as plain member reads, the
// compiler sinks each load below the branches that precede its first use, and a tree level then pays one L2 round trip after
// another (state label -> terminal test -> transition bookkeeping -> widening test -> children statistics: three dependent
// trips per level in the SASS of the first version) instead of one.
template <class T>
__device__ __forceinline__ void load_now(const T* p, T& out) {
    static_assert(sizeof(T) % 32 == 0 && alignof(T) >= 32, "node-table PODs are whole 32-byte sectors");
#ifdef MCTSB200_HOST_EMU
    out = *p;
#else
    constexpr int kChunks = (int)(sizeof(T) / 32);
    unsigned long long w[4 * kChunks];
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
#ifdef __CUDACC_RTC__  // (the PTX version NVRTC emits for run-time plugins has no 256-bit vector loads: two 128-bit ones)
        asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(w[4 * i]), "=l"(w[4 * i + 1]) : "l"(reinterpret_cast<const char*>(p) + 32 * i) : "memory");
        asm volatile("ld.global.v2.u64 {%0,%1}, [%2];"
                     : "=l"(w[4 * i + 2]), "=l"(w[4 * i + 3])
                     : "l"(reinterpret_cast<const char*>(p) + 32 * i + 16)
                     : "memory");
#else
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(w[4 * i]), "=l"(w[4 * i + 1]), "=l"(w[4 * i + 2]), "=l"(w[4 * i + 3])
                     : "l"(reinterpret_cast<const char*>(p) + 32 * i)
                     : "memory");
#endif
    }
    memcpy(&out, w, sizeof(T));
#endif
}

// small loads that are issued where they are written (the children scans of continuous-action trees request several children's
// indices, then their nodes, before looking at any of them: four L2 round trips overlap instead of queueing)
__device__ __forceinline__ int ld_int_now(const int* p) {
#ifdef MCTSB200_HOST_EMU
    return *p;
#else
    int v;
    asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void ld_2u64_now(const void* p, unsigned long long& lo, unsigned long long& hi) {
#ifdef MCTSB200_HOST_EMU
    lo = reinterpret_cast<const unsigned long long*>(p)[0];
    hi = reinterpret_cast<const unsigned long long*>(p)[1];
#else
    asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
#endif
}

// multiplicity of a transition record += 1, fire-and-forget (RED: the warp does not wait for the old value)
__device__ __forceinline__ void bump_count(int* count) {
#ifdef MCTSB200_HOST_EMU
    *count += 1;
#else
    atomicAdd(count, 1);
#endif
}

template <class M, int G>
struct Search {
    using State = typename M::State;
    using Key = typename M::Key;
    static constexpr int A = M::kActions;

    const KernelArgs<M>& a;
    Tile<G> tile;
    // this tile's tables
    Row<M>* rows;
    Aux<M>* aux;
    TransRec* trans;
    PathEntry* path;
    uint16_t* dmap;
    HashSlot* hmap;
    uint64_t tree_index;
    // header mirror (registers, tile-uniform)
    int n_state, n_action, n_trans, n_plog, status;
    int* plog;  // this tree's push-log arena
    BoxANode* anodes;  // this tree's action nodes (continuous action spaces)
    // per-lane work counters (per-tree counts fit 32 bits: n_iterations*depth <= 2e9 is validated on the host;
    // rollout steps can exceed that and are kept in 64 bits)
    unsigned cnt[CNT_COUNT];
    unsigned long long rollout_steps;
    int path_stride;  // entries between consecutive levels of this tile's path
    // the heads of the two UCB tables in shared memory (c folded into the first: the same single rounded multiply), and the
    // first widening thresholds in registers: a level then waits for no table load after its row has arrived
    const double* s_cs;
    const double* s_rs;
    int nmin_s[4];      // nmin_state[0..3]
    int nmin_a[A + 1];  // nmin_action[0..A]
#ifdef MCTSB200_PHASE_TIMING
    long long cyc[4] = {0, 0, 0, 0};  // descent, leaf evaluation, backup, trips
#define PT_CLOCK(var) const long long var = clock64()
#define PT_ADD(i, v) cyc[i] += (v)
#else
#define PT_CLOCK(var)
#define PT_ADD(i, v)
#endif

    __device__ __forceinline__ Search(const KernelArgs<M>& args, long long tree) : a(args) {
        rows = args.rows + tree * args.geo.NS;
        aux = args.aux ? args.aux + tree * args.geo.NS : nullptr;
        trans = args.trans ? args.trans + tree * args.geo.NT : nullptr;
        plog = args.plog ? args.plog + tree * args.geo.NP : nullptr;
        anodes = args.anodes ? args.anodes + tree * args.geo.NA : nullptr;
        n_plog = 0;
        path = args.path + tree * args.geo.depth;
        path_stride = 1;
        dmap = M::kDense ? (uint16_t*)args.map + tree * args.geo.MAP : nullptr;
        hmap = M::kDense ? nullptr : (HashSlot*)args.map + tree * args.geo.MAP;
        tree_index = (uint64_t)(args.tree_index_base + tree);
#pragma unroll
        for (int i = 0; i < CNT_COUNT; ++i) cnt[i] = 0;
        rollout_steps = 0;
        s_cs = s_rs = nullptr;
#if MCTSB200_HOIST_NMIN
#pragma unroll
        for (int i = 0; i < 4; ++i) nmin_s[i] = (args.cfg.nmin_state && i < args.cfg.nmin_state_n) ? __ldg(args.cfg.nmin_state + i) : 0x7fffffff;
#pragma unroll
        for (int i = 0; i <= A; ++i) nmin_a[i] = args.cfg.nmin_action[i];
#endif
    }

    __device__ __forceinline__ bool lead() const { return tile.lane == 0; }

    // exact int -> double for 0 <= n < 2^31 on the FP64 pipe (I2F.F64 would go through the 16-lane XU pipe)
    __device__ __forceinline__ static double i2d(int n) { return DM_SUB(__hiloint2double(0x43300000, n), 4503599627370496.0); }

    // element `slot` of a small per-child array held in registers (a dynamic index would move the array to local memory)
    template <class T2>
    __device__ __forceinline__ static T2 pick_a(const T2 (&v)[M::kActions], int slot) {
        T2 out = v[0];
#pragma unroll
        for (int k = 1; k < M::kActions; ++k)
            if (slot == k) out = v[k];
        return out;
    }

    template <int N2>
    __device__ __forceinline__ static int pick_n(const int (&v)[N2], int i) {
        int out = v[0];
#pragma unroll
        for (int k = 1; k < N2; ++k)
            if (i == k) out = v[k];
        return out;
    }

    // ---- numeric hooks (src/domain_knowledge.jl:10,82,91 + table stand-ins) ----
    __device__ __forceinline__ double init_Q(const State& s, int act) const {
        if (M::kDense && a.cfg.init_q_table) return __ldg(a.cfg.init_q_table + M::dense_index(a.mdp, s) * A + act);
        return a.cfg.init_Q;
    }
    __device__ __forceinline__ int init_N(const State& s, int act) const {
        if (M::kDense && a.cfg.init_n_table) return __ldg(a.cfg.init_n_table + M::dense_index(a.mdp, s) * A + act);
        return a.cfg.init_N;
    }
    __device__ __forceinline__ double log_n(int N) const {
        if (N < a.cfg.log_tab_n) return __ldg(a.cfg.log_tab + N);
        return det_log((double)N);
    }

    // ---- state -> node map -----------------------------------------------------------------
    __device__ __forceinline__ int map_lookup(const State& s) const {
        if (M::kDense) {
            return (int)dmap[M::dense_index(a.mdp, s)] - 1;
        } else {
            const uint32_t tag = hash_of(s) | 1u;
            const int msk = a.geo.MAP - 1;
            const Key key = M::key_of(a.mdp, s);
            for (int i = (int)(tag >> 1) & msk;; i = (i + 1) & msk) {
                const HashSlot sl = hmap[i];
                if (sl.tag == 0) return -1;
                if (sl.tag == tag && M::key_equal(rows[sl.node].key, key)) return sl.node;
            }
        }
    }
    __device__ __forceinline__ void map_insert(const State& s, int node) const {  // leader only
        if (M::kDense) {
            dmap[M::dense_index(a.mdp, s)] = (uint16_t)(node + 1);
        } else {
            // s_lookup[s] = snode: a state inserted again (unmerged states with keep_tree, src/dpw.jl:127) maps to its newest node
            const uint32_t tag = hash_of(s) | 1u;
            const int msk = a.geo.MAP - 1;
            const Key key = M::key_of(a.mdp, s);
            int i = (int)(tag >> 1) & msk;
            while (hmap[i].tag != 0 && !(hmap[i].tag == tag && hmap[i].node != node && M::key_equal(rows[hmap[i].node].key, key))) i = (i + 1) & msk;
            hmap[i] = HashSlot{tag, node};
        }
    }
    __device__ __forceinline__ static uint32_t hash_of(const State& s) {
        if constexpr (M::kDense) return 0u;
        else return M::hash(s);
    }

    // ---- leaf evaluation ---------------------------------------------------------------------
    // POMDPTools simulate(::RolloutSimulator, mdp, policy, s0)
    __device__ __forceinline__ double rollout(State s, int max_steps, Stream& rs, unsigned& steps) const {
        const double eps = a.cfg.rollout_eps;
        const double gamma = M::discount(a.mdp);
        double disc = 1.0, r_total = 0.0;
        int step = 1;
        while (disc > eps && !M::is_terminal(a.mdp, s) && step <= max_steps) {
            State sp;
            double r;
            if constexpr (M::kActionDoubles > 0) {  // RandomPolicy on a continuous action space: rand(rng, actions(mdp, s))
                typename M::Action act;
                M::policy_action(a.mdp, a.cfg.policy, a.cfg.pparam, s, rs, act);
                // (these models run one rollout per leaf: every lane of the tile is here, with the same stream)
                if constexpr (G >= 2 && M::kCoopGen) M::gen_coop(tile, a.mdp, s, act, rs, sp, r);
                else M::gen(a.mdp, s, act, rs, sp, r);
            } else {
                int act;
                if (M::kDense && a.cfg.policy == MCTSB200_POLICY_TABLE) act = (int)a.cfg.policy_table[M::dense_index(a.mdp, s)];  // FunctionPolicy(f), tabulated
                else act = M::policy_action(a.mdp, a.cfg.policy, a.cfg.pparam, s, rs);
                M::gen(a.mdp, s, act, rs, sp, r);
            }
            r_total = DM_ADD(r_total, DM_MUL(disc, r));
            s = sp;
            disc = DM_MUL(disc, gamma);
            ++step;
            ++steps;
        }
        return r_total;
    }

    // estimate_value(estimator, mdp, s, remaining_depth)
    __device__ __forceinline__ double estimate(const State& s, int remaining_depth, uint32_t it) {
        if (a.cfg.estimate_kind == MCTSB200_EST_CONSTANT) return a.cfg.estimate_const;
        if (a.cfg.estimate_kind == MCTSB200_EST_TABLE) {
            if constexpr (M::kDense) return __ldg(a.cfg.value_table + M::dense_index(a.mdp, s));
            else return 0.0;
        }
        const int max_steps = a.cfg.rollout_max_depth == -1 ? remaining_depth : a.cfg.rollout_max_depth;
        const int K = a.cfg.K;
        if (K == 1) {  // every lane runs the same stream: no divergence, no shuffle
            Stream rs;
            rs.init(a.cfg.seed, tree_index, it, 1u);
            unsigned steps = 0;
            const double v = rollout(s, max_steps, rs, steps);
            if (lead()) {
                cnt[CNT_ROLLOUTS] += 1;
                rollout_steps += steps;
            }
            return v;
        }
        // leaf-parallel: lane l runs rollouts l, l+G, ...; summed in stream order 0..K-1
        double sum = 0.0;
        for (int base = 0; base < K; base += G) {
            const int k = base + tile.lane;
            double mine = 0.0;
            if (k < K) {
                Stream rs;
                rs.init(a.cfg.seed, tree_index, it, 1u + (uint32_t)k);
                unsigned steps = 0;
                mine = rollout(s, max_steps, rs, steps);
                cnt[CNT_ROLLOUTS] += 1;
                rollout_steps += steps;
            }
            const int m = min(G, K - base);
            for (int j = 0; j < m; ++j) sum = DM_ADD(sum, tile.bcast(mine, j));
        }
        return DM_DIV(sum, (double)K);
    }

    // ---- backup (src/vanilla.jl:272-274, src/dpw.jl:149-151), deepest level first ---------------
    // Two passes. What chains the levels together is only q_lvl = r_lvl + gamma * q_(lvl+1): two FP64 operations per level, run
    // first, leaving q_lvl in the path entry. The statistics updates (n += 1, total_n += 1, q += (q_lvl - q) / n: three loads,
    // a division, three stores) then depend on one another only where the same state node occurs twice on the path, so they run
    // two levels at a time with their loads and divisions interleaved, and one after the other where two neighbouring levels
    // share their node (self-loops and cycles: the deeper level's update must land first, as in the reference's unwinding
    // recursion).
    __device__ __forceinline__ void update_stats(Row<M>* rw, int slot, double qv) {
        // all three reads before the first write: a load that follows a store to the same line waits for L2
        const int nn = rw->n[slot] + 1;
        const int tn = rw->total_n + 1;
        const double qo = rw->q[slot];
        rw->n[slot] = nn;
        rw->total_n = tn;
        rw->q[slot] = DM_ADD(qo, div_by_count(DM_SUB(qv, qo), i2d(nn)));
    }
    __device__ __forceinline__ void backup(int top, double v) {
        tile.sync();
        if (lead()) {
            const double gamma = M::discount(a.mdp);
            for (int lvl = top - 1; lvl >= 0; --lvl) {
                PathEntry* e = path + lvl * path_stride;
                v = DM_ADD(e->r, DM_MUL(gamma, v));
                e->r = v;
            }
            int lvl = top - 1;
            for (; lvl >= 1; lvl -= 2) {
                const PathEntry e0 = path[lvl * path_stride], e1 = path[(lvl - 1) * path_stride];
                Row<M>* r0 = rows + e0.node;
                Row<M>* r1 = rows + e1.node;
                if (e0.node != e1.node) {
                    const int nn0 = r0->n[e0.slot] + 1, tn0 = r0->total_n + 1;
                    const int nn1 = r1->n[e1.slot] + 1, tn1 = r1->total_n + 1;
                    const double qo0 = r0->q[e0.slot], qo1 = r1->q[e1.slot];
                    const double d0 = div_by_count_nb(DM_SUB(e0.r, qo0), i2d(nn0));
                    const double d1 = div_by_count_nb(DM_SUB(e1.r, qo1), i2d(nn1));
                    r0->n[e0.slot] = nn0;
                    r0->total_n = tn0;
                    r0->q[e0.slot] = DM_ADD(qo0, d0);
                    r1->n[e1.slot] = nn1;
                    r1->total_n = tn1;
                    r1->q[e1.slot] = DM_ADD(qo1, d1);
                } else {
                    update_stats(r0, e0.slot, e0.r);
                    update_stats(r1, e1.slot, e1.r);
                }
            }
            if (lvl == 0) {
                const PathEntry e0 = path[0];
                update_stats(rows + e0.node, e0.slot, e0.r);
            }
        }
        tile.sync();
    }

    // =====================================================================================
    // Vanilla UCT
    // =====================================================================================
    // insert_node! (src/vanilla.jl:292-307); returns the new 0-based node id or -1 on overflow
    __device__ __forceinline__ int uct_insert(const State& s) {
        const int id = n_state;
        if (id >= a.geo.NS) {
            status = MCTSB200_ERR_CAPACITY;
            return -1;
        }
        if (lead()) {
            Row<M>* rw = rows + id;
            int tot = 0;
#pragma unroll
            for (int k = 0; k < A; ++k) {
                const int n0 = init_N(s, k);
                tot += n0;
                rw->n[k] = n0;
                rw->q[k] = init_Q(s, k);
            }
            rw->total_n = tot;
            {
                uint32_t pk = 0;
#pragma unroll
                for (int k = 0; k < A; ++k) pk = Row<M>::pack_add(pk, k, k);
                rw->packed = pk;
            }
            rw->key = M::key_of(a.mdp, s);
            map_insert(s, id);
            cnt[CNT_STATE_INSERTS] += 1;
            cnt[CNT_ACTION_INSERTS] += A;
            if (kSuccCache && aux) {
#pragma unroll
                for (int k = 0; k < A; ++k) aux[id].sp1[k] = -1;  // successors not looked up yet
            }
        }
        n_state = id + 1;
        n_action += A;
        return id;
    }

    // Vanilla UCT on a hashed-state model whose (s, a) has a single successor (MountainCar, plugins registered with
    // max_branch = 1; their gen draws no random numbers): the node a child leads to is remembered in Aux::sp1 / r1 the first time it
    // is looked up, and later visits follow it without the transition arithmetic and the hash probe.
    static constexpr bool kSuccCache = M::kMaxBranch == 1 && !M::kDense;

    // ---- selection ---------------------------------------------------------------------------
    static constexpr int T = (A + G - 1) / G;  // children per lane
    static constexpr int kNone = 0x7fffffff;
    struct Kids {  // this lane's children: slot j = lane + t*G
        int n[T];
        double q[T];
    };

    // one row = three vector loads when a lane owns all four children
    __device__ __forceinline__ void load_kids(const Row<M>* rw, int nchild, Kids& k, int& total_n) const {
        if constexpr (G == 1 && A == 4) {
            const double2 q01 = *reinterpret_cast<const double2*>(&rw->q[0]);
            const double2 q23 = *reinterpret_cast<const double2*>(&rw->q[2]);
            const int4 n4 = *reinterpret_cast<const int4*>(&rw->n[0]);
            k.q[0] = q01.x; k.q[1] = q01.y; k.q[2] = q23.x; k.q[3] = q23.y;
            k.n[0] = n4.x; k.n[1] = n4.y; k.n[2] = n4.z; k.n[3] = n4.w;
#pragma unroll
            for (int t = 0; t < T; ++t)
                if (t >= nchild) {
                    k.n[t] = 1;
                    k.q[t] = 0.0;
                }
        } else {
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int j = tile.lane + t * G;
                k.n[t] = j < nchild ? rw->n[j] : 1;
                k.q[t] = j < nchild ? rw->q[j] : 0.0;
            }
        }
        total_n = rw->total_n;
    }

    // the same from a row that is already in registers (load_now)
    __device__ __forceinline__ void kids_of(const Row<M>& rw, int nchild, Kids& k, int& total_n) const {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = tile.lane + t * G;
            k.n[t] = j < nchild ? pick_a(rw.n, j) : 1;
            k.q[t] = j < nchild ? pick_a(rw.q, j) : 0.0;
        }
        total_n = rw.total_n;
    }

    // argmax(q, children): strict '>' scan from -Inf, default first child
    __device__ __forceinline__ int argmax_q(const Kids& k, int nchild) const {
        double best = -CUDART_INF;
        int best_i = kNone;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = tile.lane + t * G;
            if (j < nchild && k.q[t] > best) {
                best = k.q[t];
                best_i = j;
            }
        }
        tile.argmax(best, best_i);
        return best_i == kNone ? 0 : best_i;
    }

    // The reference's score, operation for operation: q + c*sqrt(log(N)/n). Static (no access to `this`) so
    // that the out-of-line near-tie handler below does not force the per-tile state into local memory.
    __device__ __forceinline__ static int exact_scores(const Tile<G>& tile, const DevCfg& cfg, const Kids& k, int nchild, int N, bool dpw) {
        const double c = cfg.c;
        const double lN = N < cfg.log_tab_n ? __ldg(cfg.log_tab + N) : det_log((double)N);
        double best = -CUDART_INF;
        int best_i = kNone;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = tile.lane + t * G;
            if (j < nchild) {
                double ucb;
                if (dpw && lN <= 0.0 && k.n[t] == 0) ucb = k.q[t];
                else ucb = DM_ADD(k.q[t], DM_MUL(c, DM_SQRT(DM_DIV(lN, i2d(k.n[t])))));
                if (ucb > best) {
                    best = ucb;
                    best_i = j;
                }
            }
        }
        tile.argmax(best, best_i);
        return best_i == kNone ? 0 : best_i;
    }
    __device__ __forceinline__ int select_exact(const Kids& k, int nchild, int N, bool dpw) {
        if (lead()) cnt[CNT_SELECT_EXACT] += 1;
        return exact_scores(tile, a.cfg, k, nchild, N, dpw);
    }

    // A near-tie of the fast-path scores: children inside the margin that carry bit-identical (n, q) have
    // bit-identical reference scores (tie -> lowest index, as the strict '>' scan); anything else is decided
    // with the reference's exact operation sequence. Returns slot | 256 if the exact scores were needed.
    __device__ __noinline__ static int select_near_tie(Tile<G> tile, const DevCfg* cfg, Kids k, int nchild, int N, bool dpw, double cs,
                                                       double margin) {
        double S[T];
        double best = -CUDART_INF, bq = 0.0, bt = 0.0;
        int best_i = kNone, bn = 0;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = tile.lane + t * G;
            const double term = DM_MUL(cs, __ldg(cfg->rsqrt_tab + k.n[t]));
            S[t] = DM_ADD(k.q[t], term);
            if (j < nchild && S[t] > best) {
                best = S[t];
                best_i = j;
                bq = k.q[t];
                bn = k.n[t];
                bt = term;
            }
        }
        tile.argmax_payload(best, best_i, bq, bn, bt);
        bool need_exact = best_i == kNone;  // every score NaN / -Inf
        int tie_i = best_i;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = tile.lane + t * G;
            if (j < nchild && j != best_i && !(DM_SUB(best, S[t]) > margin)) {  // also catches NaN / Inf - Inf
                if (k.n[t] == bn && k.q[t] == bq) tie_i = min(tie_i, j);
                else need_exact = true;
            }
        }
        if (tile.any(need_exact)) return exact_scores(tile, *cfg, k, nchild, N, dpw) | 256;
        return tile.min_int(tie_i);
    }

    // best_sanode_UCB for both solvers (src/vanilla.jl:354-386, src/dpw.jl:179-199); returns the child slot.
    //
    // Fast path: S_j = q_j + (c*sqrt(log N)) * n_j^(-1/2) from two host-built tables -- 1 load, 1 mul, 1 add per
    // child, no division, square root or XU-pipe instruction. S_j differs from the reference's rounded score by
    // at most ~12 ulp of (|q_j| + term_j); the arg-max is accepted only if the runner-up is further away than
    // 2^-44 * sum_j(|q_j| + term_j) (a >= 32x safety factor); otherwise select_near_tie decides.
    __device__ __forceinline__ int select_child(const Row<M>* rw, int nchild, bool dpw) {
        Kids k;
        int N;
        load_kids(rw, nchild, k, N);
        return select_among(k, N, nchild, dpw);
    }
    __device__ __forceinline__ int select_child(const Row<M>& rw, int nchild, bool dpw) {
        Kids k;
        int N;
        kids_of(rw, nchild, k, N);
        return select_among(k, N, nchild, dpw);
    }
    __device__ __forceinline__ int select_among(const Kids& k, int N, int nchild, bool dpw) {
        const double c = a.cfg.c;
        if (c == 0.0) return argmax_q(k, nchild);
        int first_zero = kNone;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = tile.lane + t * G;
            if (j < nchild && k.n[t] == 0) first_zero = min(first_zero, j);
        }
        first_zero = tile.min_int(first_zero);
        if (first_zero != kNone) {
            if (!dpw) return first_zero;  // vanilla: "if action was not used, use it"
            return select_exact(k, nchild, N, true);
        }
        if (dpw && N <= 1) return argmax_q(k, nchild);  // log(N) <= 0 with every n > 0: the exploration term is exactly 0
        // (total_n is the sum of its children's n -- both start from init_N and grow together -- so N bounds every index below)
        if (N >= a.cfg.log_tab_n) return select_exact(k, nchild, N, dpw);

        const double cs = (s_cs && N < kTileTab) ? s_cs[N] : DM_MUL(c, __ldg(a.cfg.sqrtlog_tab + N));
        double best = -CUDART_INF, second = -CUDART_INF, mag = 0.0;
        int best_i = kNone;
        // the margin: 2^-44 * sum(|q_j| + term_j) <= 2^-42 * (Q bound + c sqrt(log N)) for up to four children (1/sqrt(n) <= 1), known
        // before the children's terms are; without a bound of |Q| it is summed from the children
#ifndef MCTSB200_TILE_QBOUND
#define MCTSB200_TILE_QBOUND 0  // (measured on B200, profiles/r2k_*: the bound-based margin is 2 % SLOWER on config 4 than summing it)
#endif
        const bool have_bound = MCTSB200_TILE_QBOUND && A <= 4 && a.cfg.q_bound_tile == a.cfg.q_bound_tile;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = tile.lane + t * G;
            if (j < nchild) {
                const double term = DM_MUL(cs, (s_rs && k.n[t] < kTileTab) ? s_rs[k.n[t]] : __ldg(a.cfg.rsqrt_tab + k.n[t]));
                const double S = DM_ADD(k.q[t], term);
                if (!have_bound) mag = DM_ADD(mag, DM_ADD(fabs(k.q[t]), term));
                if (S > best) {
                    second = best;
                    best = S;
                    best_i = j;
                } else if (S > second) {
                    second = S;
                }
            }
        }
        if constexpr (G > 1) {
            const double lane_best = best;
            tile.argmax(best, best_i);
            const bool mine = best_i != kNone && (best_i % G) == tile.lane;  // the winner is one of this lane's children
            second = tile.max_double(mine ? second : lane_best);
            if (!have_bound) mag = tile.sum_double(mag);
        }
        const double margin = have_bound ? DM_MUL(2.2737367544323206e-13, DM_ADD(a.cfg.q_bound_tile, fabs(cs)))
                                         : DM_MUL(5.6843418860808015e-14, mag);  // (NaN if any input is: the near-tie path decides)
        if (best_i != kNone && DM_SUB(best, second) > margin) return best_i;
        const int r = select_near_tie(tile, &a.cfg, k, nchild, N, dpw, cs, margin);
        if ((r & 256) && lead()) cnt[CNT_SELECT_EXACT] += 1;
        return r & 255;
    }

    // one iteration of src/vanilla.jl:229-235: simulate(root, depth) unrolled into three warp-synchronous phases:
    // descent (one tree level per trip), ONE leaf-evaluation phase for all lanes together (lanes reach their
    // leaves at different levels; evaluating inside the descent loop would run a separate, mostly empty rollout
    // loop per level), and the deepest-first backup.
    __device__ __forceinline__ void uct_iteration(int root, uint32_t it) {
        Stream rng;
        rng.init(a.cfg.seed, tree_index, it, 0u);
        int node = root, d = a.cfg.depth, top = 0;
        State s;
        int leaf = 0;  // 0: terminal (value 0.0), 1: estimate_value(s, d), 2: table overflow
        PT_CLOCK(t0);
        const bool cache = kSuccCache && aux != nullptr;
        for (;;) {
            PT_ADD(3, 1);
            // the node's row (state label + children statistics) and, for single-successor hashed models, its successor cache:
            // one batch of loads before the first branch (load_now)
            Row<M> R;
            Aux<M> X;
            load_now(rows + node, R);
            if (cache) load_now(aux + node, X);
            s = M::state_of(a.mdp, R.key);
            if (M::is_terminal(a.mdp, s)) break;
            if (d == 0) {
                leaf = 1;
                break;
            }
            if (lead()) cnt[CNT_LEVELS] += 1;
            const int slot = select_child(R, A, false);
            State sp;
            double r;
            int next;
            const int known = cache ? pick_a(X.sp1, slot) : -1;
            if (known >= 0) {
                next = known;
                r = pick_a(X.r1, slot);
            } else {
                M::gen(a.mdp, s, slot, rng, sp, r);
                next = map_lookup(sp);
            }
            if (lead()) path[top * path_stride] = PathEntry{node, slot, r};
            ++top;
            --d;
            if (next < 0) {
                s = sp;
                next = uct_insert(sp);  // new node: estimate_value(sp, depth-1), also when sp is terminal
                leaf = next < 0 ? 2 : 1;
            }
            if (cache && known < 0 && next >= 0 && lead()) {
                aux[node].sp1[slot] = next;
                aux[node].r1[slot] = r;
            }
            if (a.vis && next >= 0 && lead()) {  // record_visit!(tree, said, spid) (src/vanilla.jl:268-270, 328-334)
                VisRec* v = a.vis + ((long long)(tree_index - (uint64_t)a.tree_index_base) * a.geo.NS + node) * (A * a.vis_branch) + slot * a.vis_branch;
                for (int b = 0; b < a.vis_branch; ++b) {
                    if (v[b].count == 0) {
                        v[b] = VisRec{next, 1};
                        break;
                    }
                    if (v[b].sp == next) {
                        v[b].count += 1;
                        break;
                    }
                }
            }
            if (leaf) break;
            node = next;
        }
        if (leaf == 2) return;  // capacity error: abandon this simulation
        tile.sync();
        PT_CLOCK(t1);
        const double v = leaf == 1 ? estimate(s, d, it) : 0.0;
        tile.sync();
        PT_CLOCK(t2);
        backup(top, v);
        PT_CLOCK(t3);
        PT_ADD(0, t1 - t0);
        PT_ADD(1, t2 - t1);
        PT_ADD(2, t3 - t2);
        if (lead()) cnt[CNT_SIMULATIONS] += 1;
    }

    // =====================================================================================
    // DPW
    // =====================================================================================
    // insert_state_node! (src/dpw_types.jl:162-171)
    __device__ __forceinline__ int dpw_insert_state(const State& s, bool maintain_lookup) {
        const int id = n_state;
        if (id >= a.geo.NS) {
            status = MCTSB200_ERR_CAPACITY;
            return -1;
        }
        if (lead()) {
            Row<M>* rw = rows + id;
            rw->total_n = 0;
            rw->packed = 0;
            rw->key = M::key_of(a.mdp, s);
            if (maintain_lookup) map_insert(s, id);
            cnt[CNT_STATE_INSERTS] += 1;
            if constexpr (M::kActionDoubles > 0) {  // the (empty) children list of a continuous-action state node
                aux[id].nac[0] = 0;
                aux[id].pbase[0] = 0;
                aux[id].pcap[0] = 0;
            }
        }
        n_state = id + 1;
        return id;
    }

    // insert_action_node! (src/dpw_types.jl:174-186) into the next free slot of the row. R / X are the level's register copies
    // of the node's row and transition bookkeeping: every lane updates them, the leader writes the same values through.
    __device__ __forceinline__ void dpw_insert_action(int node, int slot, const State& s, int act, Row<M>& R, Aux<M>& X) {
        const int n0 = init_N(s, act);
        const double q0 = init_Q(s, act);
        R.packed = Row<M>::pack_add(R.packed, slot, act);
        R.total_n += n0;  // tree.total_n[snode] += n0
#pragma unroll
        for (int k = 0; k < A; ++k)
            if (k == slot) {
                R.n[k] = n0;
                R.q[k] = q0;
                X.nac[k] = 0;
                X.thead[k] = -1;
                X.tpush[k] = 0;
                X.pbase[k] = 0;
                X.pcap[k] = 0;
            }
        if (lead()) {
            Row<M>* rw = rows + node;
            Aux<M>* ax = aux + node;
            rw->n[slot] = n0;
            rw->q[slot] = q0;
            rw->packed = R.packed;
            rw->total_n = R.total_n;
            ax->nac[slot] = 0;
            ax->thead[slot] = -1;
            ax->tpush[slot] = 0;
            ax->pbase[slot] = 0;
            ax->pcap[slot] = 0;
            ax->seq[slot] = n_action + 1;
            cnt[CNT_ACTION_INSERTS] += 1;
        }
        n_action += 1;
    }

    // element `slot` of a small array held in registers (a dynamic index would move the array to local memory)
    template <class T2>
    __device__ __forceinline__ static T2 pick(const T2 (&v)[A], int slot) {
        T2 out = v[0];
#pragma unroll
        for (int k = 1; k < A; ++k)
            if (slot == k) out = v[k];
        return out;
    }

    // unbounded branching (a run-time plugin registered with max_branch = 0): transitions[sanode] is kept as the reference keeps
    // it, a vector with one entry per push (Aux::pbase / pcap, the plog arena); otherwise as distinct records with multiplicities
    static constexpr bool kPushLog = M::kMaxBranch == 0;

    // one iteration of src/dpw.jl:48-56: simulate(dpw, snode, depth) as descent + backup
    __device__ __forceinline__ void dpw_iteration(int root, uint32_t it) {
        Stream rng;
        rng.init(a.cfg.seed, tree_index, it, 0u);
        int node = root, d = a.cfg.depth, top = 0;
        State s;
        int leaf = 0;  // 0: terminal (value 0.0), 1: estimate_value(s, d), 2: table overflow (see uct_iteration)
        PT_CLOCK(t0);
        for (;;) {
            PT_ADD(3, 1);
            // Everything this level needs from the node is requested at once, before the first branch: the row (children statistics
            // and the state label) and the per-child transition bookkeeping, so the level pays ONE trip to L2 / HBM instead of a
            // chain label -> bookkeeping -> statistics (load_now: 256-bit loads the compiler cannot sink).
            Row<M>* rw = rows + node;
            Aux<M>* ax = aux + node;
            Row<M> R;
            Aux<M> X;
            load_now(rw, R);
            load_now(ax, X);
            s = M::state_of(a.mdp, R.key);
            if (M::is_terminal(a.mdp, s)) break;
            if (d == 0) {
                leaf = 1;
                break;
            }
            int nchild = R.nchild();
            // ---- action progressive widening (src/dpw.jl:94-114) ----
            if (a.cfg.enable_action_pw) {
                // length(children) <= k_action * total_n^alpha_action  <=>  total_n >= nmin_action[len]
#if MCTSB200_HOIST_NMIN
                if (R.total_n >= pick_n(nmin_a, nchild)) {
#else
                if (R.total_n >= a.cfg.nmin_action[nchild]) {
#endif
                    // next_action(::RandomActionGenerator) always draws; a node that owns every action cannot gain one (check_repeat_action
                    // is always on here, see capi.cu), so the word is consumed unread
                    bool tabulated = false;
                    if constexpr (M::kDense) tabulated = a.cfg.next_action_table != nullptr;
                    if (nchild == A) {
                        if (!tabulated) rng.skip();
                    } else {
                        // next_action(f::Function, mdp, s, snode) = f(mdp, s, snode) (src/action_gen.jl:14), tabulated: no draw
                        int act;
                        if (tabulated) act = (int)a.cfg.next_action_table[M::dense_index(a.mdp, s)];
                        else act = rng.uniform_int(A);
                        bool present = false;
                        const uint32_t pk = R.packed;
#pragma unroll
                        for (int j = 0; j < A; ++j) present |= (j < nchild) & ((int)((pk >> (4 * j)) & 15u) == act);
                        if (!present) {
                            dpw_insert_action(node, nchild, s, act, R, X);
                            ++nchild;
                            tile.sync();
                        }
                    }
                }
            } else if (nchild == 0) {
                for (int act = 0; act < A; ++act) dpw_insert_action(node, act, s, act, R, X);
                nchild = A;
                tile.sync();
            }
            if (lead()) {
                cnt[CNT_LEVELS] += 1;
                cnt[CNT_CHILDREN_SEEN] += (unsigned)nchild;
            }
            const int slot = select_child(R, nchild, true);
            const int act = R.alabel(slot);

            // ---- state progressive widening (src/dpw.jl:119-141) ----
            bool new_node = false;
            int spnode;
            double r;
            State sp;
            const int nac = pick(X.nac, slot);
            bool widen = nac == 0;
            if (!widen && a.cfg.enable_state_pw) {
                // n_a_children <= k_state * n^alpha_state  <=>  n >= nmin_state[n_a_children]
#if MCTSB200_HOIST_NMIN
                const int need = nac < 4 ? pick_n(nmin_s, nac) : (nac < a.cfg.nmin_state_n ? __ldg(a.cfg.nmin_state + nac) : 0x7fffffff);
#else
                const int need = nac < a.cfg.nmin_state_n ? __ldg(a.cfg.nmin_state + nac) : 0x7fffffff;
#endif
                widen = pick(R.n, slot) >= need;
            }
            if (M::kMaxBranch == 1 && widen && nac == 1 && a.cfg.check_repeat_state) {
                // A model whose (s, a) has exactly one successor, with merged states: gen would return the state this action
                // node already leads to and the lookup would find its node -- push!(transitions[sanode], ...) on the one
                // record, without the transition arithmetic and the hash probe (MountainCar: gen draws no random number).
                spnode = pick(X.sp1, slot);
                r = pick(X.r1, slot);
                if (lead()) {
                    bump_count(&trans[pick(X.thead, slot)].count);
                    ax->tpush[slot] = pick(X.tpush, slot) + 1;
                }
            } else if (widen) {
                M::gen(a.mdp, s, act, rng, sp, r);
                spnode = a.cfg.check_repeat_state ? map_lookup(sp) : -1;
                if (spnode < 0) {
                    spnode = dpw_insert_state(sp, a.cfg.maintain_lookup != 0);
                    if (spnode < 0) {
                        leaf = 2;
                        break;
                    }
                    new_node = true;
                }
                // push!(transitions[sanode], (spnode, r)) in multiset form; unique_transitions bookkeeping
                int rec = -1;
                if (!new_node) {
                    for (rec = pick(X.thead, slot); rec >= 0; rec = trans[rec].next)
                        if (trans[rec].sp == spnode) break;
                }
                if (rec < 0 && n_trans >= a.geo.NT) {
                    status = MCTSB200_ERR_CAPACITY;
                    leaf = 2;
                    break;
                }
                if constexpr (kPushLog) {
                    // push!(transitions[sanode], ...) literally: append the record id to the action node's log
                    const int pushes = pick(X.tpush, slot), cap = pick(X.pcap, slot);
                    int base = pick(X.pbase, slot);
                    if (pushes == cap) {  // full (or not allocated yet): move to a chunk twice as large at the end of the arena
                        const int ncap = cap ? 2 * cap : 4;
                        if (n_plog + ncap > a.geo.NP) {
                            status = MCTSB200_ERR_CAPACITY;
                            leaf = 2;
                            break;
                        }
                        const int nbase = n_plog;
                        for (int i = tile.lane; i < pushes; i += G) plog[nbase + i] = plog[base + i];
                        if (lead()) {
                            ax->pbase[slot] = nbase;
                            ax->pcap[slot] = ncap;
                        }
                        n_plog += ncap;
                        base = nbase;
                    }
                    if (lead()) plog[base + pushes] = rec >= 0 ? rec : n_trans;
                }
                if (lead()) {
                    if (rec >= 0) {
                        bump_count(&trans[rec].count);
                    } else {
                        TransRec t;
                        t.r = r;
                        t.sp = spnode;
                        t.next = pick(X.thead, slot);
                        t.count = 1;
                        t.pad[0] = t.pad[1] = t.pad[2] = 0;
                        trans[n_trans] = t;
                        ax->thead[slot] = n_trans;
                        ax->nac[slot] = nac + 1;
                        ax->sp1[slot] = spnode;
                        ax->r1[slot] = r;
                    }
                    ax->tpush[slot] = pick(X.tpush, slot) + 1;
                }
                if (rec < 0) n_trans += 1;
            } else {
                // rand(rng, transitions[sanode]): index j of the push-ordered multiset; records are chained
                // newest-first, so walk with the mirrored index
                const int pushes = pick(X.tpush, slot);
                const int j = rng.uniform_int(pushes);
                if (M::kMaxBranch == 1 && a.cfg.check_repeat_state) {  // a single record: every draw picks it
                    spnode = pick(X.sp1, slot);
                    r = pick(X.r1, slot);
                } else if (kPushLog) {  // transitions[sanode][j]
                    const int rec = plog[pick(X.pbase, slot) + j];
                    spnode = trans[rec].sp;
                    r = trans[rec].r;
                } else {
                    int jm = pushes - 1 - j, cum = 0, rec = pick(X.thead, slot);
                    for (;;) {
                        const int cnt_r = trans[rec].count;
                        if (jm < cum + cnt_r || trans[rec].next < 0) break;
                        cum += cnt_r;
                        rec = trans[rec].next;
                    }
                    spnode = trans[rec].sp;
                    r = trans[rec].r;
                }
                if (lead()) cnt[CNT_TRANS_SAMPLES] += 1;
            }
            if (lead()) path[top * path_stride] = PathEntry{node, slot, r};
            ++top;
            tile.sync();
            node = spnode;
            --d;
            if (new_node) {
                s = sp;
                leaf = 1;  // estimate_value(sp, d-1)
                break;
            }
        }
        if (leaf == 2) return;
        PT_CLOCK(t1);
        const double v = leaf == 1 ? estimate(s, d, it) : 0.0;
        PT_CLOCK(t2);
        backup(top, v);
        PT_CLOCK(t3);
        PT_ADD(0, t1 - t0);
        PT_ADD(1, t2 - t1);
        PT_ADD(2, t3 - t2);
        if (lead()) cnt[CNT_SIMULATIONS] += 1;
    }

    // =====================================================================================
    // DPW over a continuous action space (models with kActionDoubles > 0; test/discussion_514.jl)
    // =====================================================================================
    // A growable int vector in the tree's arena -- the children of a state node, or transitions[sanode] -- receives `value` at
    // position `len`: a full (or still empty) vector moves to a chunk twice as large at the end of the arena. base / cap are
    // the caller's register copies; the leader writes the new location through g_base / g_cap. false = the arena is full.
    __device__ __forceinline__ bool vec_push(int& base, int& cap, int len, int value, int* g_base, int* g_cap) {
        if (len == cap) {
            const int ncap = cap ? 2 * cap : 4;
            if (n_plog + ncap > a.geo.NP) {
                status = MCTSB200_ERR_CAPACITY;
                return false;
            }
            const int nbase = n_plog;
            for (int i = tile.lane; i < len; i += G) plog[nbase + i] = plog[base + i];
            if (lead()) {
                *g_base = nbase;
                *g_cap = ncap;
            }
            n_plog += ncap;
            base = nbase;
            cap = ncap;
        }
        if (lead()) plog[base + len] = value;
        tile.sync();
        return true;
    }

    // one iteration of src/dpw.jl:48-56 when actions(mdp, s) is a continuum: a state node owns up to k_action * N^alpha_action
    // action nodes (BoxANode table, children listed in the arena), next_action draws a vector, and the lanes of the tile share
    // the scans over a node's children (repeat check, UCB arg-max) -- the place where a wide tile pays.
    __device__ __forceinline__ void dpw_box_iteration(int root, uint32_t it) {
        using Act = typename M::Action;
        constexpr int AD = M::kActionDoubles;
        BoxANode* const an = anodes;
        Stream rng;
        rng.init(a.cfg.seed, tree_index, it, 0u);
        int node = root, d = a.cfg.depth, top = 0;
        State s;
        int leaf = 0;  // 0: terminal (value 0.0), 1: estimate_value(s, d), 2: table overflow
        for (;;) {
            Row<M>* rw = rows + node;
            Aux<M>* ax = aux + node;
            Row<M> R;
            Aux<M> X;
            load_now(rw, R);
            load_now(ax, X);
            s = M::state_of(a.mdp, R.key);
            if (M::is_terminal(a.mdp, s)) break;
            if (d == 0) {
                leaf = 1;
                break;
            }
            int nchild = X.nac[0], cbase = X.pbase[0], ccap = X.pcap[0];
            int total_n = R.total_n;
            // ---- action progressive widening (src/dpw.jl:94-105); the host refuses enable_action_pw = false for these models ----
            // length(children) <= k_action * total_n^alpha_action  <=>  total_n >= nmin_action_tab[len]
            const int need_a = nchild < a.cfg.nmin_action_n ? __ldg(a.cfg.nmin_action_tab + nchild) : 0x7fffffff;
            if (total_n >= need_a) {
                Act act;
                M::next_action(a.mdp, s, rng, act);  // next_action(::RandomActionGenerator): always draws
                bool present = false;
                if (a.cfg.check_repeat_action) {  // haskey(tree.a_lookup, (snode, a)): isequal on the vector = bitwise
                    constexpr int U = MCTSB200_BOX_SCAN_U;  // children per lane and trip: their indices, then their labels, are requested together
                    for (int j0 = tile.lane; j0 < nchild; j0 += U * G) {
                        int id[U];
                        unsigned long long lab[U][4];
#pragma unroll
                        for (int u = 0; u < U; ++u) id[u] = ld_int_now(plog + cbase + (j0 + u * G < nchild ? j0 + u * G : j0));
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            ld_2u64_now(an[id[u]].a, lab[u][0], lab[u][1]);
                            if constexpr (AD > 2) ld_2u64_now(an[id[u]].a + 2, lab[u][2], lab[u][3]);
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            bool eq = j0 + u * G < nchild;
#pragma unroll
                            for (int i = 0; i < AD; ++i) eq = eq && lab[u][i] == (unsigned long long)__double_as_longlong(act.v[i]);
                            present = present || eq;
                        }
                    }
                    present = tile.any(present);
                }
                if (!present) {
                    if (n_action >= a.geo.NA) {
                        status = MCTSB200_ERR_CAPACITY;
                        leaf = 2;
                        break;
                    }
                    const int id = n_action;
                    const int n0 = a.cfg.init_N;
                    if (lead()) {  // insert_action_node! (src/dpw_types.jl:174-186)
                        BoxANode nn;
                        nn.q = a.cfg.init_Q;
                        nn.n = n0;
                        nn.nac = 0;
                        nn.thead = -1;
                        nn.tpush = 0;
                        nn.pbase = 0;
                        nn.pcap = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) nn.a[i] = i < AD ? act.v[i < AD ? i : 0] : 0.0;
                        an[id] = nn;
                        cnt[CNT_ACTION_INSERTS] += 1;
                    }
                    if (!vec_push(cbase, ccap, nchild, id, &ax->pbase[0], &ax->pcap[0])) {
                        leaf = 2;
                        break;
                    }
                    ++nchild;
                    total_n += n0;  // tree.total_n[snode] += n0
                    if (lead()) {
                        ax->nac[0] = nchild;
                        rw->total_n = total_n;
                    }
                    n_action += 1;
                    tile.sync();
                }
            }
            if (lead()) {
                cnt[CNT_LEVELS] += 1;
                cnt[CNT_CHILDREN_SEEN] += (unsigned)nchild;
            }
            // ---- best_sanode_UCB (src/dpw.jl:179-199), the reference's expression operation for operation; lane j scores children
            // j, j + G, ... in order, the tile keeps the first maximum ----
            const double c = a.cfg.c;
            const double ltn = log_n(total_n);
            double best = -CUDART_INF;
            int best_j = kNone;
            {
                constexpr int U = MCTSB200_BOX_SCAN_U;  // as in the repeat check: several children's (q, n) in flight per lane
                for (int j0 = tile.lane; j0 < nchild; j0 += U * G) {
                    int id[U];
                    unsigned long long qb[U], nb[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) id[u] = ld_int_now(plog + cbase + (j0 + u * G < nchild ? j0 + u * G : j0));
#pragma unroll
                    for (int u = 0; u < U; ++u) ld_2u64_now(an + id[u], qb[u], nb[u]);  // (q, n | nac)
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int j = j0 + u * G;
                        if (j < nchild) {
                            const double q = __longlong_as_double((long long)qb[u]);
                            const int nn = (int)(unsigned)nb[u];
                            double ucb;
                            if ((ltn <= 0.0 && nn == 0) || c == 0.0) ucb = q;
                            else ucb = DM_ADD(q, DM_MUL(c, DM_SQRT(DM_DIV(ltn, i2d(nn)))));
                            if (ucb > best) {  // (children of a lane are taken in increasing j: the first maximum)
                                best = ucb;
                                best_j = j;
                            }
                        }
                    }
                }
            }
            tile.argmax(best, best_j);
            if (best_j == kNone) best_j = 0;  // (every score NaN or -Inf: the reference asserts; the first child keeps the search going)
            const int sa = plog[cbase + best_j];
            BoxANode AN;
            load_now(an + sa, AN);
            Act act;
#pragma unroll
            for (int i = 0; i < AD; ++i) act.v[i] = AN.a[i];

            // ---- state progressive widening (src/dpw.jl:119-141) ----
            bool new_node = false;
            int spnode;
            double r;
            State sp;
            bool widen = AN.nac == 0;
            if (!widen && a.cfg.enable_state_pw) {
                const int need = AN.nac < a.cfg.nmin_state_n ? __ldg(a.cfg.nmin_state + AN.nac) : 0x7fffffff;
                widen = AN.n >= need;
            }
            if (widen) {
                if constexpr (G >= 2 && M::kCoopGen) M::gen_coop(tile, a.mdp, s, act, rng, sp, r);
                else M::gen(a.mdp, s, act, rng, sp, r);
                spnode = a.cfg.check_repeat_state ? map_lookup(sp) : -1;
                if (spnode < 0) {
                    spnode = dpw_insert_state(sp, a.cfg.maintain_lookup != 0);
                    if (spnode < 0) {
                        leaf = 2;
                        break;
                    }
                    new_node = true;
                }
                // push!(transitions[sanode], (spnode, r)); n_a_children counts the distinct (sanode, spnode) pairs
                int rec = -1;
                if (!new_node) {
                    for (rec = AN.thead; rec >= 0; rec = trans[rec].next)
                        if (trans[rec].sp == spnode) break;
                }
                if (rec < 0 && n_trans >= a.geo.NT) {
                    status = MCTSB200_ERR_CAPACITY;
                    leaf = 2;
                    break;
                }
                int pb = AN.pbase, pc = AN.pcap;
                if (!vec_push(pb, pc, AN.tpush, rec >= 0 ? rec : n_trans, &an[sa].pbase, &an[sa].pcap)) {
                    leaf = 2;
                    break;
                }
                if (lead()) {
                    if (rec >= 0) {
                        bump_count(&trans[rec].count);
                    } else {
                        TransRec t;
                        t.r = r;
                        t.sp = spnode;
                        t.next = AN.thead;
                        t.count = 1;
                        t.pad[0] = t.pad[1] = t.pad[2] = 0;
                        trans[n_trans] = t;
                        an[sa].thead = n_trans;
                        an[sa].nac = AN.nac + 1;
                    }
                    an[sa].tpush = AN.tpush + 1;
                }
                if (rec < 0) n_trans += 1;
            } else {
                const int j = rng.uniform_int(AN.tpush);  // rand(rng, transitions[sanode])
                const int rec = plog[AN.pbase + j];
                spnode = trans[rec].sp;
                r = trans[rec].r;
                if (lead()) cnt[CNT_TRANS_SAMPLES] += 1;
            }
            if (lead()) path[top * path_stride] = PathEntry{node, sa, r};
            ++top;
            tile.sync();
            node = spnode;
            --d;
            if (new_node) {
                s = sp;
                leaf = 1;  // estimate_value(sp, d-1)
                break;
            }
        }
        if (leaf == 2) return;
        const double v0 = leaf == 1 ? estimate(s, d, it) : 0.0;
        // backup (src/dpw.jl:149-151), deepest level first
        tile.sync();
        if (lead()) {
            const double gamma = M::discount(a.mdp);
            double v = v0;
            for (int lvl = top - 1; lvl >= 0; --lvl) {
                const PathEntry e = path[lvl * path_stride];
                v = DM_ADD(e.r, DM_MUL(gamma, v));
                BoxANode* c = an + e.slot;
                const int nn = c->n + 1;
                const double qo = c->q;
                rows[e.node].total_n += 1;
                c->n = nn;
                c->q = DM_ADD(qo, div_by_count(DM_SUB(v, qo), i2d(nn)));
            }
        }
        tile.sync();
        if (lead()) cnt[CNT_SIMULATIONS] += 1;
    }
};

// -------------------------------------------------------------------------------------------------
// The search kernel: iterations [iter_begin, iter_end) of every tree of the call.
// -------------------------------------------------------------------------------------------------
template <class M, int KIND, int G>
__global__ void __launch_bounds__(kBlockThreads) search_kernel(const __grid_constant__ KernelArgs<M> args) {
    __shared__ unsigned long long s_cnt[CNT_COUNT];
#ifndef MCTSB200_HOST_EMU
    if (threadIdx.x < CNT_COUNT) s_cnt[threadIdx.x] = 0ull;
#if MCTSB200_TILE_TAB > 0
    __shared__ double s_ucb[2][kTileTab];
    for (int i = threadIdx.x; i < kTileTab; i += kBlockThreads) {
        s_ucb[0][i] = i < args.cfg.log_tab_n ? DM_MUL(args.cfg.c, args.cfg.sqrtlog_tab[i]) : 0.0;
        s_ucb[1][i] = i < args.cfg.rsqrt_tab_n ? args.cfg.rsqrt_tab[i] : 0.0;
    }
#endif
    __syncthreads();
#endif

    const long long slot = ((long long)blockIdx.x * kBlockThreads + threadIdx.x) / G;
    const long long tree = slot >= args.n_slots ? -1 : args.order ? (long long)args.order[slot] : slot;
    if (tree >= 0) {
        Search<M, G> S(args, tree);
#ifndef MCTSB200_HOST_EMU
#if MCTSB200_TILE_TAB > 0
        S.s_cs = s_ucb[0];
        S.s_rs = s_ucb[1];
#endif
        if (args.path_smem) {  // [level][tile]: tiles at the same level touch consecutive 16 B entries
            extern __shared__ __align__(16) unsigned char dyn_smem[];
            // (spread batches keep their trees in the first path_tiles tiles of each warp: only those get a stack)
            const int tpw = args.path_tiles ? args.path_tiles : 32 / G;
            S.path = reinterpret_cast<PathEntry*>(dyn_smem) + (threadIdx.x >> 5) * tpw + (threadIdx.x & 31) / G;
            S.path_stride = (kBlockThreads / 32) * tpw;
        }
#endif
        TreeHeader* hd = args.hdr + tree;
        TreeHeader h = *hd;
        S.n_state = h.n_state;
        S.n_action = h.n_action;
        S.n_trans = h.n_trans;
        S.n_plog = h.n_plog;
        S.status = h.status;
        int root = h.root;

        if (args.iter_begin == 0 && h.iterations == 0 && h.status == MCTSB200_OK) {
            // root lookup / insertion (src/vanilla.jl:219-224, src/dpw.jl:22-42)
            const typename M::State s0 = M::load(args.mdp, args.roots[tree * args.root_stride]);
            if (KIND == MCTSB200_KIND_DPW && M::is_terminal(args.mdp, s0)) {
                S.status = MCTSB200_ERR_TERMINAL_ROOT;
            } else {
                int id = S.n_state > 0 ? S.map_lookup(s0) : -1;
                if (id < 0) {
                    if constexpr (M::kActionDoubles > 0) id = S.dpw_insert_state(s0, S.n_state > 0 ? true : args.cfg.check_repeat_state != 0);
                    else if (KIND == MCTSB200_KIND_UCT) id = S.uct_insert(s0);
                    else id = S.dpw_insert_state(s0, S.n_state > 0 ? true : args.cfg.check_repeat_state != 0);
                }
                root = id < 0 ? 0 : id;
            }
            S.tile.sync();
        }

        int it = args.iter_begin;
        for (; it < args.iter_end && S.status == MCTSB200_OK; ++it) {
            if constexpr (M::kActionDoubles > 0) S.dpw_box_iteration(root, (uint32_t)it);
            else if (KIND == MCTSB200_KIND_UCT) S.uct_iteration(root, (uint32_t)it);
            else S.dpw_iteration(root, (uint32_t)it);
        }

#ifdef MCTSB200_PHASE_TIMING
#ifdef MCTSB200_PT_ALL
        if ((threadIdx.x & 31) == 0)
#else
        if ((threadIdx.x & 31) == 0 && (blockIdx.x % 97) == 0 && threadIdx.x < 64)
#endif
            printf("PT block %d warp %d: iters %d descent %lld leaf %lld backup %lld (cycles/iter) trips/iter %.2f\n", (int)blockIdx.x, (int)(threadIdx.x >> 5),
                   args.iter_end - args.iter_begin, S.cyc[0] / (args.iter_end - args.iter_begin), S.cyc[1] / (args.iter_end - args.iter_begin),
                   S.cyc[2] / (args.iter_end - args.iter_begin), (double)S.cyc[3] / (args.iter_end - args.iter_begin));
#endif
        if (S.lead()) {
            h.n_state = S.n_state;
            h.n_action = S.n_action;
            h.n_trans = S.n_trans;
            h.n_plog = S.n_plog;
            h.status = S.status;
            h.root = root;
            h.iterations = S.status == MCTSB200_OK ? args.iter_end : it;
            *hd = h;
        }
        S.cnt[CNT_ROLLOUT_STEPS] = 0;
#pragma unroll
        for (int i = 0; i < CNT_COUNT; ++i)
            if (S.cnt[i]) atomicAdd(&s_cnt[i], (unsigned long long)S.cnt[i]);
        if (S.rollout_steps) atomicAdd(&s_cnt[CNT_ROLLOUT_STEPS], S.rollout_steps);
    }
#ifdef MCTSB200_HOST_EMU
    for (int i = 0; i < CNT_COUNT; ++i) {  // the emulator runs one thread at a time
        args.counters[i] += s_cnt[i];
        s_cnt[i] = 0;
    }
#else
    __syncthreads();
    if (threadIdx.x < CNT_COUNT && s_cnt[threadIdx.x]) atomicAdd(args.counters + threadIdx.x, s_cnt[threadIdx.x]);
#endif
}

// A plan call that continues kept trees (reuse_tree / keep_tree): the trees stay, the search starts over.
static __global__ void begin_call_kernel(TreeHeader* hdr, long long n_trees) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    hdr[t].iterations = 0;
    hdr[t].status = MCTSB200_OK;
}

// Episode driver (mctsb200_run_episodes): the environment side of simulate(::RolloutSimulator, mdp, planner, s0).
template <class M>
__global__ void episode_init_kernel(const typename M::Params mdp, const typename M::AbiState* s, long long n, double* disc, double* ret, int* steps,
                                    int* active, int* n_active) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool run = !M::is_terminal(mdp, M::load(mdp, s[i]));
    disc[i] = 1.0;
    ret[i] = 0.0;
    steps[i] = 0;
    active[i] = run ? 1 : 0;
    if (run) atomicAdd(n_active, 1);
}
// one environment step of every running episode: sp, r = gen(mdp, s, a, rng); r_total += disc*r; disc *= discount
template <class M>
__global__ void episode_step_kernel(const typename M::Params mdp, typename M::AbiState* s, const int* action, const double* action_vec, long long n,
                                    unsigned long long env_seed, long long base, int t, double* disc, double* ret, int* steps, int* active,
                                    int* n_active) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !active[i]) return;
    Stream rng;
    rng.init(env_seed, (uint64_t)(base + i), (uint32_t)t, 0u);
    const typename M::State cur = M::load(mdp, s[i]);
    typename M::State sp;
    double r;
    if constexpr (M::kActionDoubles > 0) {  // the planner's decision is a vector (decide_box_kernel wrote it)
        typename M::Action av;
#pragma unroll
        for (int k = 0; k < M::kActionDoubles; ++k) av.v[k] = action_vec[i * M::kActionDoubles + k];
        M::gen(mdp, cur, av, rng, sp, r);
    } else {
        M::gen(mdp, cur, action[i], rng, sp, r);
    }
    ret[i] = DM_ADD(ret[i], DM_MUL(disc[i], r));
    disc[i] = DM_MUL(disc[i], M::discount(mdp));
    s[i] = M::store(mdp, sp);
    steps[i] += 1;
    if (M::is_terminal(mdp, sp)) active[i] = 0;
    else atomicAdd(n_active, 1);
}

// Root decision: best_sanode_Q (src/vanilla.jl:339-349) / best_sanode (src/dpw.jl:163-173).
template <class M>
__global__ void decide_kernel(const TreeHeader* hdr, const Row<M>* rows, int NS, long long n_trees, int* action_out,
                              double* best_q_out, int* status_out, unsigned long long* status_flags) {
    const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    const TreeHeader h = hdr[tree];
    int act = 0;
    double bq = 0.0;
    if (h.status != MCTSB200_ERR_TERMINAL_ROOT && h.n_state > 0) {
        const Row<M>* rw = rows + tree * NS + h.root;
        double best = -CUDART_INF;
        const int nchild = rw->nchild();
        int slot = nchild > 0 ? 0 : -1;  // vanilla default: first child; DPW: 0 children -> none
        for (int j = 0; j < nchild; ++j) {
            if (rw->q[j] > best) {
                best = rw->q[j];
                slot = j;
            }
        }
        if (slot >= 0) {
            act = rw->alabel(slot);
            bq = rw->q[slot];
        }
    }
    if (action_out) action_out[tree] = act;
    if (best_q_out) best_q_out[tree] = bq;
    if (status_out) status_out[tree] = h.status;
    // the call's worst status reaches the host with the counters even when the caller passed no status buffer
    if (h.status != MCTSB200_OK) atomicOr(status_flags, h.status == MCTSB200_ERR_TERMINAL_ROOT ? 1ull : 2ull);
}

// Root decision of a continuous-action tree: best_sanode (src/dpw.jl:163-173) over the root's children list; writes the chosen
// child's position, its Q and its action vector.
template <class M>
__global__ void decide_box_kernel(const TreeHeader* hdr, const Aux<M>* aux, const BoxANode* anodes, const int* plog, Geometry geo, long long n_trees,
                                  int* action_out, double* best_q_out, int* status_out, double* action_vec_out, unsigned long long* status_flags) {
    const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    const TreeHeader h = hdr[tree];
    int pos = -1;
    double bq = 0.0;
    double av[4] = {0.0, 0.0, 0.0, 0.0};
    if (h.status != MCTSB200_ERR_TERMINAL_ROOT && h.n_state > 0) {
        const Aux<M>& ax = aux[tree * geo.NS + h.root];
        const BoxANode* an = anodes + tree * geo.NA;
        const int* kids = plog + tree * geo.NP + ax.pbase[0];
        double best = -CUDART_INF;
        for (int j = 0; j < ax.nac[0]; ++j) {
            const BoxANode& c = an[kids[j]];
            if (c.q > best) {
                best = c.q;
                pos = j;
            }
        }
        if (pos >= 0) {
            const BoxANode& c = an[kids[pos]];
            bq = c.q;
            for (int i = 0; i < 4; ++i) av[i] = c.a[i];
        }
    }
    if (action_out) action_out[tree] = pos;
    if (best_q_out) best_q_out[tree] = bq;
    if (status_out) status_out[tree] = h.status;
    if (action_vec_out)
        for (int i = 0; i < M::kActionDoubles; ++i) action_vec_out[tree * M::kActionDoubles + i] = av[i];
    if (h.status != MCTSB200_OK) atomicOr(status_flags, h.status == MCTSB200_ERR_TERMINAL_ROOT ? 1ull : 2ull);
}

// Root-parallel partial sums (BASELINE config 5): sum_t N_t(a) and sum_t N_t(a)*Q_t(a) over the roots of all trees, in INTEGER
// arithmetic so that the result does not depend on the order of the trees, on how they are sharded over devices or on the
// reduction order of the allreduce that merges the devices' partial sums: each (once-rounded) double product N*Q is taken as a
// multiple of 2^-64 (rounded toward zero) in 128-bit two's complement and accumulated in four 32-bit limbs held in int64
// counters; the host recombines the limbs and rounds once to double (capi.cu, root_sums_to_double).
// out: [n[A] | limbs[A][4] | inexact flag]  (kRootSumsWords(A) int64 words)
__host__ __device__ constexpr int kRootSumsWords(int A) { return 5 * A + 1; }
// false if p is not finite or |p| >= 2^40 (then the sum is flagged inexact and the decision reports NaN)
__host__ __device__ __forceinline__ bool fix64_limbs(double p, long long (&c)[4]) {
    const uint64_t b = dm_bits(p);
    const int ex = (int)((b >> 52) & 0x7ffu);
    if (ex == 0x7ff) return false;
    const uint64_t mant = (b & 0x000fffffffffffffULL) | (ex ? 0x0010000000000000ULL : 0ULL);
    const int e = (ex ? ex : 1) - 1075 + 64;  // fixed = mant * 2^e
    if (e > 51) return false;                 // |p| >= 2^40
    uint64_t lo, hi;
    if (e >= 0) {
        lo = mant << e;
        hi = e == 0 ? 0ULL : mant >> (64 - e);
    } else {
        lo = -e >= 64 ? 0ULL : mant >> (-e);
        hi = 0ULL;
    }
    if (b >> 63) {  // two's complement negation
        lo = ~lo + 1ULL;
        hi = ~hi + (lo == 0ULL ? 1ULL : 0ULL);
    }
    c[0] = (long long)(lo & 0xffffffffULL);
    c[1] = (long long)(lo >> 32);
    c[2] = (long long)(hi & 0xffffffffULL);
    c[3] = (long long)hi >> 32;  // signed top limb
    return true;
}
template <class M>
__global__ void root_sums_kernel(const TreeHeader* hdr, const Row<M>* rows, int NS, long long n_trees, unsigned long long* out) {
    constexpr int A = M::kActions, W = kRootSumsWords(A);
    __shared__ unsigned long long acc[W];
#ifdef MCTSB200_HOST_EMU
    unsigned long long* dst = out;  // (the emulator runs one thread at a time)
#else
    unsigned long long* dst = acc;
    for (int i = threadIdx.x; i < W; i += blockDim.x) acc[i] = 0ULL;
    __syncthreads();
#endif
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_trees) {
        const TreeHeader h = hdr[t];
        if (h.n_state != 0) {
            const Row<M>* rw = rows + t * NS + h.root;
            for (int j = 0; j < rw->nchild(); ++j) {
                const int act = rw->alabel(j);
                const int nn = rw->n[j];
                long long c[4];
                atomicAdd(dst + act, (unsigned long long)(long long)nn);
                if (fix64_limbs(DM_MUL((double)nn, rw->q[j]), c)) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) atomicAdd(dst + A + 4 * act + i, (unsigned long long)c[i]);
                } else {
                    atomicOr(dst + 5 * A, 1ULL);
                }
            }
        }
    }
#ifndef MCTSB200_HOST_EMU
    __syncthreads();
    for (int i = threadIdx.x; i < W; i += blockDim.x)
        if (acc[i]) {
            if (i == 5 * A) atomicOr(out + i, acc[i]);
            else atomicAdd(out + i, acc[i]);
        }
#endif
}

// Lane ordering for dense-state MDPs: a counting sort of the trees by root state, so that trees planning from
// equal roots (statistically alike, and for deterministic models identical, amounts of work per iteration) run
// in the same warps and a warp's lanes stay converged. Every group of equal roots is padded to whole warps
// (idle slots hold -1), and the warps are then dealt round-robin to the thread blocks, so that every block (and
// SM) receives an evenly spaced sample of the groups instead of a run of warps with the same, possibly
// long-running, root. The order only decides which lane works on which tree; results do not depend on it.
template <class M>
__global__ void order_hist_kernel(const typename M::Params mdp, const typename M::AbiState* roots, long long n_trees, int* bins) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_trees) atomicAdd(bins + M::dense_index(mdp, M::load(mdp, roots[t])), 1);
}
// bins[0..n): group sizes -> first slot of each group, every group rounded up to a multiple of `tpw` slots (one block)
// slot -> tree map that leaves lanes idle: only the first `active` lanes of every warp hold a tree. Trees whose lanes
// diverge (stochastic transitions / policies) then wait for the slowest of `active` neighbours instead of 32, and a
// batch too small to fill the SMs gets more warps to hide latency with.
static __global__ void order_spread_kernel(int* order, long long n_slots, int active, long long n_trees) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const long long warp = slot >> 5;
    const int lane = (int)(slot & 31);
    const long long tree = warp * active + lane;
    order[slot] = (lane < active && tree < n_trees) ? (int)tree : -1;
}

static __global__ void order_scan_kernel(int* bins, int n, int tpw) {
#ifdef MCTSB200_HOST_EMU
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < n; ++i) {
            const int c = (bins[i] + tpw - 1) / tpw * tpw;
            bins[i] = run;
            run += c;
        }
    }
#else
    __shared__ int part[1024];
    const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += (bins[i] + tpw - 1) / tpw * tpw;
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) {
            const int c = part[i];
            part[i] = run;
            run += c;
        }
    }
    __syncthreads();
    int run = part[threadIdx.x];
    for (int i = lo; i < hi; ++i) {
        const int c = (bins[i] + tpw - 1) / tpw * tpw;
        bins[i] = run;
        run += c;
    }
#endif
}
// slot = group start + arrival rank; its warp w goes to block w mod n_blocks, position w / n_blocks in that block
template <class M>
__global__ void order_scatter_kernel(const typename M::Params mdp, const typename M::AbiState* roots, long long n_trees, int* bins, int* order,
                                     int tpw, long long n_blocks, long long warps_per_block) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    const int slot = atomicAdd(bins + M::dense_index(mdp, M::load(mdp, roots[t])), 1);
    const long long w = slot / tpw;
    order[((w % n_blocks) * warps_per_block + w / n_blocks) * tpw + slot % tpw] = (int)t;
}

// dynamic shared memory of one block of the G variant: the path stacks, if they fit
__host__ __device__ static inline size_t path_smem_bytes(int depth, int G, int path_tiles = 0) {
    return sizeof(PathEntry) * (size_t)depth * (size_t)(kBlockThreads / 32) * (size_t)(path_tiles ? path_tiles : 32 / G);
}
constexpr size_t kPathSmemLimit = 48 * 1024 - 2 * sizeof(double) * kTileTab - 512;  // (static + dynamic shared memory of a block stay below the 48 KB that need no opt-in)

#ifndef __CUDACC_RTC__  // (host-side launchers: not part of a run-time compiled plugin program, plugin_rt.cu)
template <class M, int KIND>
void launch_search(const KernelArgs<M>& args, int G, cudaStream_t stream);
template <class M, int KIND>
int search_occupancy(int G, size_t smem_bytes);  // resident blocks per SM of the G variant
#endif

}  // namespace mctsb200
