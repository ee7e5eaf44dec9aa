// uct_fast.cuh -- the lean vanilla-UCT kernel for small dense-state MDPs whose model tables fit in shared memory
// (SimpleGridWorld up to 127 cells). Same algorithm, same arithmetic and same results as
// search_kernel<M, UCT, 1> (search.cuh); what differs is the machine mapping, chosen after profiling the tile
// kernel (profiles/):
//
//   The search of one tree is a strictly serial chain of dependent steps (iteration n+1 reads the statistics
//   iteration n wrote; src/vanilla.jl:229-235) and a batch has far fewer trees than a B200 has lanes, so what
//   counts is the LATENCY of one step of one tree and the number of LSU wavefronts it costs (ncu: the L1/shared
//   data pipe is the busiest unit). This kernel shortens the chain and the traffic:
//     * one lane per tree, one block per SM; the model (one 16-byte entry per (state, action): three sampling
//       thresholds + four successor deltas; rewards; a tabulated rollout policy) and the heads of the two UCB
//       tables live in shared memory, so a rollout step is ONE 128-bit LDS and a few integer selects and touches
//       no global memory;
//     * a state node is 64 bytes = 4 children x {q f64, n i32, tag u32}, indexed directly by the state's dense
//       index (no state -> node map) and read with two 256-bit loads; the successor's row is requested as soon as
//       the successor is known. The rows of the 32 trees of a warp are interleaved, [warp][state][lane], so the
//       warp's working set is one contiguous ~200 KB region (TLB reach) and lanes that stand on the same state
//       read neighbouring rows;
//     * a global store evicts its line from L1, so in a read-modify-write backup every load pays an L2 round trip
//       (~350 cycles, scripts/microbench/lat.cu). The backup here loads nothing: the descent leaves a 16-byte
//       record per level in shared memory ([level][lane], conflict-free) with the visited (node, child) and that
//       child's (n, q) as read for selection; the backup computes and issues one 128-bit store per level. A
//       (node, child) pair that occurs twice on one path (self-loops at walls, cycles) is handled exactly: the
//       descent tags the child with (iteration, level), a repeated pair finds its previous level in the tag and the
//       backup forwards the new (n, q) into that level's record;
//     * random words come from whole Philox blocks (one per 4 levels / 2-4 rollout steps); UCB arg-max is a
//       two-round tournament on table-based scores with an error-margin test, exact ties (bit-identical (n, q))
//       resolved inline to the lowest index and the reference's exact expression as the last resort.
//   After the last launch finalize_rows_kernel rewrites the rows in place into the Row<M> layout (total_n = sum of
//   the children's n; key = insertion sequence number) that the decision / query / export code reads.
//
// Reference lines restated: simulate src/vanilla.jl:240-276, insert_node! :292-307, best_sanode_UCB :354-386,
// estimate_value(::SolvedRolloutEstimator) src/domain_knowledge.jl:65-73 + the POMDPTools RolloutSimulator loop.
#pragma once
#include <cstdlib>
#ifdef MCTSB200_PHASE_TIMING
#include <cstdio>
#endif
#include <type_traits>

#include "search.cuh"

namespace mctsb200 {

#ifdef MCTSB200_PHASE_TIMING
__device__ __forceinline__ unsigned __smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
#endif

constexpr int kFastMaxThreads = 512;
#ifndef MCTSB200_FAST_TAB
#define MCTSB200_FAST_TAB 2048
#endif
constexpr int kFastTab = MCTSB200_FAST_TAB;  // entries of sqrt(log N) and 1/sqrt(n) staged in shared memory

enum { FAST_POLICY_RANDOM = 0, FAST_POLICY_CONSTANT = 1, FAST_POLICY_TABLE = 2 };
enum { FAST_MODE_PLAIN = 0, FAST_MODE_AHEAD = 1, FAST_MODE_DET = 2 };

// words b0..b3 of block `blk` of stream (seed, tree, iteration, stream_id) -- the contract of include/mctsb200.h
__device__ __forceinline__ void philox_block(uint64_t seed, uint64_t tree, uint32_t it, uint32_t stream_id, uint32_t blk, uint32_t& b0, uint32_t& b1,
                                             uint32_t& b2, uint32_t& b3) {
    Stream st;
    st.init(seed, tree, it, stream_id);
    st.c0 = blk;
    st.refill();
    b0 = st.b0;
    b1 = st.b1;
    b2 = st.b2;
    b3 = st.b3;
}

// One state node of the fast kernel: child k = action k.
//   tag == 0             the state is not in the tree (rows are zero-filled before the search)
//   tag = stamp<<8|lvl   stamp = (iteration+1) mod 2^24 of the last descent that took this child, lvl = its level
//                        on that path (0xff = none; a fresh node carries stamp 0)
struct __align__(16) FastChild {
    double q;
    int n;
    uint32_t tag;
};
struct __align__(32) FastRow {
    FastChild c[4];
};
static_assert(sizeof(FastRow) == 64, "one row = two 32-byte sectors");

__device__ __forceinline__ void load_row(const FastRow* rw, FastChild (&k)[4]) {
#ifdef MCTSB200_HOST_EMU
    for (int j = 0; j < 4; ++j) k[j] = rw->c[j];
#else
    unsigned long long a0, a1, a2, a3, b0, b1, b2, b3;
#ifdef MCTSB200_ROW_CG  // (experiment: rows bypass L1)
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(rw) : "memory");
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b0), "=l"(b1), "=l"(b2), "=l"(b3) : "l"(&rw->c[2]) : "memory");
#else
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(rw) : "memory");
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b0), "=l"(b1), "=l"(b2), "=l"(b3) : "l"(&rw->c[2]) : "memory");
#endif
    k[0].q = __longlong_as_double((long long)a0); k[0].n = (int)(uint32_t)a1; k[0].tag = (uint32_t)(a1 >> 32);
    k[1].q = __longlong_as_double((long long)a2); k[1].n = (int)(uint32_t)a3; k[1].tag = (uint32_t)(a3 >> 32);
    k[2].q = __longlong_as_double((long long)b0); k[2].n = (int)(uint32_t)b1; k[2].tag = (uint32_t)(b1 >> 32);
    k[3].q = __longlong_as_double((long long)b2); k[3].n = (int)(uint32_t)b3; k[3].tag = (uint32_t)(b3 >> 32);
#endif
}
// a line towards L2 without touching L1 or a register (HBM -> L2 ahead of the load that will need it)
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef MCTSB200_HOST_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
__device__ __forceinline__ void store_child(FastChild* c, double q, int n, uint32_t tag) {
    *reinterpret_cast<int4*>(c) = int4{__double2loint(q), __double2hiint(q), n, (int)tag};
}

// the reference's exact scores, out of line: needed for ~0.1 % of the selections
template <class M>
__device__ __noinline__ int exact_scores_cold(const DevCfg* cfg, typename Search<M, 1>::Kids k, int N) {
    return Search<M, 1>::exact_scores(Tile<1>(), *cfg, k, M::kActions, N, false);
}

// bytes of dynamic shared memory of one block
template <class M>
static inline size_t fast_smem_bytes(const typename M::Params& p, int depth, int block_threads) {
    return M::fast_table_bytes(p) + 2 * sizeof(double) * kFastTab + (size_t)((M::fast_states(p) + 15) / 16 * 16) +
           sizeof(uint4) * (size_t)depth * (size_t)block_threads;
}

// MODE selects the descent / transition code:
//   FAST_MODE_PLAIN   the chosen action's successor is sampled after the selection
//   FAST_MODE_AHEAD   the successors of all A actions are looked up while the node's row is still in flight (pays when the
//                     warps have idle issue slots: lanes in lockstep, or few warps per SM)
//   FAST_MODE_DET     AHEAD for an MDP whose every transition has a single outcome (host-checked): no random word is drawn
//                     for transitions at all -- the sampled successor does not depend on it
template <class M, int POLICY, int MODE>
__global__ void __launch_bounds__(kFastMaxThreads, 1) uct_fast_kernel(const __grid_constant__ KernelArgs<M> args) {
    constexpr int A = M::kActions;
    static_assert(A == 4, "a row holds exactly four children");
    constexpr bool AHEAD = MODE != FAST_MODE_PLAIN, DET = MODE == FAST_MODE_DET;
    using SG = Search<M, 1>;
#ifdef MCTSB200_HOST_EMU
    static unsigned char fast_smem[256 * 1024];
#else
    extern __shared__ __align__(16) unsigned char fast_smem[];
#endif
    __shared__ unsigned long long s_cnt[CNT_COUNT];
    const int BD = (int)blockDim.x;
    const DevCfg& cfg = args.cfg;

    // ---- stage the tables -------------------------------------------------------------------------
    typename M::FastTables ft;
    unsigned char* sm = M::fast_stage(args.mdp, fast_smem, ft);
    double* s_rsqrt = reinterpret_cast<double*>(sm);
    double* s_sqrtlog = s_rsqrt + kFastTab;
    uint8_t* s_pol = reinterpret_cast<uint8_t*>(s_sqrtlog + kFastTab);
    const int n_states = M::fast_states(args.mdp);
    // per level: x = node | child << 16 | (previous level of the same (node, child) on this path, 0xff = none) << 24,
    // y = n[child], (z, w) = q[child] as read by the descent
    uint4* s_path = reinterpret_cast<uint4*>(s_pol + (n_states + 15) / 16 * 16) + threadIdx.x;  // [level][lane]
#ifdef MCTSB200_HOST_EMU
    const int st0 = 0, st1 = 1;  // the emulator runs one thread at a time: everybody stages everything
#else
    const int st0 = (int)threadIdx.x, st1 = BD;
    if (threadIdx.x < CNT_COUNT) s_cnt[threadIdx.x] = 0ull;
#endif
    for (int i = st0; i < kFastTab; i += st1) {
        s_rsqrt[i] = i < cfg.rsqrt_tab_n ? cfg.rsqrt_tab[i] : 0.0;
        s_sqrtlog[i] = i < cfg.log_tab_n ? DM_MUL(cfg.c, cfg.sqrtlog_tab[i]) : 0.0;  // c * sqrt(log N), the product the scores start from
    }
    if (POLICY == FAST_POLICY_TABLE)
        for (int i = st0; i < n_states; i += st1) s_pol[i] = cfg.policy_table[i];
    __syncthreads();

    const long long slot_id = (long long)blockIdx.x * BD + threadIdx.x;
    const long long tree = slot_id >= args.n_slots ? -1 : args.order ? (long long)args.order[slot_id] : slot_id;
    if (tree >= 0) {
        // row of state c of this lane's tree: rows[c * 32]  ([warp][state][lane] interleaving)
        FastRow* const rows = reinterpret_cast<FastRow*>(args.rows_fast) + (slot_id >> 5) * ((long long)args.geo.NS * 32) + (slot_id & 31);
        // (one 32 x 32-bit multiply-add per row address instead of 64-bit index arithmetic)
        auto row_at = [&](int c) -> FastRow* { return reinterpret_cast<FastRow*>(reinterpret_cast<char*>(rows) + (uint32_t)c * (32u * (uint32_t)sizeof(FastRow))); };
        uint16_t* const seq = reinterpret_cast<uint16_t*>(args.map) + tree * args.geo.MAP;  // insertion sequence number per state
        TreeHeader* const hd = args.hdr + tree;
        TreeHeader h = *hd;
        int n_state = h.n_state, root = h.root;
        const int status = h.status;
        const uint64_t tree_index = (uint64_t)(args.tree_index_base + tree);
        const uint64_t seed = cfg.seed;
        const int TERM = ft.term;
        const double gamma = M::discount(args.mdp), eps = cfg.rollout_eps, c_ucb = cfg.c, q_bound = cfg.q_bound;
        const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
        unsigned n_levels = 0, n_inserts = 0, n_rollouts = 0, n_sims = 0, n_exact = 0;
        unsigned long long n_steps = 0;

        // insert_node! (src/vanilla.jl:292-307): all A children at once
        auto insert = [&](int cell) {
            FastRow* rw = row_at(cell);
#pragma unroll
            for (int k = 0; k < A; ++k) store_child(&rw->c[k], cfg.init_Q, cfg.init_N, 0xffu);
            seq[cell] = (uint16_t)(++n_state);
            ++n_inserts;
        };

        if (args.iter_begin == 0 && h.iterations == 0 && status == MCTSB200_OK) {  // root lookup / insertion (src/vanilla.jl:219-224)
            root = M::dense_index(args.mdp, M::load(args.mdp, args.roots[tree * args.root_stride]));
            if (row_at(root)->c[0].tag == 0) insert(root);
        }

#ifdef MCTSB200_PHASE_TIMING
        long long cyc[3] = {0, 0, 0};
        long long trips[3] = {0, 0, 0};
#define PT_CLOCK(var) const long long var = clock64()
#define PT_ADD(i, v) cyc[i] += (v)
#define PT_TRIP(i, v) trips[i] += (v)
        long long lv[5] = {0, 0, 0, 0, 0};  // inside a descent level: wait for the row, select, gen, record + issue, rest
#define PT_LV(i, v) lv[i] += (v)
#else
#define PT_LV(i, v)
#define PT_CLOCK(var)
#define PT_ADD(i, v)
#define PT_TRIP(i, v)
#endif
        int it = args.iter_begin;
        if (status == MCTSB200_OK) {
            for (; it < args.iter_end; ++it) {
                PT_CLOCK(t0);
                const uint32_t stamp = ((uint32_t)it + 1u) & 0xffffffu;
                // ---- descent: simulate(planner, node, depth), one level per trip --------------------------------
                int cell = root, d = cfg.depth, top = 0;
                bool leaf = false;
                uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
                FastChild k[A];
                load_row(row_at(root), k);
                for (;;) {
                    PT_CLOCK(l0);
                    // While this node's row is still on its way from L2: everything that does not need it -- the random word of
                    // this level and the successor each action would lead to (sp, r = @gen(:sp,:r)(mdp, s, a, rng) for all a).
                    // The batch's node tables are several times the L2 (65 536 trees x 101 rows x 64 B = 424 MB) and a warp's lanes walk
                    // different trees, so without help nearly every level of a warp waits for HBM on behalf of one lane or another.
                    // Whatever the action and the sampled outcome, the successor is this cell, the terminal state or one of the four
                    // grid neighbours: their rows are asked for now, a whole level before one of them is loaded.
                    if (!AHEAD && args.prefetch_l2) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const FastRow* nb = row_at(M::fast_neighbour(args.mdp, cell, j));
                            prefetch_l2(nb);
                            if (args.prefetch_l2 > 1) prefetch_l2(&nb->c[2]);  // (the row's second 32-byte sector)
                        }
                    }
                    uint32_t w = 0;
                    if (!DET) {
                        if ((top & 3) == 0) philox_block(seed, tree_index, (uint32_t)it, 0u, (uint32_t)(top >> 2), b0, b1, b2, b3);
                        w = (top & 2) ? ((top & 1) ? b3 : b2) : ((top & 1) ? b1 : b0);  // word `top` of stream 0
                    }
                    int spc[A];
                    if (AHEAD) {
#pragma unroll
                        for (int j = 0; j < A; ++j) spc[j] = DET ? M::fast_gen_det(ft, cell, j) : M::fast_gen(ft, cell, j, w);
                    }
#ifdef MCTSB200_PHASE_TIMING
                    if (AHEAD && spc[0] + spc[1] + spc[2] + spc[3] < 0) break;  // (never: makes the clock below wait for the successors)
                    PT_CLOCK(lA);
                    if (k[0].tag == 0xdeadbeefu) break;  // (never: makes the clock below wait for the row)
                    PT_CLOCK(lB);
                    PT_LV(2, lA - l0);
                    PT_LV(0, lB - lA);
#endif
                    // first use of the row (AHEAD; otherwise the row was checked as soon as it was requested, below)
                    if (AHEAD && k[0].tag == 0) {  // new node: estimate_value(sp, depth-1), also when sp is terminal
                        insert(cell);
                        leaf = true;
                        break;
                    }
                    if (cell == TERM) break;  // isterminal: value 0.0, no statistics update at this node
                    if (d == 0) {
                        leaf = true;  // estimate_value(s, 0)
                        break;
                    }
                    ++n_levels;

                    // ---- best_sanode_UCB (src/vanilla.jl:354-386) ----
                    int slot;
                    if (k[0].n == 0 || k[1].n == 0 || k[2].n == 0 || k[3].n == 0) {
                        slot = k[0].n == 0 ? 0 : k[1].n == 0 ? 1 : k[2].n == 0 ? 2 : 3;  // "if action was not used, use it"
                    } else {
                        // fast-path scores S_j = q_j + (c*sqrt(log N)) * n_j^(-1/2) from the two tables (see
                        // Search::select_child for the error bound behind the margin)
                        const int N = (k[0].n + k[1].n) + (k[2].n + k[3].n);  // total_n(snode): always the sum over its children
                        double S[A], T[A], cs;
                        if (N < kFastTab) {  // then every n_j is below kFastTab too: all five factors come from shared memory
                            cs = s_sqrtlog[N];
#pragma unroll
                            for (int j = 0; j < A; ++j) T[j] = DM_MUL(cs, s_rsqrt[k[j].n]);
                        } else if (N < cfg.log_tab_n) {
                            cs = DM_MUL(c_ucb, __ldg(cfg.sqrtlog_tab + N));
#pragma unroll
                            for (int j = 0; j < A; ++j) T[j] = DM_MUL(cs, __ldg(cfg.rsqrt_tab + k[j].n));
                        } else {
                            // beyond the tables (a max_time search that ran long): NaN scores fail every test below and send the
                            // selection to the reference's exact expression
                            cs = kNaN;
#pragma unroll
                            for (int j = 0; j < A; ++j) T[j] = kNaN;
                        }
#pragma unroll
                        for (int j = 0; j < A; ++j) S[j] = DM_ADD(k[j].q, T[j]);
                        // 2^-44 * sum(|q_j| + |term_j|) <= 2^-42 * (max|q| + |c sqrt(log N)|): q_bound is the host's bound of every Q
                        // value this search can produce (NaN when there is none, which sends every selection to the exact path)
                        const double margin = DM_MUL(2.2737367544323206e-13, DM_ADD(q_bound, fabs(cs)));
                        // two-round tournament; equal scores keep the lower index, like the reference's strict '>' scan
                        const bool m01 = S[1] > S[0], m23 = S[3] > S[2];
                        const double w01 = m01 ? S[1] : S[0], l01 = m01 ? S[0] : S[1];
                        const double w23 = m23 ? S[3] : S[2], l23 = m23 ? S[2] : S[3];
                        const bool mf = w23 > w01;
                        const double best = mf ? w23 : w01;
                        const double o1 = mf ? w01 : w23, o2 = mf ? l23 : l01;  // the other pair's winner, the loser beside the best
                        const int best_i = mf ? (m23 ? 3 : 2) : (m01 ? 1 : 0);
                        if (DM_SUB(best, o1) > margin && DM_SUB(best, o2) > margin) {
                            slot = best_i;
                        } else {
                            // near-tie: a child inside the margin with (n, q) bit-identical to the winner's has a bit-identical
                            // score -- in the reference too -- and the tournament above already kept the lowest index among equal
                            // scores; anything else inside the margin is decided by the reference's exact expression
                            const int nb = best_i == 0 ? k[0].n : best_i == 1 ? k[1].n : best_i == 2 ? k[2].n : k[3].n;
                            const double qb = best_i == 0 ? k[0].q : best_i == 1 ? k[1].q : best_i == 2 ? k[2].q : k[3].q;
                            bool ok = best == best;
#pragma unroll
                            for (int j = 0; j < A; ++j) ok &= (DM_SUB(best, S[j]) > margin) | ((k[j].n == nb) & (k[j].q == qb));
                            if (ok) {
                                slot = best_i;
                            } else {
                                ++n_exact;
                                typename SG::Kids kk;
#pragma unroll
                                for (int j = 0; j < A; ++j) {
                                    kk.q[j] = k[j].q;
                                    kk.n[j] = k[j].n;
                                }
                                slot = exact_scores_cold<M>(&cfg, kk, N);
                            }
                        }
                    }

                    // the successor of the chosen action; the reward is re-read from the table during backup
                    const int sp = AHEAD ? (slot == 0 ? spc[0] : slot == 1 ? spc[1] : slot == 2 ? spc[2] : spc[3]) : M::fast_gen(ft, cell, slot, w);
#ifdef MCTSB200_PHASE_TIMING
                    if (sp < 0) break;  // (never: makes the clock below wait for the selection)
#endif
                    PT_CLOCK(l1);
                    // what the backup needs from this level
                    const int ns = slot == 0 ? k[0].n : slot == 1 ? k[1].n : slot == 2 ? k[2].n : k[3].n;
                    const double qs = slot == 0 ? k[0].q : slot == 1 ? k[1].q : slot == 2 ? k[2].q : k[3].q;
                    const uint32_t tg = slot == 0 ? k[0].tag : slot == 1 ? k[1].tag : slot == 2 ? k[2].tag : k[3].tag;
                    {
                        // record the level; tag the child so that a later level of this path on the same (node, child) finds this one
                        const uint32_t link = (tg >> 8) == stamp ? (tg & 0xffu) : 0xffu;
                        s_path[top * BD] = uint4{(uint32_t)cell | ((uint32_t)slot << 16) | (link << 24), (uint32_t)ns, (uint32_t)__double2loint(qs),
                                                 (uint32_t)__double2hiint(qs)};
                        row_at(cell)->c[slot].tag = (stamp << 8) | (uint32_t)top;  // (before the successor's row is read: it may be this row)
                    }
                    ++top;
                    --d;
                    cell = sp;
                    load_row(row_at(sp), k);
                    if (!AHEAD && k[0].tag == 0) {  // (divergent lanes: leaving the loop here saves the trip to the top)
                        insert(cell);
                        leaf = true;
                        break;
                    }
                    PT_CLOCK(l3);
#ifdef MCTSB200_PHASE_TIMING
                    PT_LV(1, l1 - lB);
                    PT_LV(3, l3 - l1);
#endif
                }
                PT_CLOCK(t1);
                PT_TRIP(0, top);

                // ---- leaf evaluation: one rollout (POMDPTools RolloutSimulator loop) --------------------------------
                double v = 0.0;
                if (leaf) {
                    ++n_rollouts;
                    const int max_steps = cfg.rollout_max_depth == -1 ? d : cfg.rollout_max_depth;
                    constexpr int kWordsPerStep = POLICY == FAST_POLICY_RANDOM ? 2 : 1;
                    constexpr int kStepsPerBlock = 4 / kWordsPerStep;
                    double disc = 1.0, r_total = 0.0;
                    int step = 1, cc = cell;
                    uint32_t blk = 0;
                    // `disc > eps` of the reference's loop guard is provably always true for eps = 0 and these depths: that
                    // version of the loop carries no FP64 compare
                    auto run = [&](auto DCHK) {
                        constexpr bool dchk = decltype(DCHK)::value;
                        while ((!dchk || disc > eps) && cc != TERM && step <= max_steps) {
                            uint32_t w[4] = {0u, 0u, 0u, 0u};
                            if (!(DET && POLICY != FAST_POLICY_RANDOM)) philox_block(seed, tree_index, (uint32_t)it, 1u, blk++, w[0], w[1], w[2], w[3]);
#pragma unroll
                            for (int u = 0; u < kStepsPerBlock; ++u) {
                                if (u > 0 && !((!dchk || disc > eps) && cc != TERM && step <= max_steps)) break;
                                int act;
                                uint32_t wg;
                                if (POLICY == FAST_POLICY_RANDOM) {
                                    act = (int)__umulhi(w[2 * u], (uint32_t)A);  // rand(rng, actions(mdp, s))
                                    wg = w[2 * u + 1];
                                } else {
                                    act = POLICY == FAST_POLICY_CONSTANT ? cfg.pparam[0] : (int)s_pol[cc];
                                    wg = w[u];
                                }
                                if (M::fast_has_reward(ft, cc)) r_total = DM_ADD(r_total, DM_MUL(disc, M::fast_reward(ft, cc)));
                                cc = DET ? M::fast_gen_det(ft, cc, act) : M::fast_gen(ft, cc, act, wg);
                                disc = DM_MUL(disc, gamma);
                                ++step;
                            }
                        }
                    };
                    if (cfg.disc_always_positive) run(std::false_type{});
                    else run(std::true_type{});
                    n_steps += (unsigned)(step - 1);
                    v = r_total;
                }
                PT_CLOCK(t2);

                // ---- backup (src/vanilla.jl:272-274), deepest level first; no loads from the node table ----------------
                for (int lvl = top - 1; lvl >= 0; --lvl) {
                    const uint4 e = s_path[lvl * BD];
                    const int node = (int)(e.x & 0xffffu), slot = (int)((e.x >> 16) & 0xffu);
                    const uint32_t link = e.x >> 24;
                    double qv = DM_MUL(gamma, v);
                    if (M::fast_has_reward(ft, node)) qv = DM_ADD(M::fast_reward(ft, node), qv);
                    const int nn = (int)e.y + 1;
                    const double qo = __hiloint2double((int)e.w, (int)e.z);
                    const double qn = DM_ADD(qo, div_by_count(DM_SUB(qv, qo), SG::i2d(nn)));
                    store_child(&row_at(node)->c[slot], qn, nn, (stamp << 8) | 0xffu);
                    if (link != 0xffu) {  // the same (node, child) higher up on this path backs up on top of this update
                        uint4* up = s_path + link * BD;
                        up->y = (uint32_t)nn;
                        up->z = (uint32_t)__double2loint(qn);
                        up->w = (uint32_t)__double2hiint(qn);
                    }
                    v = qv;
                }
                PT_CLOCK(t3);
                PT_TRIP(2, top);
                PT_ADD(0, t1 - t0);
                PT_ADD(1, t2 - t1);
                PT_ADD(2, t3 - t2);
                ++n_sims;
            }
        }
#ifdef MCTSB200_PHASE_TIMING
        if ((threadIdx.x & 31) == 0 && n_sims)  // PT block warp smid | levels/iter steps/iter | cycles/iter: descent rollout backup | per level, step, level
            printf("PT %d %d %d | %.2f %.2f | %lld %lld %lld | %.0f %.0f %.0f | level: rowwait %.0f select %.0f pre-gen %.0f record+issue %.0f\n", (int)blockIdx.x,
                   (int)(threadIdx.x >> 5), (int)__smid(), (double)trips[0] / n_sims, (double)trips[1] / n_sims, cyc[0] / n_sims, cyc[1] / n_sims,
                   cyc[2] / n_sims, (double)cyc[0] / (double)max(trips[0], 1LL), (double)cyc[1] / (double)max(trips[1], 1LL),
                   (double)cyc[2] / (double)max(trips[2], 1LL), (double)lv[0] / (double)max(trips[0], 1LL), (double)lv[1] / (double)max(trips[0], 1LL),
                   (double)lv[2] / (double)max(trips[0], 1LL), (double)lv[3] / (double)max(trips[0], 1LL));
#endif

        h.n_state = n_state;
        h.n_action = A * n_state;
        h.n_trans = 0;
        h.root = root;
        h.iterations = status == MCTSB200_OK ? args.iter_end : it;
        *hd = h;
#ifdef MCTSB200_HOST_EMU
        unsigned long long* dst = args.counters;
#else
        unsigned long long* dst = s_cnt;
#endif
        atomicAdd(dst + CNT_SIMULATIONS, (unsigned long long)n_sims);
        atomicAdd(dst + CNT_LEVELS, (unsigned long long)n_levels);
        atomicAdd(dst + CNT_STATE_INSERTS, (unsigned long long)n_inserts);
        atomicAdd(dst + CNT_ACTION_INSERTS, (unsigned long long)n_inserts * A);
        atomicAdd(dst + CNT_ROLLOUTS, (unsigned long long)n_rollouts);
        atomicAdd(dst + CNT_ROLLOUT_STEPS, n_steps);
        atomicAdd(dst + CNT_SELECT_EXACT, (unsigned long long)n_exact);
    }
#ifndef MCTSB200_HOST_EMU
    __syncthreads();
    if (threadIdx.x < CNT_COUNT && s_cnt[threadIdx.x]) atomicAdd(args.counters + threadIdx.x, s_cnt[threadIdx.x]);
#endif
}

// After the last search launch: copy every row from the kernel's interleaved FastRow table into the Row<M> table
// [tree][state] (q[A], n[A], total_n = sum of the children's n -- src/vanilla.jl:296-305 and :272-273 keep them
// equal --, all-children label word, key = 1-based insertion sequence number), which is what decide_kernel, the root
// queries and the tree export read. One thread per (slot, state); consecutive threads read consecutive lanes.
template <class M>
__global__ void finalize_rows_kernel(const FastRow* rows_fast, const int* order, const uint16_t* seq, Row<M>* rows, int NS, long long n_slots,
                                     unsigned long long* max_visits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // ((warp * NS) + state) * 32 + lane
    const long long warp = i / ((long long)NS * 32);
    const int state = (int)((i / 32) % NS), lane = (int)(i & 31);
    const long long slot = warp * 32 + lane;
    if (slot >= n_slots) return;
    const long long tree = order ? (long long)order[slot] : slot;
    if (tree < 0) return;
    const FastRow fr = rows_fast[i];
    Row<M> out;
    int tot = 0;
    uint32_t pk = 0;
#pragma unroll
    for (int k = 0; k < M::kActions; ++k) {
        out.q[k] = fr.c[k].q;
        out.n[k] = fr.c[k].n;
        tot += fr.c[k].n;
        pk = Row<M>::pack_add(pk, k, k);
    }
    const bool present = fr.c[0].tag != 0;
    out.total_n = present ? tot : 0;
    out.packed = present ? pk : 0u;
    out.key = (typename M::Key)seq[tree * NS + state];
    rows[tree * NS + state] = out;
    // the largest visit count bounds the next call's UCB-table indices when the trees are kept
    if (present && (unsigned long long)tot > *(volatile unsigned long long*)max_visits) atomicMax(max_visits, (unsigned long long)tot);
}

template <class M>
int32_t launch_uct_fast(const KernelArgs<M>& args, int policy, int mode, int grid, int block, size_t smem, cudaStream_t stream);

}  // namespace mctsb200
