// uct_fast.cuh -- the lean vanilla-UCT kernel for dense-state MDPs whose model tables fit in shared memory
// (SimpleGridWorld). Same algorithm, same arithmetic and same results as search_kernel<M, UCT, 1> (search.cuh);
// what differs is the machine mapping, chosen after profiling the generic kernel (profiles/):
//
//   The search of one tree is a strictly serial chain of dependent steps (iteration n+1 reads the statistics
//   iteration n wrote; src/vanilla.jl:229-235), and a batch has far fewer trees than a B200 has lanes, so the
//   cost that matters is the LATENCY of one step of one tree, not instruction throughput. This kernel shortens
//   that chain:
//     * one lane per tree, one block per SM; every model table (fused per-(cell, action) transition entries,
//       rewards, tabulated rollout policy) and the head of the two UCB tables are staged in shared memory once
//       per block, so a rollout step touches no global memory at all and its critical path is one 29-cycle LDS
//       plus a handful of integer selects;
//     * rows are indexed directly by the state's dense index (no state -> node map probe between two levels);
//       the row of the next level is fetched as soon as the successor state is known;
//     * a global store evicts its line from L1, so in a read-modify-write backup every load pays an L2 round trip
//       (~350 cycles measured, scripts/microbench/lat.cu). The backup here loads nothing: the descent leaves a
//       16-byte record per level in shared memory ([level][lane], conflict-free) with the visited (node, child) and
//       that child's (n, q) as it was read for selection, and the backup only computes and stores. A (node, child)
//       pair that occurs twice on one path (self-loops at walls, cycles) is handled exactly: the descent stamps each
//       row it passes with the iteration and the level per child, so a repeated pair finds its previous level
//       and the backup forwards the new (n, q) into that level's record; total_n is not kept during the search
//       (it is always the sum of the children's n) and is written once by total_n_kernel afterwards;
//     * random words come from whole Philox blocks generated once per 4 levels / 2-4 rollout steps, with no
//       per-word buffer bookkeeping; every run-time option that the generic kernel tests per step (rollout
//       policy kind, estimator kind, leaf-parallel count) is a template parameter or an applicability rule here.
//
// Reference lines restated: simulate src/vanilla.jl:240-276, insert_node! :292-307, best_sanode_UCB :354-386,
// estimate_value(::SolvedRolloutEstimator) src/domain_knowledge.jl:65-73 + the POMDPTools RolloutSimulator loop.
#pragma once
#ifdef MCTSB200_PHASE_TIMING
#include <cstdio>
#endif
#include "search.cuh"

namespace mctsb200 {

constexpr int kFastMaxThreads = 512;
constexpr int kFastTab = 1024;  // entries of sqrt(log N) and 1/sqrt(n) staged in shared memory
enum { FAST_POLICY_RANDOM = 0, FAST_POLICY_CONSTANT = 1, FAST_POLICY_TABLE = 2 };

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

// bytes of dynamic shared memory of one block
template <class M>
static inline size_t fast_smem_bytes(const typename M::Params& p, int depth, int block_threads) {
    return M::fast_table_bytes(p) + 2 * sizeof(double) * kFastTab + (size_t)((M::fast_states(p) + 15) / 16 * 16) +
           sizeof(uint4) * (size_t)depth * (size_t)block_threads;
}

template <class M, int POLICY>
__global__ void __launch_bounds__(kFastMaxThreads, 1) uct_fast_kernel(const __grid_constant__ KernelArgs<M> args) {
    constexpr int A = M::kActions;
    static_assert(A == 4, "rows are read with 16-byte vector loads of four children");
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
        s_sqrtlog[i] = i < cfg.log_tab_n ? cfg.sqrtlog_tab[i] : 0.0;
    }
    if (POLICY == FAST_POLICY_TABLE)
        for (int i = st0; i < n_states; i += st1) s_pol[i] = cfg.policy_table[i];
    __syncthreads();

    const long long slot_id = (long long)blockIdx.x * BD + threadIdx.x;
    const long long tree = slot_id >= args.n_slots ? -1 : args.order ? (long long)args.order[slot_id] : slot_id;
    if (tree >= 0) {
        Row<M>* const rows = args.rows + tree * args.geo.NS;
        TreeHeader* const hd = args.hdr + tree;
        TreeHeader h = *hd;
        int n_state = h.n_state, root = h.root;
        const int status = h.status;
        const uint64_t tree_index = (uint64_t)(args.tree_index_base + tree);
        const uint64_t seed = cfg.seed;
        const int TERM = ft.term;
        const double gamma = M::discount(args.mdp), eps = cfg.rollout_eps;
        unsigned n_levels = 0, n_inserts = 0, n_rollouts = 0, n_sims = 0, n_exact = 0;
        unsigned long long n_steps = 0;

        // insert_node! (src/vanilla.jl:292-307): all A children at once; the row's key field holds the 1-based
        // insertion sequence number (= the reference's state-node id)
        auto insert = [&](int cell) {
            Row<M>* rw = rows + cell;
            const double iq = cfg.init_Q;
            const int in = cfg.init_N;
            *reinterpret_cast<double2*>(&rw->q[0]) = double2{iq, iq};
            *reinterpret_cast<double2*>(&rw->q[2]) = double2{iq, iq};
            *reinterpret_cast<int4*>(&rw->n[0]) = int4{in, in, in, in};
            uint32_t pk = 0;
#pragma unroll
            for (int k = 0; k < A; ++k) pk = Row<M>::pack_add(pk, k, k);
            ++n_state;
            // tail: [stamp (total_n's slot) | packed | seq (key) + level bytes of children 0,1 | children 2,3]
            *reinterpret_cast<int4*>(&rw->total_n) = int4{0, (int)pk, n_state, 0};
            ++n_inserts;
        };

        if (args.iter_begin == 0 && h.iterations == 0 && status == MCTSB200_OK) {  // root lookup / insertion (src/vanilla.jl:219-224)
            root = M::dense_index(args.mdp, M::load(args.mdp, args.roots[tree * args.root_stride]));
            if (rows[root].packed == 0) insert(root);
        }

#ifdef MCTSB200_PHASE_TIMING
        long long cyc[3] = {0, 0, 0};
        long long trips[3] = {0, 0, 0};
#define PT_CLOCK(var) const long long var = clock64()
#define PT_ADD(i, v) cyc[i] += (v)
#define PT_TRIP(i, v) trips[i] += (v)
#else
#define PT_CLOCK(var)
#define PT_ADD(i, v)
#define PT_TRIP(i, v)
#endif
        int it = args.iter_begin;
        if (status == MCTSB200_OK) {
            for (; it < args.iter_end; ++it) {
                PT_CLOCK(t0);
                // ---- descent: simulate(planner, node, depth), one level per trip --------------------------------
                int cell = root, d = cfg.depth, top = 0;
                bool leaf = false;
                uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
                double2 q01 = *reinterpret_cast<const double2*>(&rows[root].q[0]);
                double2 q23 = *reinterpret_cast<const double2*>(&rows[root].q[2]);
                int4 n4 = *reinterpret_cast<const int4*>(&rows[root].n[0]);
                int4 tail = *reinterpret_cast<const int4*>(&rows[root].total_n);
                for (;;) {
                    if (cell == TERM) break;  // isterminal: value 0.0, no statistics update at this node
                    if (d == 0) {
                        leaf = true;  // estimate_value(s, 0)
                        break;
                    }
                    ++n_levels;
                    if ((top & 3) == 0) philox_block(seed, tree_index, (uint32_t)it, 0u, (uint32_t)(top >> 2), b0, b1, b2, b3);
                    const uint32_t w = (top & 2) ? ((top & 1) ? b3 : b2) : ((top & 1) ? b1 : b0);
                    // best_sanode_UCB
                    typename SG::Kids k;
                    k.q[0] = q01.x; k.q[1] = q01.y; k.q[2] = q23.x; k.q[3] = q23.y;
                    k.n[0] = n4.x; k.n[1] = n4.y; k.n[2] = n4.z; k.n[3] = n4.w;
                    const int N = k.n[0] + k.n[1] + k.n[2] + k.n[3];  // total_n(snode): always the sum over its children
                    int slot;
                    if (cfg.c == 0.0) {
                        slot = 0;
                        double best = k.q[0];
#pragma unroll
                        for (int j = 1; j < A; ++j)
                            if (k.q[j] > best) {
                                best = k.q[j];
                                slot = j;
                            }
                    } else if (k.n[0] == 0 || k.n[1] == 0 || k.n[2] == 0 || k.n[3] == 0) {
                        slot = k.n[0] == 0 ? 0 : k.n[1] == 0 ? 1 : k.n[2] == 0 ? 2 : 3;  // "if action was not used, use it"
                    } else {
                        const int n_max = max(max(k.n[0], k.n[1]), max(k.n[2], k.n[3]));
                        if (N >= cfg.log_tab_n || n_max >= cfg.rsqrt_tab_n) {
                            ++n_exact;
                            slot = SG::exact_scores(Tile<1>(), cfg, k, A, N, false);
                        } else {
                            // fast-path scores S_j = q_j + (c*sqrt(log N)) * n_j^(-1/2) (see Search::select_child)
                            const double cs = DM_MUL(cfg.c, N < kFastTab ? s_sqrtlog[N] : __ldg(cfg.sqrtlog_tab + N));
                            double S[A], mag = 0.0;
#pragma unroll
                            for (int j = 0; j < A; ++j) {
                                const double term = DM_MUL(cs, k.n[j] < kFastTab ? s_rsqrt[k.n[j]] : __ldg(cfg.rsqrt_tab + k.n[j]));
                                S[j] = DM_ADD(k.q[j], term);
                                mag = DM_ADD(mag, DM_ADD(fabs(k.q[j]), term));
                            }
                            double best = -CUDART_INF, second = -CUDART_INF;
                            int best_i = SG::kNone;
#pragma unroll
                            for (int j = 0; j < A; ++j) {
                                if (S[j] > best) {
                                    second = best;
                                    best = S[j];
                                    best_i = j;
                                } else if (S[j] > second) {
                                    second = S[j];
                                }
                            }
                            const double margin = DM_MUL(5.6843418860808015e-14, mag);
                            if (best_i != SG::kNone && DM_SUB(best, second) > margin) {
                                slot = best_i;
                            } else {
                                const int r = SG::select_near_tie(Tile<1>(), &cfg, k, A, N, false, cs, margin);
                                if (r & 256) ++n_exact;
                                slot = r & 255;
                            }
                        }
                    }
                    // sp, r = @gen(:sp,:r)(mdp, s, a, rng); the reward is re-read from the table during backup
                    const int sp = M::fast_gen(ft, cell, slot, w);
                    {
                        // record the level; stamp the row so that a later level of this path on the same (node, child) finds this one
                        const uint32_t stamp = (uint32_t)it + 1u;
                        uint32_t lv = ((uint32_t)tail.z >> 16) | ((uint32_t)tail.w << 16);  // level byte per child
                        if ((uint32_t)tail.x != stamp) lv = 0xffffffffu;
                        const uint32_t link = (lv >> (8 * slot)) & 0xffu;
                        lv = (lv & ~(0xffu << (8 * slot))) | ((uint32_t)top << (8 * slot));
                        const int ns = slot == 0 ? k.n[0] : slot == 1 ? k.n[1] : slot == 2 ? k.n[2] : k.n[3];
                        const double qs = slot == 0 ? k.q[0] : slot == 1 ? k.q[1] : slot == 2 ? k.q[2] : k.q[3];
                        s_path[top * BD] = uint4{(uint32_t)cell | ((uint32_t)slot << 16) | (link << 24), (uint32_t)ns, (uint32_t)__double2loint(qs),
                                                 (uint32_t)__double2hiint(qs)};
                        *reinterpret_cast<int4*>(&rows[cell].total_n) =
                            int4{(int)stamp, tail.y, (int)(((uint32_t)tail.z & 0xffffu) | (lv << 16)), (int)(lv >> 16)};
                    }
                    ++top;
                    --d;
                    cell = sp;
                    const Row<M>* rw = rows + cell;
                    q01 = *reinterpret_cast<const double2*>(&rw->q[0]);
                    q23 = *reinterpret_cast<const double2*>(&rw->q[2]);
                    n4 = *reinterpret_cast<const int4*>(&rw->n[0]);
                    tail = *reinterpret_cast<const int4*>(&rw->total_n);
                    if (tail.y == 0) {  // new node: estimate_value(sp, depth-1), also when sp is terminal
                        insert(cell);
                        leaf = true;
                        break;
                    }
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
                    unsigned steps_here = 0;
                    while (disc > eps && cc != TERM && step <= max_steps) {
                        uint32_t w[4];
                        philox_block(seed, tree_index, (uint32_t)it, 1u, blk++, w[0], w[1], w[2], w[3]);
#pragma unroll
                        for (int u = 0; u < kStepsPerBlock; ++u) {
                            if (u > 0 && !(disc > eps && cc != TERM && step <= max_steps)) break;
                            int act;
                            uint32_t wg;
                            if (POLICY == FAST_POLICY_RANDOM) {
                                act = (int)__umulhi(w[2 * u], (uint32_t)A);  // rand(rng, actions(mdp, s))
                                wg = w[2 * u + 1];
                            } else {
                                act = POLICY == FAST_POLICY_CONSTANT ? cfg.pparam[0] : (int)s_pol[cc];
                                wg = w[u];
                            }
                            const double r = M::fast_reward(ft, cc);
                            cc = M::fast_gen(ft, cc, act, wg);
                            r_total = DM_ADD(r_total, DM_MUL(disc, r));
                            disc = DM_MUL(disc, gamma);
                            ++step;
                            ++steps_here;
                        }
                    }
                    n_steps += steps_here;
                    PT_TRIP(1, steps_here);
                    v = r_total;
                }
                PT_CLOCK(t2);

                // ---- backup (src/vanilla.jl:272-274), deepest level first; no loads from the node table ----------------
                for (int lvl = top - 1; lvl >= 0; --lvl) {
                    const uint4 e = s_path[lvl * BD];
                    const int node = (int)(e.x & 0xffffu), slot = (int)((e.x >> 16) & 0xffu);
                    const uint32_t link = e.x >> 24;
                    const double qv = DM_ADD(M::fast_reward(ft, node), DM_MUL(gamma, v));
                    const int nn = (int)e.y + 1;
                    const double qo = __hiloint2double((int)e.w, (int)e.z);
                    const double qn = DM_ADD(qo, DM_DIV(DM_SUB(qv, qo), SG::i2d(nn)));
                    Row<M>* rw = rows + node;
                    rw->n[slot] = nn;
                    rw->q[slot] = qn;
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
        if ((threadIdx.x & 31) == 0 && (blockIdx.x % 37) == 0 && threadIdx.x < 64 && n_sims)
            printf("PT block %d warp %d: iters %u | descent %lld cyc/iter %.2f levels/iter -> %.0f cyc/level | rollout %lld cyc/iter %.2f steps/iter -> %.0f cyc/step | backup %lld cyc/iter -> %.0f cyc/level\n",
                   (int)blockIdx.x, (int)(threadIdx.x >> 5), n_sims, cyc[0] / n_sims, (double)trips[0] / n_sims, (double)cyc[0] / (double)max(trips[0], 1LL),
                   cyc[1] / n_sims, (double)trips[1] / n_sims, (double)cyc[1] / (double)max(trips[1], 1LL), cyc[2] / n_sims,
                   (double)cyc[2] / (double)max(trips[2], 1LL));
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

// total_n(snode) = sum of its children's n (src/vanilla.jl:296-305 sets it to the sum of init_N, :272-273 add one to
// both per backup). Written once after the search, one thread per row.
template <class M>
__global__ void total_n_kernel(Row<M>* rows, long long n_rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    Row<M>* rw = rows + i;
    if (rw->packed == 0) return;
    int tot = 0;
#pragma unroll
    for (int k = 0; k < M::kActions; ++k) tot += rw->n[k];
    rw->total_n = tot;
}

template <class M>
int32_t launch_uct_fast(const KernelArgs<M>& args, int policy, int grid, int block, size_t smem, cudaStream_t stream);

}  // namespace mctsb200
