// common.cuh -- shared device/host PODs of libmctsb200: Philox streams, per-tree tables, kernel arguments.
//
// Storage model ("node tables", one set per ctx, all in HBM):
//   for each component X there is ONE array covering all trees of the plan call, indexed
//   [tree][entry] (struct-of-arrays at the table level):
//     hdr   [n_trees]                 TreeHeader   counters, root id, status
//     rows  [n_trees][NS]             Row<M>       one 64 B row per state node: q[A], key, n[A], total_n,
//                                                  child count + child action labels (2 sectors, read with
//                                                  16 B vector loads)
//     aux   [n_trees][NS]             Aux<M>       DPW only: n_a_children, transition list head, push
//                                                  count and insertion sequence number per child
//     trans [n_trees][NT]             TransRec     DPW only: distinct (spnode, r) records with counts
//     map   [n_trees][MAP]            u16 / HashSlot   state -> node id (dense index or open addressing)
//     path  [n_trees][depth]          PathEntry    the explicit recursion stack of one simulation (only used when
//                                                  the stacks of a block do not fit in shared memory)
// Node ids on the device are 0-based; the ABI export converts to the reference's 1-based ids.
#pragma once
#include <stdint.h>

#include "../../include/mctsb200.h"
#include "../../include/mctsb200_detmath.h"

// Kernel launch. MCTSB200_HOST_EMU is defined only by tests/emu (a single-threaded host build of this
// library used to exercise the host logic without a GPU); the shipped library never defines it.
#ifdef MCTSB200_HOST_EMU
#define MCTSB200_LAUNCH(kernel, grid, block, smem, stream, ...) mctsb200_emu::run((grid), (block), [&] { kernel(__VA_ARGS__); })
#else
#define MCTSB200_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#if !defined(MCTSB200_HOST_EMU) && !defined(__CUDACC_RTC__)
#include <cstdlib>
// Shared memory and L1 share the SM's 256 KB. Left to itself the driver gives a kernel with a large dynamic request the maximum
// carve-out (L1: 28 KB); these kernels want every KB of L1 their block leaves over (32 768 roots: 29.2 -> 27.8 ms,
// profiles/r2_experiments.md), so the launchers of the ONE-BLOCK-PER-SM kernels (uct_fast, dpw_fast) ask for the smallest carve-out
// that holds the block. Not for the tile kernel: its small blocks share an SM, and a carve-out sized for one of them leaves room
// for one of them (config 4: 445 -> 612 ms, a second wave of blocks); the driver's default is right there.
// MCTSB200_CARVEOUT=<percent> overrides (experiments).
namespace mctsb200 {
template <class K>
inline void prefer_smallest_carveout(K kernel, size_t dynamic_smem) {
    static const int max_smem = [] {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        return v > 0 ? v : 228 * 1024;
    }();
    int pct = (int)(((dynamic_smem + 2048) * 100 + (size_t)max_smem - 1) / (size_t)max_smem);
    if (const char* ev = std::getenv("MCTSB200_CARVEOUT"))
        if (ev[0]) pct = std::atoi(ev);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
}
}  // namespace mctsb200
#endif

#define MCTSB200_MAX_ACTIONS_PER_ROW 7

namespace mctsb200 {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 streams (contract: include/mctsb200.h)
// ---------------------------------------------------------------------------------------------
struct Stream {
    uint32_t k0, k1;
    uint32_t c0, c1, c2, c3;
    uint32_t b0, b1, b2, b3;
    int have;  // unread words of the current block, kept at the front: b0 is the next one
    int pend;  // words consumed unread since the last block ran out (skip): their blocks are never computed

    __host__ __device__ __forceinline__ void init(uint64_t seed, uint64_t tree, uint32_t iteration, uint32_t stream_id) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        c0 = 0;
        c1 = iteration;
        c2 = (uint32_t)tree;
        c3 = (uint32_t)(((tree >> 32) & 0xffffu) << 16) | (stream_id & 0xffffu);
        have = 0;
        pend = 0;
    }
    __host__ __device__ __forceinline__ void refill() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, ka = k0, kb = k1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int round = 0; round < 10; ++round) {
#if defined(__CUDA_ARCH__)
            const uint32_t lo0 = 0xD2511F53u * x0, hi0 = __umulhi(0xD2511F53u, x0);
            const uint32_t lo1 = 0xCD9E8D57u * x2, hi1 = __umulhi(0xCD9E8D57u, x2);
#else
            const uint64_t p0 = (uint64_t)0xD2511F53u * x0, p1 = (uint64_t)0xCD9E8D57u * x2;
            const uint32_t lo0 = (uint32_t)p0, hi0 = (uint32_t)(p0 >> 32), lo1 = (uint32_t)p1, hi1 = (uint32_t)(p1 >> 32);
#endif
            const uint32_t n0 = hi1 ^ x1 ^ ka, n2 = hi0 ^ x3 ^ kb;
            x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
            ka += 0x9E3779B9u;
            kb += 0xBB67AE85u;
        }
        b0 = x0; b1 = x1; b2 = x2; b3 = x3;
        c0 += 1;
        have = 4;
    }
    __host__ __device__ __forceinline__ uint32_t next() {
        if (have == 0) {
            if (pend) {  // words skipped at a block boundary: jump over their whole blocks, drop the rest from the front of this one
                c0 += (uint32_t)pend >> 2;
                const int drop = pend & 3;
                pend = 0;
                refill();
                if (drop >= 1) { b0 = b1; b1 = b2; b2 = b3; }
                if (drop >= 2) { b0 = b1; b1 = b2; }
                if (drop >= 3) { b0 = b1; }
                have -= drop;
            } else {
                refill();
            }
        }
        const uint32_t w = b0;
        b0 = b1; b1 = b2; b2 = b3;
        --have;
        return w;
    }
    // Consumes one word without looking at it. A counter-based stream makes that free: no block is computed for words nobody
    // reads. (A DPW level whose state node already owns every action still draws next_action, src/dpw.jl:96, but the draw cannot
    // change anything: the level walks on without the ten Philox rounds and later levels find the stream where the reference's is.)
    __host__ __device__ __forceinline__ void skip() {
        if (have > 0) {
            b0 = b1; b1 = b2; b2 = b3;
            --have;
        } else {
            ++pend;
        }
    }
    __host__ __device__ __forceinline__ int uniform_int(int n) {
#if defined(__CUDA_ARCH__)
        return (int)__umulhi(next(), (uint32_t)n);
#else
        return (int)(((uint64_t)next() * (uint64_t)n) >> 32);
#endif
    }
    __host__ __device__ __forceinline__ double uniform01() { return (double)next() * (1.0 / 4294967296.0); }
};

// ---------------------------------------------------------------------------------------------
// Per-tree tables
// ---------------------------------------------------------------------------------------------
struct TreeHeader {
    int n_state;     // state nodes in use
    int n_action;    // action nodes inserted so far (DPW insertion sequence; UCT: A * n_state)
    int n_trans;     // transition records in use
    int root;        // 0-based root node id
    int status;      // MCTSB200_OK / ERR_TERMINAL_ROOT / ERR_CAPACITY
    int iterations;  // iterations completed
    int n_plog;      // push-log arena entries in use (unbounded-branching models)
    int pad;
};

struct __align__(16) PathEntry {
    int node;
    int slot;
    double r;
};

struct __align__(32) TransRec {
    double r;
    int sp;     // 0-based state node
    int next;   // next (older) record of the same action node, -1 = end
    int count;  // multiplicity in the reference's transitions[sanode] vector
    int pad[3];
};

// one (action node => next state node) edge of a vanilla tree's _vis_stats (enable_tree_vis): count == 0 = free
struct __align__(8) VisRec {
    int sp;
    int count;
};

struct __align__(8) HashSlot {
    uint32_t tag;  // 0 = empty; else (hash | 1)
    int node;
};

// One state node (64 B for both built-in MDPs = two 32 B sectors).
// UCT: children are all A actions in actions(mdp) order (child k = action k; reference action-node id =
// A*node + k + 1). DPW: children live in slots [0, nchild) in insertion order; `packed` holds the child count
// (bits 28-31) and each child's action index (4 bits per slot).
// Field order keeps q[] and n[] 16 B aligned for A = 4 so one row is read with three vector loads.
template <class M>
struct __align__(32) Row {
    double q[M::kActions];
    int n[M::kActions];
    int total_n;
    uint32_t packed;
    typename M::Key key;

    __host__ __device__ __forceinline__ int nchild() const { return (int)(packed >> 28); }
    __host__ __device__ __forceinline__ int alabel(int slot) const { return (int)((packed >> (4 * slot)) & 15u); }
    __host__ __device__ __forceinline__ static uint32_t pack_add(uint32_t pk, int slot, int act) {
        return (pk & 0x0fffffffu & ~(15u << (4 * slot))) | ((uint32_t)act << (4 * slot)) | ((uint32_t)(slot + 1) << 28);
    }
};
static_assert(MCTSB200_MAX_ACTIONS_PER_ROW == 7, "4-bit labels, 7 slots + count");

template <class M>
struct __align__(32) Aux {
    int nac[M::kActions];    // n_a_children
    int thead[M::kActions];  // newest transition record, -1 = none
    int tpush[M::kActions];  // length of the reference's transitions[sanode] (pushes)
    int seq[M::kActions];    // 1-based insertion sequence of the action node (reference id)
    // the action node's newest transition record, mirrored (models with a single successor per (s, a) follow it without
    // touching the record table: search.cuh, dpw_iteration)
    int sp1[M::kActions];
    double r1[M::kActions];
    // models with unbounded branching (run-time plugins registered with max_branch = 0): the action node's push log -- the
    // reference's transitions[sanode] vector itself, one record id per push, so rand(rng, transitions[sanode]) is one indexed
    // load instead of a walk over up to k*n^alpha records. [pbase, pbase + pcap) in the tree's plog arena, grown by doubling.
    int pbase[M::kActions];
    int pcap[M::kActions];
};

// One action node of a continuous-action DPW tree (models with kActionDoubles > 0, e.g. GaussianBoxDev): a state node there has
// up to k_action * N^alpha_action children, so the action nodes live in a table of their own, ids in insertion order (the
// reference's action-node ids), and a state node lists its children in the tree's int arena (plog: Aux::pbase / pcap / nac of slot 0).
// 64 B = two sectors; (q, n) lead so that the selection scan reads one 16-byte word per child.
struct __align__(32) BoxANode {
    double q;
    int n;
    int nac;    // n_a_children
    int thead;  // newest transition record, -1 = none
    int tpush;  // length of transitions[sanode]
    int pbase;  // the push-ordered transitions[sanode] vector (record ids) in the arena: [pbase, pbase + pcap)
    int pcap;
    double a[4];  // the action label (kActionDoubles <= 4 coordinates)
};
static_assert(sizeof(BoxANode) == 64, "one action node = two sectors");

// geometry of the tables of one plan call
struct Geometry {
    int NS;         // state-node capacity per tree
    int NT;         // transition-record capacity per tree
    int MAP;        // map entries per tree (dense: n_dense u16; hash: power-of-two HashSlots)
    int depth;      // path entries per tree
    int NP;         // push-log arena entries per tree (0 unless the model has unbounded branching)
    int NA;         // action-node capacity per tree (continuous action spaces only, else 0)
};

enum CounterId {
    CNT_SIMULATIONS = 0, CNT_LEVELS, CNT_CHILDREN_SEEN, CNT_STATE_INSERTS, CNT_ACTION_INSERTS, CNT_ROLLOUTS, CNT_ROLLOUT_STEPS,
    CNT_TRANS_SAMPLES, CNT_SELECT_EXACT,
    CNT_MAX_VISITS,  // not a sum: the largest node visit count, written by the fast kernels' finalize pass
    CNT_STATUS_FLAGS,  // not a sum: OR over the trees' status (bit 0: a terminal DPW root, bit 1: a table overflowed), written by the decision kernels
    CNT_COUNT
};

// device view of the solver configuration
struct DevCfg {
    long long n_iterations;
    double c;
    double init_Q;
    double estimate_const;
    double rollout_eps;
    unsigned long long seed;
    const double* init_q_table;  // device, may be null
    const int* init_n_table;     // device, may be null
    const double* value_table;   // device, may be null
    const double* log_tab;       // device: det_log(N) for N < log_tab_n
    const double* sqrtlog_tab;   // device: sqrt(det_log(N)) for N < log_tab_n (fast-path score)
    const double* rsqrt_tab;     // device: 1/sqrt(n) for n < rsqrt_tab_n     (fast-path score)
    const int* nmin_state;       // device: min n[sanode] for which n_a_children == len may widen
    const unsigned char* policy_table;  // device, may be null: tabulated state-only rollout policy [dense index] (uct_fast.cuh)
    const unsigned char* next_action_table;  // device, may be null: DPW's next_action as a state-only function [dense index]
    int log_tab_n;
    int rsqrt_tab_n;
    int nmin_state_n;
    int nmin_action[16];         // min total_n for which a state node with len children may widen
    const int* nmin_action_tab;  // device, continuous action spaces: the same for len < nmin_action_n (a node may have many children)
    int nmin_action_n;
    int check_repeat_action;
    int depth;
    int K;
    int init_N;
    int estimate_kind;
    int policy;
    int pparam[4];
    int rollout_max_depth;
    int enable_action_pw, enable_state_pw, check_repeat_state, maintain_lookup;
    double q_bound;              // fast kernels: bound of |Q| over everything this search (and the kept trees) can hold; NaN = none
    int disc_always_positive;    // eps == 0 and discount^steps cannot underflow: the rollout guard `disc > eps` is always true
    double q_bound_tile;         // tile kernel: the same bound when it covers every value the search can back up (rollout estimates,
                                 // no init_Q table), else NaN -- then the selection margin is summed from the children instead
};

template <class M>
struct KernelArgs {
    typename M::Params mdp;
    DevCfg cfg;
    Geometry geo;
    TreeHeader* hdr;
    Row<M>* rows;
    void* rows_fast;                    // uct_fast.cuh / dpw_fast.cuh: FastRow table, [warp][state][lane]
    void* trans_fast;                   // dpw_fast.cuh: DpwTransRow table, [warp][state][lane]
    Aux<M>* aux;
    TransRec* trans;
    int* plog;                          // push-log arena, [n_trees][NP] (null unless the model has unbounded branching)
    BoxANode* anodes;                   // action nodes of continuous-action trees, [n_trees][NA] (else null)
    double* action_vec_out;             // continuous action spaces: the decision's action vector per tree (may be null)
    VisRec* vis;                        // enable_tree_vis: [n_trees][NS * kActions * vis_branch] edge visit counts (else null)
    int vis_branch;
    void* map;
    PathEntry* path;
    const typename M::AbiState* roots;  // device
    unsigned long long* counters;       // device, CNT_COUNT entries
    long long n_trees;
    long long tree_index_base;
    long long root_stride;              // 1 = one root per tree, 0 = every tree plans roots[0]
    const int* order;                   // device, may be null: tile slot -> tree or -1 (see order_*_kernel in search.cuh)
    long long n_slots;                  // tile slots of the launch (= n_trees without an order)
    int path_smem;                      // 1: the path stacks of a block live in dynamic shared memory, [level][tile]
    int path_tiles;                     // tiles per warp that own a path stack (0 = all 32/G; fewer when the trees are spread one per warp)
    int prefetch_l2;                    // fast kernels: ask L2 for the rows of a level's possible successors while the level's own row is in flight
    int iter_begin, iter_end;
};

}  // namespace mctsb200
