// models.cuh -- device-side MDP plugins (transition / reward / terminal test / rollout policies).
//
// A plugin is a trait struct:
//   kActions            |A| (both built-ins have a state-independent action set)
//   kDense              true if states have a dense index (lets the tree skip hashing)
//   kMaxBranch          max distinct successors of one (s,a) (bounds the transition-record table)
//   AbiState            the state as the host passes it (statetype(mdp) bytes)
//   State               the state as kept in registers
//   Key                 what a tree row stores to identify its state
//   Params              POD parameters (tables are device pointers)
//   load/store          AbiState <-> State
//   is_terminal, gen, discount, key_of, state_of, dense_index / hash
//   policy_action       rollout-policy trait
// "Registering" an MDP = adding such a struct and one line to the dispatch table in capi.cu.
//
// The arithmetic restates POMDPModels.jl's SimpleGridWorld and MountainCar (not vendored in the
// reference), with every floating-point operation individually rounded.
#pragma once
#include "common.cuh"

#ifndef MCTSB200_MC_COS_NB
#define MCTSB200_MC_COS_NB 0  // 1: MountainCar's cosine without the quadrant branch (experiment, profiles/r2_experiments.md)
#endif

namespace mctsb200 {

// ---------------------------------------------------------------------------------------------
// SimpleGridWorld
// ---------------------------------------------------------------------------------------------
struct GridWorldDev {
    static constexpr int kActions = 4;
    static constexpr bool kDense = true;
    static constexpr int kMaxBranch = 5;
    static constexpr int kMdpId = MCTSB200_MDP_GRIDWORLD;
    static constexpr bool kCustom = false;  // (custom.cuh: plugins registered at run time)
    static constexpr int kActionDoubles = 0;  // actions are indices into actions(mdp) (> 0: vectors of a continuous action space)

    struct AbiState { long long x, y; };
    // On the device a state is its dense cell index: (x-1) + (y-1)*nx, and nx*ny for the terminal state (-1,-1).
    struct State { int cell; };
    typedef uint16_t Key;

    // One 16 B record per cell (one vector load per gen call).
    struct __align__(16) CellInfo {
        double reward;       // get(rewards, s, 0.0)
        uint32_t term_from;  // s in terminate_from, or the terminal state itself
        uint32_t row;        // 4 * ob_mask: first threshold row of this cell's boundary class
    };
    // Transition sampling for one (boundary class, intended action). The reference draws u = sum(p)*rand() and
    // takes the first destination in [stay, up, down, left, right] whose running total exceeds u (SparseCat).
    // With rand() = w * 2^-32 that choice is a monotone step function of the 32-bit word w, so the host
    // evaluates the reference's floating-point expression (sample_destination) to find the step positions and
    // the device only compares integers:  j = #{i : w > t[i]},  destination = (codes >> 4j) & 15.
    // Unused slots hold t = 0xffffffff (never exceeded).
    struct __align__(32) Thresholds {
        uint32_t t[4];
        uint32_t codes;
        uint32_t pad[3];
    };

    // The same sampling table fused per (cell, action) for the shared-memory kernel (uct_fast.cuh), 16 bytes per
    // entry: a GridWorld move has at most 4 outcomes with non-zero probability (stay only at the border, where at
    // least one direction is folded into it), i.e. at most 3 steps, and on grids of up to 127 cells every successor
    // is within +-127 of the source index:  sp = cell + d[#{i : w > t[i]}].  terminate_from cells and the terminal
    // state have no step and d[0] = terminal - cell.
    struct __align__(16) FastEntry {
        uint32_t t[3];  // ascending, unused = 0xffffffff
        uint32_t d;     // four signed 8-bit successor deltas, plateau 0 in the low byte
    };
    struct FastTables {
        const FastEntry* tt;   // shared memory, [ncells + 1][4]
        const double* reward;  // shared memory, [ncells + 1]
        const uint8_t* rflag;  // shared memory, [ncells + 1]: reward(c) is not +0.0 (adding +0.0 is the exact identity, so those
                               // cells skip the arithmetic)
        int term;              // dense index of the terminal state
    };

    struct Params {
        int nx, ny, ncells;
        int delta[8];    // cell-index step of destination code [stay, up, down, left, right]
        double tprob;    // probability of the intended move
        double p_other;  // (1.0 - tprob) / 3
        double discount;
        const CellInfo* cells;      // [nx*ny + 1]
        const Thresholds* thr;      // [16 boundary classes][4 actions]
        const FastEntry* fast_tt;   // [nx*ny + 1][4], null if the grid does not qualify for the 16-byte entries
        uint32_t fast_rbits[4];
    };

    static constexpr bool kFast = true;
    __host__ __device__ __forceinline__ static int fast_states(const Params& p) { return p.ncells + 1; }
    __host__ __device__ __forceinline__ static size_t fast_table_bytes(const Params& p) {
        return ((sizeof(FastEntry) * 4 + sizeof(double) + 1) * (size_t)(p.ncells + 1) + 15) / 16 * 16;
    }
    // copies the tables into shared memory (all threads of the block; the caller synchronises); returns the
    // first free 16-byte aligned address
    __device__ __forceinline__ static unsigned char* fast_stage(const Params& p, unsigned char* smem, FastTables& ft) {
        const int n = p.ncells + 1;
        uint4* dst = reinterpret_cast<uint4*>(smem);
        const uint4* src = reinterpret_cast<const uint4*>(p.fast_tt);
        double* rew = reinterpret_cast<double*>(smem + sizeof(FastEntry) * 4 * (size_t)n);
#ifdef MCTSB200_HOST_EMU
        const int i0 = 0, di = 1;
#else
        const int i0 = (int)threadIdx.x, di = (int)blockDim.x;
#endif
        uint8_t* flg = reinterpret_cast<uint8_t*>(rew + n);
        for (int i = i0; i < n * 4; i += di) dst[i] = src[i];
        for (int i = i0; i < n; i += di) {
            rew[i] = p.cells[i].reward;
            flg[i] = (uint8_t)((p.fast_rbits[i >> 5] >> (i & 31)) & 1u);
        }
        ft.tt = reinterpret_cast<const FastEntry*>(smem);
        ft.reward = rew;
        ft.rflag = flg;
        ft.term = p.ncells;
        return smem + fast_table_bytes(p);
    }
    // dense index of the j-th state (j < 4) a move from `cell` can lead to besides `cell` itself and the terminal state: its four
    // grid neighbours, whatever the action and the sampled outcome (the fast kernels ask L2 for their rows a level ahead)
    __device__ __forceinline__ static int fast_neighbour(const Params& p, int cell, int j) {
        const int d = j == 0 ? p.nx : j == 1 ? -p.nx : j == 2 ? -1 : 1;
        return min(max(cell + d, 0), p.ncells);
    }
    __device__ __forceinline__ static bool fast_has_reward(const FastTables& ft, int cell) { return ft.rflag[cell] != 0; }
    __device__ __forceinline__ static double fast_reward(const FastTables& ft, int cell) { return ft.reward[cell]; }
    // PRMT with sign-replicating selector nibbles (the __byte_perm intrinsic only honours three bits per nibble)
    __device__ __forceinline__ static int prmt_sext(uint32_t d, uint32_t sel) {
#ifdef MCTSB200_HOST_EMU
        return (int)__byte_perm(d, 0u, sel);  // (the emulator's shim implements the full PRMT selector)
#else
        uint32_t r;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(d), "r"(0u), "r"(sel));
        return (int)r;
#endif
    }
    // (the thresholds ascend, so the last one exceeded names the outcome; the byte-permute selector j | 0x8880 + 0x1110 j picks
    // delta byte j sign-extended)
    __device__ __forceinline__ static int fast_gen(const FastTables& ft, int cell, int a, uint32_t w) {
        const uint4 e = *reinterpret_cast<const uint4*>(ft.tt + (cell * 4 + a));
        uint32_t sel = 0x8880u;
        if (w > e.x) sel = 0x9991u;
        if (w > e.y) sel = 0xaaa2u;
        if (w > e.z) sel = 0xbbb3u;
        return cell + prmt_sext(e.w, sel);
    }
    // every (state, action) has the single outcome of plateau 0 (Params::fast_det, checked by the host)
    __device__ __forceinline__ static int fast_gen_det(const FastTables& ft, int cell, int a) {
        return cell + (int)(signed char)(ft.tt[cell * 4 + a].d & 0xffu);
    }
    // Same, also returning the outcome index j < 4. Distinct outcome indices of one (state, action) are distinct
    // successors (the plateaus of the sampling table are [stay, up, down, left, right] in that order, each at most once),
    // so an action node's set of seen successors is a 4-bit mask (dpw_fast.cuh).
    __device__ __forceinline__ static int fast_gen(const FastTables& ft, int cell, int a, uint32_t w, int& j) {
        const uint4 e = *reinterpret_cast<const uint4*>(ft.tt + (cell * 4 + a));
        uint32_t sel = 0x8880u;
        if (w > e.x) sel = 0x9991u;
        if (w > e.y) sel = 0xaaa2u;
        if (w > e.z) sel = 0xbbb3u;
        j = (int)(sel & 3u);
        return cell + prmt_sext(e.w, sel);
    }
    __device__ __forceinline__ static int fast_outcome(const FastTables& ft, int cell, int a, int j) {
        return cell + (int)(signed char)(ft.tt[cell * 4 + a].d >> (8 * j));
    }
    __host__ __device__ __forceinline__ static int fast_outcome_global(const Params& p, int cell, int a, int j) {
        return cell + (int)(signed char)(p.fast_tt[cell * 4 + a].d >> (8 * j));
    }

    __device__ __forceinline__ static State load(const Params& p, const AbiState& a) {
        return State{(a.x < 0 || a.y < 0) ? p.ncells : (int)(a.x - 1) + (int)(a.y - 1) * p.nx};
    }
    __host__ __device__ __forceinline__ static AbiState store(const Params& p, const State& s) {
        if (s.cell == p.ncells) return AbiState{-1, -1};
        return AbiState{s.cell % p.nx + 1, s.cell / p.nx + 1};
    }
    __device__ __forceinline__ static bool is_terminal(const Params& p, const State& s) { return s.cell == p.ncells; }  // any(s .< 0)
    __device__ __forceinline__ static double discount(const Params& p) { return p.discount; }
    __device__ __forceinline__ static int dense_index(const Params&, const State& s) { return s.cell; }
    __device__ __forceinline__ static Key key_of(const Params&, const State& s) { return (Key)s.cell; }
    __host__ __device__ __forceinline__ static State state_of(const Params&, Key k) { return State{(int)k}; }
    __device__ __forceinline__ static bool key_equal(Key a, Key b) { return a == b; }

    // The reference's floating-point sampling expression for one (boundary class, action, word): which of the
    // five destinations [stay, up, down, left, right] is chosen. Only the host evaluates it (threshold table).
    __host__ __device__ __forceinline__ static int sample_destination(double tprob, double p_other, unsigned ob_mask, int a, uint32_t w) {
        const bool ob_up = ob_mask & 1u, ob_down = ob_mask & 2u, ob_left = ob_mask & 4u, ob_right = ob_mask & 8u;
        const double pu = a == 0 ? tprob : p_other, pd = a == 1 ? tprob : p_other;
        const double pl = a == 2 ? tprob : p_other, pr = a == 3 ? tprob : p_other;
        double p0 = 0.0;  // out-of-bounds mass folds into "stay", accumulated in action order
        if (ob_up) p0 = DM_ADD(p0, pu);
        if (ob_down) p0 = DM_ADD(p0, pd);
        if (ob_left) p0 = DM_ADD(p0, pl);
        if (ob_right) p0 = DM_ADD(p0, pr);
        const double p1 = ob_up ? 0.0 : pu, p2 = ob_down ? 0.0 : pd, p3 = ob_left ? 0.0 : pl, p4 = ob_right ? 0.0 : pr;
        const double c1 = DM_ADD(p0, p1), c2 = DM_ADD(c1, p2), c3 = DM_ADD(c2, p3), c4 = DM_ADD(c3, p4);
        const double u = DM_MUL(c4, DM_MUL((double)w, 1.0 / 4294967296.0));
        if (u < p0) return 0;
        if (u < c1) return 1;
        if (u < c2) return 2;
        if (u < c3) return 3;
        return 4;  // u < c4 always holds for rand() < 1
    }

    // transition(mdp, s, a) + rand(rng, SparseCat) + reward(mdp, s, a); one 32-bit draw unless the source is
    // a terminate_from cell or terminal (Deterministic next state, no draw).
    __device__ __forceinline__ static void gen(const Params& p, const State& s, int a, Stream& rng, State& sp, double& r) {
        const int4 raw = __ldg(reinterpret_cast<const int4*>(p.cells + s.cell));
        r = __hiloint2double(raw.y, raw.x);
        if (raw.z) {
            sp.cell = p.ncells;
            return;
        }
        const uint32_t w = rng.next();
        const Thresholds* row = p.thr + (raw.w + a);
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(row->t));
        const uint32_t codes = __ldg(&row->codes);
        const int j = (int)(w > t.x) + (int)(w > t.y) + (int)(w > t.z) + (int)(w > t.w);
        sp.cell = s.cell + p.delta[(codes >> (4 * j)) & 15u];
    }

    __device__ __forceinline__ static int policy_action(const Params& p, int policy, const int* pp, const State& s, Stream& rng) {
        switch (policy) {
            case MCTSB200_POLICY_RANDOM: return rng.uniform_int(kActions);
            case MCTSB200_POLICY_CONSTANT: return pp[0];
            case MCTSB200_POLICY_GW_SEEK: {
                const int x = s.cell % p.nx + 1, y = s.cell / p.nx + 1;
                if (pp[0] > x) return 3;
                else if (pp[0] < x) return 2;
                else if (pp[1] > y) return 0;
                else return 1;
            }
            default: return 0;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// MountainCar
// ---------------------------------------------------------------------------------------------
struct MountainCarDev {
    static constexpr int kActions = 3;
    static constexpr bool kDense = false;
    static constexpr int kMaxBranch = 1;
    static constexpr int kMdpId = MCTSB200_MDP_MOUNTAINCAR;
    static constexpr bool kCustom = false;
    static constexpr bool kFast = false;
    static constexpr int kActionDoubles = 0;

    struct AbiState { double x, v; };
    typedef AbiState State;
    struct Key { double x, v; };

    struct Params {
        double discount, cost, jackpot;
    };

    __device__ __forceinline__ static State load(const Params&, const AbiState& a) { return a; }
    __host__ __device__ __forceinline__ static AbiState store(const Params&, const State& s) { return s; }
    __device__ __forceinline__ static bool is_terminal(const Params&, const State& s) { return s.x >= 0.5; }
    __device__ __forceinline__ static double discount(const Params& p) { return p.discount; }
    __device__ __forceinline__ static int dense_index(const Params&, const State&) { return -1; }
    __device__ __forceinline__ static Key key_of(const Params&, const State& s) { return Key{s.x, s.v}; }
    __host__ __device__ __forceinline__ static State state_of(const Params&, const Key& k) { return State{k.x, k.v}; }
    // isequal on Float64 pairs == bitwise equality
    __device__ __forceinline__ static bool key_equal(const Key& a, const Key& b) {
        return __double_as_longlong(a.x) == __double_as_longlong(b.x) && __double_as_longlong(a.v) == __double_as_longlong(b.v);
    }
    __device__ __forceinline__ static uint32_t hash(const State& s) {
        uint64_t h = (uint64_t)__double_as_longlong(s.x) * 0x9E3779B97F4A7C15ull;
        h ^= (uint64_t)__double_as_longlong(s.v) + 0x7F4A7C159E3779B9ull + (h << 6) + (h >> 2);
        h ^= h >> 33;
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 33;
        return (uint32_t)h;
    }

    // v' = clamp(v + a*0.001 + cos(3x)*-0.0025, -0.07, 0.07); x' = x + v'; inelastic wall at -1.2
    __device__ __forceinline__ static void gen(const Params& p, const State& s, int ai, Stream&, State& sp, double& r) {
        const double push = ai == 0 ? -0.001 : (ai == 1 ? 0.0 : 0.001);  // a * 0.001 for a in (-1., 0., 1.): exact
#if MCTSB200_MC_COS_NB
        double sn_, cs_;
        det_sincos(DM_MUL(3.0, s.x), &sn_, &cs_);  // experiment: no quadrant branch (both kernels always)
        double v_ = DM_ADD(DM_ADD(s.v, push), DM_MUL(cs_, -0.0025));
#else
        double v_ = DM_ADD(DM_ADD(s.v, push), DM_MUL(det_cos(DM_MUL(3.0, s.x)), -0.0025));
#endif
        v_ = v_ > 0.07 ? 0.07 : (v_ < -0.07 ? -0.07 : v_);  // clamp (two compares: FP64 min/max are multi-instruction sequences)
        double x_ = DM_ADD(s.x, v_);
        if (x_ < -1.2) {
            x_ = -1.2;
            v_ = 0.0;
        }
        sp = State{x_, v_};
        r = is_terminal(p, sp) ? p.jackpot : p.cost;
    }

    __device__ __forceinline__ static int policy_action(const Params&, int policy, const int* pp, const State& s, Stream& rng) {
        switch (policy) {
            case MCTSB200_POLICY_RANDOM: return rng.uniform_int(kActions);
            case MCTSB200_POLICY_CONSTANT: return pp[0];
            case MCTSB200_POLICY_MC_ENERGY: return s.v < 0.0 ? 0 : 2;
            default: return 0;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// GaussianBox: continuous states AND continuous actions (include/mctsb200.h: mctsb200_gaussian_box_params; the MDP of the
// reference's test/discussion_514.jl with an optional drift and quadratic cost). Only DPW plans for it.
// ---------------------------------------------------------------------------------------------
struct GaussianBoxDev {
    static constexpr int kActions = 1;   // (no finite action set: Row / Aux keep one unused slot)
    static constexpr bool kDense = false;
    static constexpr int kMaxBranch = 0;  // continuous noise: transitions[sanode] is kept as the push-ordered vector
    static constexpr int kMdpId = MCTSB200_MDP_GAUSSIAN_BOX;
    static constexpr bool kCustom = false;
    static constexpr bool kFast = false;
    static constexpr int kActionDoubles = MCTSB200_BOX_ACTION_DOUBLES;
    static constexpr int kStateDoubles = MCTSB200_BOX_STATE_DOUBLES;

    struct AbiState { double v[kStateDoubles]; };
    typedef AbiState State;
    typedef AbiState Key;
    struct Action { double v[kActionDoubles]; };
    struct Params {
        double discount;
        double a_lo[kActionDoubles], a_hi[kActionDoubles];
        double sd[3];  // sqrt(var), rounded once on the host (the oracle does the same)
        double drift, cost;
    };

    __device__ __forceinline__ static State load(const Params&, const AbiState& a) { return a; }
    __host__ __device__ __forceinline__ static AbiState store(const Params&, const State& s) { return s; }
    __device__ __forceinline__ static bool is_terminal(const Params&, const State&) { return false; }
    __device__ __forceinline__ static double discount(const Params& p) { return p.discount; }
    __device__ __forceinline__ static int dense_index(const Params&, const State&) { return -1; }
    __device__ __forceinline__ static Key key_of(const Params&, const State& s) { return s; }
    __host__ __device__ __forceinline__ static State state_of(const Params&, const Key& k) { return k; }
    __device__ __forceinline__ static bool key_equal(const Key& a, const Key& b) {
        bool eq = true;
#pragma unroll
        for (int i = 0; i < kStateDoubles; ++i) eq = eq && __double_as_longlong(a.v[i]) == __double_as_longlong(b.v[i]);
        return eq;
    }
    __device__ __forceinline__ static uint32_t hash(const State& s) {
        uint64_t h = 0x7F4A7C159E3779B9ull;
#pragma unroll
        for (int i = 0; i < kStateDoubles; ++i) {
            h ^= (uint64_t)__double_as_longlong(s.v[i]) * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            h ^= h >> 29;
        }
        h ^= h >> 33;
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 33;
        return (uint32_t)h;
    }
    // rand(rng, actions(mdp, s)) on a Box: one uniform draw per coordinate (RandomActionGenerator, src/action_gen.jl:11-13)
    __device__ __forceinline__ static void next_action(const Params& p, const State&, Stream& rng, Action& a) {
#pragma unroll
        for (int i = 0; i < kActionDoubles; ++i) a.v[i] = DM_ADD(p.a_lo[i], DM_MUL(DM_SUB(p.a_hi[i], p.a_lo[i]), rng.uniform01()));
    }
    // the rollout policy: only RandomPolicy (MdpAccess<GaussianBoxDev>::policy_ok)
    __device__ __forceinline__ static void policy_action(const Params& p, int, const int*, const State& s, Stream& rng, Action& a) { next_action(p, s, rng, a); }
    // sp = s + drift * a + sd * z, z ~ N(0, I): two Box-Muller pairs from four words (z0, z1 from the first two, z2 from the others)
    __device__ __forceinline__ static void gen(const Params& p, const State& s, const Action& a, Stream& rng, State& sp, double& r) {
        const double two_pi = 6.283185307179586;
        double z[3];
        {
            const double u1 = DM_MUL(DM_ADD((double)rng.next(), 1.0), 1.0 / 4294967296.0), u2 = rng.uniform01();
            const double rad = DM_SQRT(DM_MUL(-2.0, det_log(u1))), ang = DM_MUL(two_pi, u2);
            double sn, cs;
            det_sincos(ang, &sn, &cs);  // (= det_sin(ang), det_cos(ang) with one reduction)
            z[0] = DM_MUL(rad, cs);
            z[1] = DM_MUL(rad, sn);
        }
        {
            const double u1 = DM_MUL(DM_ADD((double)rng.next(), 1.0), 1.0 / 4294967296.0), u2 = rng.uniform01();
            z[2] = DM_MUL(DM_SQRT(DM_MUL(-2.0, det_log(u1))), det_cos(DM_MUL(two_pi, u2)));
        }
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double x = s.v[i];
            if (i < kActionDoubles) x = DM_ADD(x, DM_MUL(p.drift, a.v[i]));
            x = DM_ADD(x, DM_MUL(p.sd[i], z[i]));
            sp.v[i] = x;
            ss = DM_ADD(ss, DM_MUL(x, x));
        }
        sp.v[3] = s.v[3];
        r = DM_SUB(0.0, DM_MUL(p.cost, ss));
    }
    // The same on a tile of G >= 2 lanes (which all run the same stream: search.cuh). The two Box-Muller pairs are the same
    // functions on different arguments, so even lanes take the first pair and odd lanes the second: one det_log, one square root, one
    // det_cos and one det_sin pass instead of two / two / two / one -- the same rounded operations on the same operands, hence the
    // same bits as gen(); the three deviates come back by shuffle.
    static constexpr bool kCoopGen = true;
    template <class TileT>
    __device__ __forceinline__ static void gen_coop(const TileT& tile, const Params& p, const State& s, const Action& a, Stream& rng, State& sp,
                                                    double& r) {
        const double two_pi = 6.283185307179586;
        const double u1a = DM_MUL(DM_ADD((double)rng.next(), 1.0), 1.0 / 4294967296.0), u2a = rng.uniform01();
        const double u1b = DM_MUL(DM_ADD((double)rng.next(), 1.0), 1.0 / 4294967296.0), u2b = rng.uniform01();
        const bool second = (tile.lane & 1) != 0;
        const double rad = DM_SQRT(DM_MUL(-2.0, det_log(second ? u1b : u1a))), ang = DM_MUL(two_pi, second ? u2b : u2a);
        double sn, cs;
        det_sincos(ang, &sn, &cs);               // (no quadrant branch for the lanes to split over)
        const double zc = DM_MUL(rad, cs);       // z0 on even lanes, z2 on odd ones
        const double zs = DM_MUL(rad, sn);       // z1 on even lanes
        const double z[3] = {tile.bcast(zc, 0), tile.bcast(zs, 0), tile.bcast(zc, 1)};
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double x = s.v[i];
            if (i < kActionDoubles) x = DM_ADD(x, DM_MUL(p.drift, a.v[i]));
            x = DM_ADD(x, DM_MUL(p.sd[i], z[i]));
            sp.v[i] = x;
            ss = DM_ADD(ss, DM_MUL(x, x));
        }
        sp.v[3] = s.v[3];
        r = DM_SUB(0.0, DM_MUL(p.cost, ss));
    }
};

}  // namespace mctsb200
