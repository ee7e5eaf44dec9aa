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

namespace mctsb200 {

// ---------------------------------------------------------------------------------------------
// SimpleGridWorld
// ---------------------------------------------------------------------------------------------
struct GridWorldDev {
    static constexpr int kActions = 4;
    static constexpr bool kDense = true;
    static constexpr int kMaxBranch = 5;
    static constexpr int kMdpId = MCTSB200_MDP_GRIDWORLD;

    struct AbiState { long long x, y; };
    struct State { int x, y; };
    typedef uint16_t Key;  // dense cell index; nx*ny = the terminal state (-1,-1)

    struct Params {
        int nx, ny;
        double tprob;    // probability of the intended move
        double p_other;  // (1.0 - tprob) / 3
        double discount;
        const double* cell_reward;       // [nx*ny + 1]   get(rewards, s, 0.0)
        const uint8_t* cell_term_from;   // [nx*ny + 1]   s in terminate_from
    };

    __device__ __forceinline__ static State load(const AbiState& a) { return State{(int)a.x, (int)a.y}; }
    __host__ __device__ __forceinline__ static AbiState store(const State& s) { return AbiState{s.x, s.y}; }
    __device__ __forceinline__ static bool is_terminal(const Params&, const State& s) { return s.x < 0 || s.y < 0; }
    __device__ __forceinline__ static double discount(const Params& p) { return p.discount; }
    __device__ __forceinline__ static int dense_count(const Params& p) { return p.nx * p.ny + 1; }
    __host__ __device__ __forceinline__ static int dense_index(const Params& p, const State& s) {
        return (s.x < 0 || s.y < 0) ? p.nx * p.ny : (s.x - 1) + (s.y - 1) * p.nx;
    }
    __device__ __forceinline__ static Key key_of(const Params& p, const State& s) { return (Key)dense_index(p, s); }
    __host__ __device__ __forceinline__ static State state_of(const Params& p, Key k) {
        if ((int)k == p.nx * p.ny) return State{-1, -1};
        return State{(int)k % p.nx + 1, (int)k / p.nx + 1};
    }
    __device__ __forceinline__ static bool key_equal(Key a, Key b) { return a == b; }

    // transition(mdp, s, a) + rand(rng, SparseCat) + reward(mdp, s, a); one uniform draw unless the
    // source is a terminate_from cell or terminal (Deterministic next state, no draw).
    __device__ __forceinline__ static void gen(const Params& p, const State& s, int a, Stream& rng, State& sp, double& r) {
        const int cell = dense_index(p, s);
        r = __ldg(p.cell_reward + cell);
        if (__ldg(p.cell_term_from + cell) || is_terminal(p, s)) {
            sp = State{-1, -1};
            return;
        }
        // destination order [stay, up, down, left, right]; out-of-bounds mass folds into "stay"
        const bool ob_up = s.y + 1 > p.ny, ob_down = s.y - 1 < 1, ob_left = s.x - 1 < 1, ob_right = s.x + 1 > p.nx;
        const double pu = a == 0 ? p.tprob : p.p_other, pd = a == 1 ? p.tprob : p.p_other;
        const double pl = a == 2 ? p.tprob : p.p_other, pr = a == 3 ? p.tprob : p.p_other;
        double p0 = 0.0;
        if (ob_up) p0 = DM_ADD(p0, pu);
        if (ob_down) p0 = DM_ADD(p0, pd);
        if (ob_left) p0 = DM_ADD(p0, pl);
        if (ob_right) p0 = DM_ADD(p0, pr);
        const double p1 = ob_up ? 0.0 : pu, p2 = ob_down ? 0.0 : pd, p3 = ob_left ? 0.0 : pl, p4 = ob_right ? 0.0 : pr;
        const double c1 = DM_ADD(p0, p1), c2 = DM_ADD(c1, p2), c3 = DM_ADD(c2, p3), c4 = DM_ADD(c3, p4);
        const double u = DM_MUL(c4, rng.uniform01());
        sp = s;
        if (u < p0) {
        } else if (u < c1) {
            sp.y = s.y + 1;
        } else if (u < c2) {
            sp.y = s.y - 1;
        } else if (u < c3) {
            sp.x = s.x - 1;
        } else {
            sp.x = s.x + 1;  // u < c4 always holds for u01 < 1
        }
        // an out-of-bounds destination carries probability 0 and is never selected, except through the
        // final else when its mass is 0 and rounding pushed u to c4; the reference errors there (u < c4
        // fails only if rand() == 1, which [0,1) excludes).
    }

    __device__ __forceinline__ static int policy_action(const Params&, int policy, const int* pp, const State& s, Stream& rng) {
        switch (policy) {
            case MCTSB200_POLICY_RANDOM: return rng.uniform_int(kActions);
            case MCTSB200_POLICY_CONSTANT: return pp[0];
            case MCTSB200_POLICY_GW_SEEK:
                if (pp[0] > s.x) return 3;
                else if (pp[0] < s.x) return 2;
                else if (pp[1] > s.y) return 0;
                else return 1;
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

    struct AbiState { double x, v; };
    typedef AbiState State;
    struct Key { double x, v; };

    struct Params {
        double discount, cost, jackpot;
    };

    __device__ __forceinline__ static State load(const AbiState& a) { return a; }
    __host__ __device__ __forceinline__ static AbiState store(const State& s) { return s; }
    __device__ __forceinline__ static bool is_terminal(const Params&, const State& s) { return s.x >= 0.5; }
    __device__ __forceinline__ static double discount(const Params& p) { return p.discount; }
    __device__ __forceinline__ static int dense_count(const Params&) { return 0; }
    __host__ __device__ __forceinline__ static int dense_index(const Params&, const State&) { return -1; }
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
        const double a = (double)(ai - 1);
        double v_ = DM_ADD(DM_ADD(s.v, DM_MUL(a, 0.001)), DM_MUL(det_cos(DM_MUL(3.0, s.x)), -0.0025));
        v_ = fmax(fmin(0.07, v_), -0.07);
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

}  // namespace mctsb200
