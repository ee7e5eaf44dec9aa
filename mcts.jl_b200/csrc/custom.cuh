// custom.cuh -- the trait struct behind an MDP plugin registered at run time (mctsb200_mdp_register, include/mctsb200.h).
//
// SURVEY.md 8(b) "MDP plugin (device side)" / 8(f) rank 3. A user plugin is CUDA source text that defines one struct:
//
//     struct MyMdp {
//         struct Params { double discount; /* ... any POD, <= 256 bytes in all, `discount` first ... */ };
//         // s / sp: the state, kStateDoubles float64 values (statetype(mdp) as the host passes it)
//         __device__ static bool is_terminal(const Params& p, const double* s);                       // isterminal(mdp, s)
//         __device__ static void gen(const Params& p, const double* s, int a, mctsb200::Stream& rng,
//                                    double* sp, double& r);                                         // @gen(:sp,:r)(mdp, s, a, rng)
//         // rollout policies beyond MCTSB200_POLICY_RANDOM / _CONSTANT (ids >= MCTSB200_POLICY_PLUGIN):
//         __device__ static int policy_action(const Params& p, int policy, const int* pp, const double* s, mctsb200::Stream& rng);
//     };
//
// Params is an opaque POD to the library: it may carry pointers to tables the caller keeps in device memory (a tabular model, a
// value table; tests/test_gpu_parity.py::test_plugin_params_may_carry_device_pointers).
// `a` is the 0-based index into actions(mdp); random numbers come from `rng.next()` (32-bit words of the tree's Philox
// stream), `rng.uniform_int(n)`, `rng.uniform01()`. The source sees everything search.cuh sees (DM_ADD / DM_MUL /
// det_log / det_cos / det_sin / det_sincos for reproducible arithmetic). The library compiles search_kernel<CustomDev<SD, A>, KIND, G> around
// it with NVRTC for sm_100a -- the same tile kernel the built-in MountainCar runs -- and launches the result; nothing of
// a plugin ever runs on the host (north star: no CPU fallback).
//
// Continuous action spaces (mctsb200_mdp_register_continuous; actions(mdp) = a Box, test/discussion_514.jl): the struct defines
//
//         __device__ static bool is_terminal(const Params& p, const double* s);
//         // next_action(gen, mdp, s, snode) (src/dpw.jl:98): the action DPW's action widening proposes, kActionDoubles float64 values
//         __device__ static void next_action(const Params& p, const double* s, mctsb200::Stream& rng, double* a);
//         __device__ static void gen(const Params& p, const double* s, const double* a, mctsb200::Stream& rng, double* sp, double& r);
//         // rollout policies with ids >= MCTSB200_POLICY_PLUGIN (MCTSB200_POLICY_RANDOM is next_action itself):
//         __device__ static void policy_action(const Params& p, int policy, const int* pp, const double* s, mctsb200::Stream& rng, double* a);
//
// and the library compiles search_kernel<CustomBoxDev<SD, AD>, DPW, G> around it (csrc/search.cuh: dpw_box_iteration).
//
// The host half of the library is compiled ahead of time for a fixed set of shapes (SD state doubles x A actions, see
// MCTSB200_FOR_PLUGIN in capi.cu); the layout of KernelArgs<CustomDev<SD, A>> does not depend on the plugin.
#pragma once
#include "common.cuh"

#ifndef MCTSB200_CUSTOM_MAX_BRANCH
#define MCTSB200_CUSTOM_MAX_BRANCH 1  // (only the run-time compiled kernels read it; the host sizes tables from the registry)
#endif

namespace mctsb200 {

constexpr int kCustomParamBytes = 256;

// a plugin's parameters as the library sees them
struct __align__(16) CustomParams {
    double discount;  // the first member of every plugin's Params
    unsigned char rest[kCustomParamBytes - sizeof(double)];
};

#if defined(__CUDACC_RTC__)
// isequal on Float64 tuples == bitwise equality (as for the built-in MountainCar)
template <int SD>
__device__ __forceinline__ bool custom_key_equal(const double* a, const double* b) {
    bool eq = true;
#pragma unroll
    for (int i = 0; i < SD; ++i) eq = eq && __double_as_longlong(a[i]) == __double_as_longlong(b[i]);
    return eq;
}
template <int SD>
__device__ __forceinline__ uint32_t custom_hash(const double* s) {
    uint64_t h = 0x7F4A7C159E3779B9ull;
#pragma unroll
    for (int i = 0; i < SD; ++i) {
        h ^= (uint64_t)__double_as_longlong(s[i]) * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= h >> 29;
    }
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    return (uint32_t)h;
}
#endif

// the plugin struct of a shape; specialised by the generated NVRTC program, never by the ahead-of-time build
template <int SD, int A>
struct CustomPluginOf;

template <int SD, int A>
struct CustomDev {
    static constexpr int kActions = A;
    static constexpr bool kDense = false;
    static constexpr int kMaxBranch = MCTSB200_CUSTOM_MAX_BRANCH;
    static constexpr bool kFast = false;
    static constexpr bool kCustom = true;
    static constexpr int kStateDoubles = SD;
    static constexpr int kActionDoubles = 0;

    struct AbiState { double v[SD]; };
    typedef AbiState State;
    typedef AbiState Key;
    typedef CustomParams Params;

    __host__ __device__ __forceinline__ static State load(const Params&, const AbiState& a) { return a; }
    __host__ __device__ __forceinline__ static AbiState store(const Params&, const State& s) { return s; }
    __host__ __device__ __forceinline__ static State state_of(const Params&, const Key& k) { return k; }
    __host__ __device__ __forceinline__ static Key key_of(const Params&, const State& s) { return s; }

#if defined(__CUDACC_RTC__)
    using P = typename CustomPluginOf<SD, A>::type;
    using PParams = typename P::Params;
    static_assert(sizeof(PParams) <= sizeof(Params), "a plugin's Params must fit in 256 bytes");
    __device__ __forceinline__ static const PParams& pp_of(const Params& p) { return *reinterpret_cast<const PParams*>(&p); }

    __device__ __forceinline__ static bool is_terminal(const Params& p, const State& s) { return P::is_terminal(pp_of(p), s.v); }
    __device__ __forceinline__ static double discount(const Params& p) { return p.discount; }
    __device__ __forceinline__ static int dense_index(const Params&, const State&) { return -1; }
    __device__ __forceinline__ static bool key_equal(const Key& a, const Key& b) { return custom_key_equal<SD>(a.v, b.v); }
    __device__ __forceinline__ static uint32_t hash(const State& s) { return custom_hash<SD>(s.v); }
    __device__ __forceinline__ static void gen(const Params& p, const State& s, int a, Stream& rng, State& sp, double& r) {
        State out;
        P::gen(pp_of(p), s.v, a, rng, out.v, r);
        sp = out;
    }
    __device__ __forceinline__ static int policy_action(const Params& p, int policy, const int* pp, const State& s, Stream& rng) {
        if (policy == MCTSB200_POLICY_RANDOM) return rng.uniform_int(kActions);
        if (policy == MCTSB200_POLICY_CONSTANT) return pp[0];
        return P::policy_action(pp_of(p), policy, pp, s.v, rng);
    }
#endif
};

// ---- continuous action spaces: the action is a vector of AD float64 values ---------------------------------------------------
template <int SD, int AD>
struct CustomBoxPluginOf;

template <int SD, int AD>
struct CustomBoxDev {
    static constexpr int kActions = 1;    // (no finite action set: Row / Aux keep one unused slot, as GaussianBoxDev)
    static constexpr bool kDense = false;
    static constexpr int kMaxBranch = 0;  // transitions[sanode] is kept as the push-ordered vector
    static constexpr bool kFast = false;
    static constexpr bool kCustom = true;
    static constexpr int kStateDoubles = SD;
    static constexpr int kActionDoubles = AD;
    static constexpr bool kCoopGen = false;  // (a plugin's gen runs on every lane of a tile)
    static_assert(AD >= 1 && AD <= 4, "an action node's label holds up to four doubles (BoxANode)");

    struct AbiState { double v[SD]; };
    typedef AbiState State;
    typedef AbiState Key;
    struct Action { double v[AD]; };
    typedef CustomParams Params;

    __host__ __device__ __forceinline__ static State load(const Params&, const AbiState& a) { return a; }
    __host__ __device__ __forceinline__ static AbiState store(const Params&, const State& s) { return s; }
    __host__ __device__ __forceinline__ static State state_of(const Params&, const Key& k) { return k; }
    __host__ __device__ __forceinline__ static Key key_of(const Params&, const State& s) { return s; }

#if defined(__CUDACC_RTC__)
    using P = typename CustomBoxPluginOf<SD, AD>::type;
    using PParams = typename P::Params;
    static_assert(sizeof(PParams) <= sizeof(Params), "a plugin's Params must fit in 256 bytes");
    __device__ __forceinline__ static const PParams& pp_of(const Params& p) { return *reinterpret_cast<const PParams*>(&p); }

    __device__ __forceinline__ static bool is_terminal(const Params& p, const State& s) { return P::is_terminal(pp_of(p), s.v); }
    __device__ __forceinline__ static double discount(const Params& p) { return p.discount; }
    __device__ __forceinline__ static int dense_index(const Params&, const State&) { return -1; }
    __device__ __forceinline__ static bool key_equal(const Key& a, const Key& b) { return custom_key_equal<SD>(a.v, b.v); }
    __device__ __forceinline__ static uint32_t hash(const State& s) { return custom_hash<SD>(s.v); }
    __device__ __forceinline__ static void next_action(const Params& p, const State& s, Stream& rng, Action& a) {
        Action out;
        P::next_action(pp_of(p), s.v, rng, out.v);
        a = out;
    }
    __device__ __forceinline__ static void gen(const Params& p, const State& s, const Action& a, Stream& rng, State& sp, double& r) {
        State out;
        P::gen(pp_of(p), s.v, a.v, rng, out.v, r);
        sp = out;
    }
    // action(policy, s) of the rollout policy: RandomPolicy = rand(rng, actions(mdp, s)) = next_action's draw
    __device__ __forceinline__ static void policy_action(const Params& p, int policy, const int* pp, const State& s, Stream& rng, Action& a) {
        Action out;
        if (policy == MCTSB200_POLICY_RANDOM) P::next_action(pp_of(p), s.v, rng, out.v);
        else P::policy_action(pp_of(p), policy, pp, s.v, rng, out.v);
        a = out;
    }
#endif
};

}  // namespace mctsb200
