// mcts_oracle.cpp -- CPU restatement of MCTS.jl's search loop.  TEST INFRASTRUCTURE ONLY.
//
// Nothing in the product path (libmctsb200.so, mcts.jl_b200/) may link, import or execute this file.
// It exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// leg can check and time the CUDA path against an independent statement of the reference algorithm.
//
// It deliberately keeps the reference's own shape -- recursive simulate, Vector-of-Vectors trees,
// Dict/Set lookups, 1-based node ids -- so each function can be read against the Julia it follows
// (paths relative to /root/reference):
//   VanillaPlanner::build_tree        src/vanilla.jl:209-237
//   VanillaPlanner::simulate          src/vanilla.jl:240-276
//   VanillaPlanner::insert_node       src/vanilla.jl:292-307
//   VanillaPlanner::best_sanode_UCB   src/vanilla.jl:354-386
//   VanillaPlanner::best_sanode_Q     src/vanilla.jl:339-349
//   DPWPlanner::action_info           src/dpw.jl:18-74
//   DPWPlanner::simulate              src/dpw.jl:80-154
//   DPWPlanner::best_sanode           src/dpw.jl:163-173
//   DPWPlanner::best_sanode_UCB       src/dpw.jl:179-199
//   DPWTree / insert_state_node / insert_action_node   src/dpw_types.jl:123-186
//   Estimator::estimate_value         src/domain_knowledge.jl:10,65-73
//   init_Q / init_N                   src/domain_knowledge.jl:80-91
//   next_action                       src/action_gen.jl:11-13
//
// Third-party arithmetic that is NOT under /root/reference (Project.toml:6-27 lists the packages,
// no Manifest.toml pins them) is restated from the published sources of
//   POMDPTools.jl  (compat "0.1, 1"):  RolloutSimulator loop, RandomPolicy, FunctionPolicy, SparseCat sampling
//   POMDPModels.jl (compat "0.4"):     SimpleGridWorld, MountainCar
// and anchored on the reference's call sites (src/domain_knowledge.jl:71-72, src/vanilla.jl:258,
// src/dpw.jl:122) and stored notebook outputs (tests/test_oracle_golden.py).
//
// Parity status: the search-loop arithmetic is pinned by the golden vectors of SURVEY.md 8(c)
// (notebook Q/N values, DAG/self-loop invariants, root id, VI action table).  No reference test pins
// a full seeded search bit-for-bit (they all depend on Julia's MersenneTwister), and Julia is not
// available in the build image, so stochastic parity with Julia is statistical; MountainCar parity is
// UNPINNED by the reference (the model is never imported by it).
//
// Deliberate, documented deviations from a literal Julia run:
//   * RNG: counter-based Philox4x32-10 streams (include/mctsb200.h "Random-number contract").
//   * log/cos: include/mctsb200_detmath.h (Julia's are not correctly rounded either; see that header).
//     Build with -DORACLE_LIBM to use glibc's instead (used by a test that shows actions agree).
//   * transitions[sanode] sampling has two modes: literal (vector indexed by push order, the default
//     restatement of src/dpw.jl:131,140) and multiset (distinct (spnode,r) records with counts,
//     sampled by cumulative count in first-occurrence order) -- distribution-identical;
//     the device uses the multiset form.
#include <algorithm>
#include <any>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <array>
#include <utility>
#include <vector>

#include "../include/mctsb200.h"
#include "../include/mctsb200_detmath.h"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) and the stream contract of include/mctsb200.h
// ---------------------------------------------------------------------------------------------
static inline void philox4x32_10(const uint32_t key_in[2], const uint32_t ctr_in[4], uint32_t out[4]) {
    uint32_t k0 = key_in[0], k1 = key_in[1];
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Stream {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int have = 0;  // unread words left in buf
    Stream(uint64_t seed, uint64_t tree, uint32_t iteration, uint32_t stream_id) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = 0;
        ctr[1] = iteration;
        ctr[2] = (uint32_t)tree;
        ctr[3] = (uint32_t)(((tree >> 32) & 0xffffu) << 16) | (stream_id & 0xffffu);
    }
    uint32_t next() {
        if (have == 0) {
            philox4x32_10(key, ctr, buf);
            ctr[0] += 1;
            have = 4;
        }
        const uint32_t w = buf[4 - have];
        --have;
        return w;
    }
    int uniform_int(int n) { return (int)(((uint64_t)next() * (uint64_t)n) >> 32); }
    double uniform01() { return (double)next() * (1.0 / 4294967296.0); }
};

static inline double o_log(double x) {
#ifdef ORACLE_LIBM
    return std::log(x);
#else
    return det_log(x);
#endif
}
static inline double o_cos(double x) {
#ifdef ORACLE_LIBM
    return std::cos(x);
#else
    return det_cos(x);
#endif
}

// ---------------------------------------------------------------------------------------------
// Models (POMDPModels.jl restatements)
// ---------------------------------------------------------------------------------------------
struct GWPos {
    int64_t x, y;
    bool operator==(const GWPos& o) const { return x == o.x && y == o.y; }
    bool operator<(const GWPos& o) const { return x != o.x ? x < o.x : y < o.y; }
};
struct GWPosHash {
    size_t operator()(const GWPos& s) const { return std::hash<int64_t>()(s.x * 1000003 + s.y); }
};

// SimpleGridWorld: actions (:up,:down,:left,:right); dir up(0,1) down(0,-1) left(-1,0) right(1,0).
struct GridWorld {
    static constexpr int kMdpId = MCTSB200_MDP_GRIDWORLD;
    static constexpr bool kLiteralTransitions = false;  // (the device keeps transitions[sanode] as records with multiplicities)
    using State = GWPos;
    // actions are indices into actions(mdp) (a continuous-action model names its own Action type below)
    typedef int32_t Action;
    static constexpr int kActionDoubles = 0;
    Action next_action(const State& s, Stream& rng) const { return rng.uniform_int(n_actions(s)); }  // rand(rng, actions(mdp, s))
    static int32_t action_key(Action a) { return a; }
    static constexpr int kActions = 4;
    int nx = 10, ny = 10;
    double tprob = 0.7, gamma = 0.95;
    std::map<GWPos, double> rewards;
    std::set<GWPos> terminate_from;

    explicit GridWorld(const mctsb200_gridworld_params& p) : nx(p.nx), ny(p.ny), tprob(p.tprob), gamma(p.discount) {
        for (int i = 0; i < p.n_rewards; ++i) rewards[GWPos{p.reward_x[i], p.reward_y[i]}] = p.reward_v[i];
        for (int i = 0; i < p.n_terminate; ++i) terminate_from.insert(GWPos{p.term_x[i], p.term_y[i]});
    }
    double discount() const { return gamma; }
    int n_actions(const State&) const { return kActions; }
    int max_actions() const { return kActions; }
    bool isterminal(const State& s) const { return s.x < 0 || s.y < 0; }  // any(x->x<0, s)
    bool inbounds(const State& s) const { return 1 <= s.x && s.x <= nx && 1 <= s.y && s.y <= ny; }
    double reward(const State& s) const {  // reward(mdp, s, a) = get(mdp.rewards, s, 0.0)
        auto it = rewards.find(s);
        return it == rewards.end() ? 0.0 : it->second;
    }
    // dense index used only for the optional init/value tables: cells row-major in x, terminal last
    int64_t state_index(const State& s) const { return isterminal(s) ? (int64_t)nx * ny : (s.x - 1) + (s.y - 1) * (int64_t)nx; }
    int64_t n_states() const { return (int64_t)nx * ny + 1; }

    // transition(mdp, s, a) followed by rand(rng, dist); reward from the source state.
    void gen(const State& s, int a, Stream& rng, State& sp, double& r) const {
        r = reward(s);
        if (terminate_from.count(s) || isterminal(s)) {  // Deterministic(GWPos(-1,-1)): no draw
            sp = GWPos{-1, -1};
            return;
        }
        static const int dx[4] = {0, 0, -1, 1}, dy[4] = {1, -1, 0, 0};
        State destinations[5];
        double probs[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        destinations[0] = s;
        for (int i = 0; i < 4; ++i) {
            double prob;
            if (i == a) prob = tprob;
            else prob = (1.0 - tprob) / (double)(kActions - 1);
            State dest{s.x + dx[i], s.y + dy[i]};
            destinations[i + 1] = dest;
            if (!inbounds(dest)) {  // hit an edge and come back
                probs[0] += prob;
                destinations[i + 1] = GWPos{-1, -1};
            } else {
                probs[i + 1] += prob;
            }
        }
        // rand(rng, ::SparseCat): r = sum(probs)*rand(rng); first index with r < running total
        double total = 0.0;
        for (int i = 0; i < 5; ++i) total += probs[i];
        const double u = total * rng.uniform01();
        double tot = 0.0;
        for (int i = 0; i < 5; ++i) {
            tot += probs[i];
            if (u < tot) {
                sp = destinations[i];
                return;
            }
        }
        sp = destinations[4];  // unreachable for u < 1 (Julia raises an error here)
    }
};

struct MCState {
    double x, v;
    // isequal semantics on Float64 pairs: bitwise (so -0.0 != 0.0, NaN == NaN)
    bool operator==(const MCState& o) const { return dm_bits(x) == dm_bits(o.x) && dm_bits(v) == dm_bits(o.v); }
};
struct MCStateHash {
    size_t operator()(const MCState& s) const { return std::hash<uint64_t>()(dm_bits(s.x) * 0x9E3779B97F4A7C15ull ^ dm_bits(s.v)); }
};

// MountainCar: actions [-1., 0., 1.]; gen is deterministic.
struct MountainCar {
    static constexpr int kMdpId = MCTSB200_MDP_MOUNTAINCAR;
    static constexpr bool kLiteralTransitions = false;
    using State = MCState;
    // actions are indices into actions(mdp) (a continuous-action model names its own Action type below)
    typedef int32_t Action;
    static constexpr int kActionDoubles = 0;
    Action next_action(const State& s, Stream& rng) const { return rng.uniform_int(n_actions(s)); }  // rand(rng, actions(mdp, s))
    static int32_t action_key(Action a) { return a; }
    static constexpr int kActions = 3;
    double gamma, cost, jackpot;
    explicit MountainCar(const mctsb200_mountaincar_params& p) : gamma(p.discount), cost(p.cost), jackpot(p.jackpot) {}
    double discount() const { return gamma; }
    int n_actions(const State&) const { return kActions; }
    int max_actions() const { return kActions; }
    bool isterminal(const State& s) const { return s.x >= 0.5; }
    int64_t state_index(const State&) const { return -1; }
    int64_t n_states() const { return 0; }
    void gen(const State& s, int ai, Stream&, State& sp, double& r) const {
        const double a = (double)(ai - 1);
        // v_ = v + a*0.001 + cos(3*x)*-0.0025
        double v_ = (s.v + a * 0.001) + o_cos(3.0 * s.x) * -0.0025;
        v_ = std::max(std::min(0.07, v_), -0.07);
        double x_ = s.x + v_;
        if (x_ < -1.2) {  // inelastic boundary
            x_ = -1.2;
            v_ = 0.0;
        }
        sp = MCState{x_, v_};
        r = isterminal(sp) ? jackpot : cost;
    }
};

// MountainCar with uniform velocity noise: the model of mcts.jl_b200/plugins/noisy_mountaincar.cuh, a plugin registered at run
// time through mctsb200_mdp_register (no counterpart in the reference; it checks that the tile kernel compiled around a plugin
// follows src/dpw.jl:80-154 for a continuous-state MDP with stochastic transitions). One random word per gen call.
struct NoisyMountainCarParams {
    double discount, cost, jackpot, noise;
    double push[4];     // velocity increment of action i (POMDPModels' MountainCar: a * 0.001 for a in (-1., 0., 1.))
    int32_t n_actions;  // 2..4: the plugin is registered once per action-set size (the shape is a registration argument)
    int32_t pad;
};
constexpr int kOracleMdpNoisyMountainCar = 1000;  // oracle-side id (the library's plugin ids are assigned at registration)
struct NoisyMountainCar {
    static constexpr int kMdpId = kOracleMdpNoisyMountainCar;
    static constexpr bool kLiteralTransitions = true;  // unbounded branching: the device keeps the push-ordered vector itself
    using State = MCState;
    // actions are indices into actions(mdp) (a continuous-action model names its own Action type below)
    typedef int32_t Action;
    static constexpr int kActionDoubles = 0;
    Action next_action(const State& s, Stream& rng) const { return rng.uniform_int(n_actions(s)); }  // rand(rng, actions(mdp, s))
    static int32_t action_key(Action a) { return a; }
    static constexpr int kActions = 3;
    double gamma, cost, jackpot, noise, push[4];
    int n_act;
    explicit NoisyMountainCar(const NoisyMountainCarParams& p) : gamma(p.discount), cost(p.cost), jackpot(p.jackpot), noise(p.noise), n_act(p.n_actions) {
        for (int i = 0; i < 4; ++i) push[i] = p.push[i];
    }
    double discount() const { return gamma; }
    int n_actions(const State&) const { return n_act; }
    int max_actions() const { return n_act; }
    bool isterminal(const State& s) const { return s.x >= 0.5; }
    int64_t state_index(const State&) const { return -1; }
    int64_t n_states() const { return 0; }
    void gen(const State& s, int ai, Stream& rng, State& sp, double& r) const {
        const double u = rng.uniform01();
        double v_ = (s.v + push[ai]) + o_cos(3.0 * s.x) * -0.0025;
        v_ = v_ + (u - 0.5) * noise;
        v_ = std::max(std::min(0.07, v_), -0.07);
        double x_ = s.x + v_;
        if (x_ < -1.2) {
            x_ = -1.2;
            v_ = 0.0;
        }
        sp = MCState{x_, v_};
        r = isterminal(sp) ? jackpot : cost;
    }
};

// POMDPModels.jl's InvertedPendulum as shipped in mcts.jl_b200/plugins/inverted_pendulum.cuh (the model of
// notebooks/Test_Visualization.ipynb cell 4), registered at run time like NoisyMountainCar. Euler step, uniform force noise.
struct InvertedPendulumParams {
    double discount, g, m, l, M, alpha, dt, cost;
    double force[3];
    double noise;
};
constexpr int kOracleMdpInvertedPendulum = 1001;
struct InvertedPendulum {
    static constexpr int kMdpId = kOracleMdpInvertedPendulum;
    static constexpr bool kLiteralTransitions = true;
    using State = MCState;  // (theta, omega)
    typedef int32_t Action;
    static constexpr int kActionDoubles = 0;
    Action next_action(const State& s, Stream& rng) const { return rng.uniform_int(n_actions(s)); }
    static int32_t action_key(Action a) { return a; }
    static constexpr int kActions = 3;
    InvertedPendulumParams p;
    explicit InvertedPendulum(const InvertedPendulumParams& q) : p(q) {}
    double discount() const { return p.discount; }
    int n_actions(const State&) const { return kActions; }
    int max_actions() const { return kActions; }
    bool isterminal(const State& s) const { return std::fabs(s.x) > 1.5707963267948966; }
    int64_t state_index(const State&) const { return -1; }
    int64_t n_states() const { return 0; }
    void gen(const State& s, int ai, Stream& rng, State& sp, double& r) const {
        const double th = s.x, w = s.v;
        const double u = p.force[ai] + (rng.uniform01() - 0.5) * p.noise;
        const double sn = det_sin(th), cs = o_cos(th), s2 = det_sin(2.0 * th);
        const double aml = (p.alpha * p.m) * p.l;
        const double num = (p.g * sn - ((aml * (w * w)) * s2) * 0.5) - (p.alpha * cs) * u;
        const double den = (4.0 / 3.0) * p.l - (p.alpha * p.l) * (cs * cs);
        const double acc = num / den;
        sp = MCState{th + w * p.dt, w + acc * p.dt};
        r = isterminal(sp) ? p.cost : 0.0;
    }
};

// mcts.jl_b200/plugins/continuous_pendulum.cuh: the pendulum above with a continuous torque -- a run-time plugin with a continuous
// action space (actions(mdp) = Box([-u_max], [u_max]); next_action = RandomActionGenerator's rand(rng, actions(mdp, s))).
struct ContinuousPendulumParams {
    double discount, g, m, l, M, alpha, dt, cost;
    double u_max, noise, effort, gain;
};
struct Vec1Action {
    double v[1];
};
constexpr int kOracleMdpContinuousPendulum = 1002;
struct ContinuousPendulum {
    static constexpr int kMdpId = kOracleMdpContinuousPendulum;
    static constexpr bool kLiteralTransitions = true;
    using State = MCState;  // (theta, omega)
    typedef Vec1Action Action;
    static constexpr int kActions = 1;  // (no finite action set)
    static constexpr int kActionDoubles = 1;
    ContinuousPendulumParams p;
    explicit ContinuousPendulum(const ContinuousPendulumParams& q) : p(q) {}
    double discount() const { return p.discount; }
    int n_actions(const State&) const { return 0; }
    int max_actions() const { return 1; }
    bool isterminal(const State& s) const { return std::fabs(s.x) > 1.5707963267948966; }
    int64_t state_index(const State&) const { return -1; }
    int64_t n_states() const { return 0; }
    static std::array<uint64_t, 1> action_key(const Action& a) { return {dm_bits(a.v[0])}; }
    Action next_action(const State&, Stream& rng) const {
        Action a;
        a.v[0] = (0.0 - p.u_max) + (2.0 * p.u_max) * rng.uniform01();
        return a;
    }
    void gen(const State& s, const Action& a, Stream& rng, State& sp, double& r) const {
        const double th = s.x, w = s.v;
        const double u = a.v[0] + (rng.uniform01() - 0.5) * p.noise;
        const double sn = det_sin(th), cs = o_cos(th), s2 = det_sin(2.0 * th);
        const double aml = (p.alpha * p.m) * p.l;
        const double num = (p.g * sn - ((aml * (w * w)) * s2) * 0.5) - (p.alpha * cs) * u;
        const double den = (4.0 / 3.0) * p.l - (p.alpha * p.l) * (cs * cs);
        const double acc = num / den;
        sp = MCState{th + w * p.dt, w + acc * p.dt};
        r = isterminal(sp) ? p.cost : 0.0 - p.effort * (a.v[0] * a.v[0]);
    }
    // rollout policy 100 of the plugin: proportional push against the fall; other ids: no torque
    Action plugin_policy(int policy, const State& s) const {
        Action a;
        double u = 0.0;
        if (policy == MCTSB200_POLICY_PLUGIN) {
            u = p.gain * (s.x + 0.3 * s.v);
            u = u > p.u_max ? p.u_max : (u < 0.0 - p.u_max ? 0.0 - p.u_max : u);
        }
        a.v[0] = u;
        return a;
    }
};

// The continuous-action model of include/mctsb200.h (mctsb200_gaussian_box_params): the MDP of test/discussion_514.jl
// (states / actions are Boxes, transition = MvNormal(s, Diagonal(var))) with an optional drift and quadratic cost. next_action is
// RandomActionGenerator on a Box: rand(rng, actions(mdp, s)) = one uniform draw per coordinate (src/action_gen.jl:11-13).
struct BoxState {
    double v[MCTSB200_BOX_STATE_DOUBLES];
    bool operator==(const BoxState& o) const {
        for (int i = 0; i < MCTSB200_BOX_STATE_DOUBLES; ++i)
            if (dm_bits(v[i]) != dm_bits(o.v[i])) return false;
        return true;
    }
};
struct BoxStateHash {
    size_t operator()(const BoxState& s) const {
        uint64_t h = 0x7F4A7C159E3779B9ull;
        for (int i = 0; i < MCTSB200_BOX_STATE_DOUBLES; ++i) h = (h ^ dm_bits(s.v[i])) * 0x9E3779B97F4A7C15ull;
        return std::hash<uint64_t>()(h);
    }
};
struct BoxAction {
    double v[MCTSB200_BOX_ACTION_DOUBLES];
};
struct GaussianBox {
    static constexpr int kMdpId = MCTSB200_MDP_GAUSSIAN_BOX;
    static constexpr bool kLiteralTransitions = true;  // continuous noise: transitions[sanode] is kept as the push-ordered vector
    using State = BoxState;
    typedef BoxAction Action;
    static constexpr int kActions = 1;  // (no finite action set)
    static constexpr int kActionDoubles = MCTSB200_BOX_ACTION_DOUBLES;
    mctsb200_gaussian_box_params p;
    double sd[3];
    explicit GaussianBox(const mctsb200_gaussian_box_params& q) : p(q) {
        for (int i = 0; i < 3; ++i) sd[i] = std::sqrt(q.var[i]);
    }
    double discount() const { return p.discount; }
    int n_actions(const State&) const { return 0; }
    int max_actions() const { return 1; }
    bool isterminal(const State&) const { return false; }
    int64_t state_index(const State&) const { return -1; }
    int64_t n_states() const { return 0; }
    // isequal on the action vector == bitwise equality of its coordinates
    static std::array<uint64_t, MCTSB200_BOX_ACTION_DOUBLES> action_key(const Action& a) {
        std::array<uint64_t, MCTSB200_BOX_ACTION_DOUBLES> k;
        for (int i = 0; i < MCTSB200_BOX_ACTION_DOUBLES; ++i) k[(size_t)i] = dm_bits(a.v[i]);
        return k;
    }
    Action next_action(const State&, Stream& rng) const {
        Action a;
        for (int i = 0; i < MCTSB200_BOX_ACTION_DOUBLES; ++i) a.v[i] = p.a_lo[i] + (p.a_hi[i] - p.a_lo[i]) * rng.uniform01();
        return a;
    }
    // two Box-Muller pairs from four words: z0, z1 from (w0, w1), z2 from (w2, w3)
    void gen(const State& s, const Action& a, Stream& rng, State& sp, double& r) const {
        const double two_pi = 6.283185307179586;
        double z[3];
        {
            const double u1 = ((double)rng.next() + 1.0) * (1.0 / 4294967296.0), u2 = rng.uniform01();
            const double rad = std::sqrt(-2.0 * o_log(u1)), ang = two_pi * u2;
            z[0] = rad * o_cos(ang);
            z[1] = rad * det_sin(ang);
        }
        {
            const double u1 = ((double)rng.next() + 1.0) * (1.0 / 4294967296.0), u2 = rng.uniform01();
            z[2] = std::sqrt(-2.0 * o_log(u1)) * o_cos(two_pi * u2);
        }
        double ss = 0.0;
        for (int i = 0; i < 3; ++i) {
            double x = s.v[i];
            if (i < MCTSB200_BOX_ACTION_DOUBLES) x = x + p.drift * a.v[i];
            x = x + sd[i] * z[i];
            sp.v[i] = x;
            ss = ss + x * x;
        }
        sp.v[3] = s.v[3];
        r = 0.0 - p.cost * ss;
    }
};

// ---------------------------------------------------------------------------------------------
// Work counters (same definitions as mctsb200_plan_stats)
// ---------------------------------------------------------------------------------------------
struct Counters {
    int64_t n_simulations = 0, n_tree_levels = 0, n_dpw_children_seen = 0, n_state_inserts = 0, n_action_inserts = 0,
            n_rollouts = 0, n_rollout_steps = 0, n_transition_samples = 0;
    void add(const Counters& o) {
        n_simulations += o.n_simulations; n_tree_levels += o.n_tree_levels; n_dpw_children_seen += o.n_dpw_children_seen;
        n_state_inserts += o.n_state_inserts; n_action_inserts += o.n_action_inserts; n_rollouts += o.n_rollouts;
        n_rollout_steps += o.n_rollout_steps; n_transition_samples += o.n_transition_samples;
    }
};

// ---------------------------------------------------------------------------------------------
// Leaf evaluation and node initialisers
// ---------------------------------------------------------------------------------------------
template <class M>
struct Hooks {
    const M& mdp;
    const mctsb200_solver_cfg& cfg;
    uint64_t tree_index;
    uint32_t iteration = 0;
    Counters* cnt;

    // action(policy, s) for the rollout policy
    typename M::Action policy_action(const typename M::State& s, Stream& rng) const {
        if constexpr (M::kActionDoubles > 0) {
            if constexpr (M::kMdpId == kOracleMdpContinuousPendulum) {
                if (cfg.rollout_policy != MCTSB200_POLICY_RANDOM) return mdp.plugin_policy(cfg.rollout_policy, s);
            }
            return mdp.next_action(s, rng);  // RandomPolicy: rand(rng, actions(problem, s))
        } else {
            return discrete_policy_action(s, rng);
        }
    }
    int discrete_policy_action(const typename M::State& s, Stream& rng) const {
        switch (cfg.rollout_policy) {
            case MCTSB200_POLICY_RANDOM:  // rand(rng, actions(problem, s))
                return rng.uniform_int(mdp.n_actions(s));
            case MCTSB200_POLICY_CONSTANT:
                return cfg.rollout_policy_param[0];
            case MCTSB200_POLICY_GW_SEEK: {  // notebooks/Domain_Knowledge_Example.ipynb cell 13
                const int64_t tx = cfg.rollout_policy_param[0], ty = cfg.rollout_policy_param[1];
                const int64_t sx = state_x(s), sy = state_y(s);
                if (tx > sx) return 3;       // :right
                else if (tx < sx) return 2;  // :left
                else if (ty > sy) return 0;  // :up
                else return 1;               // :down
            }
            case MCTSB200_POLICY_TABLE:  // FunctionPolicy(f) for a state-only f, tabulated over the dense state space
                return cfg.policy_table[mdp.state_index(s)];
            case MCTSB200_POLICY_PLUGIN:  // policy 100 of the shipped plugins
                if constexpr (M::kMdpId == kOracleMdpInvertedPendulum) return (s.x + 0.3 * s.v) > 0.0 ? 2 : 0;  // push against the fall
                return state_v(s) < 0.0 ? 0 : 2;                                                                    // (MountainCar family: as MC_ENERGY)
            case MCTSB200_POLICY_MC_ENERGY:
                return state_v(s) < 0.0 ? 0 : 2;
        }
        return 0;
    }
    static int64_t state_x(const GWPos& s) { return s.x; }
    static int64_t state_y(const GWPos& s) { return s.y; }
    static int64_t state_x(const MCState&) { return 0; }
    static int64_t state_y(const MCState&) { return 0; }
    static double state_v(const GWPos&) { return 0.0; }
    static double state_v(const MCState& s) { return s.v; }
    static int64_t state_x(const BoxState&) { return 0; }
    static int64_t state_y(const BoxState&) { return 0; }
    static double state_v(const BoxState&) { return 0.0; }

    // POMDPTools simulate(::RolloutSimulator, mdp, policy, s0)
    double rollout(typename M::State s, int64_t max_steps, Stream& rng) const {
        const double eps = cfg.rollout_eps;
        double disc = 1.0;
        double r_total = 0.0;
        int64_t step = 1;
        while (disc > eps && !mdp.isterminal(s) && step <= max_steps) {
            const typename M::Action a = policy_action(s, rng);
            typename M::State sp;
            double r;
            mdp.gen(s, a, rng, sp, r);
            r_total += disc * r;
            s = sp;
            disc *= mdp.discount();
            step += 1;
            cnt->n_rollout_steps += 1;
        }
        return r_total;
    }

    // estimate_value(estimator, mdp, state, remaining_depth)
    double estimate_value(const typename M::State& s, int64_t remaining_depth) const {
        switch (cfg.estimate_kind) {
            case MCTSB200_EST_CONSTANT:
                return cfg.estimate_const;  // convert(Float64, estimator): also for terminal states
            case MCTSB200_EST_TABLE:
                return cfg.value_table[mdp.state_index(s)];
            default: {
                const int64_t max_steps = cfg.rollout_max_depth == -1 ? remaining_depth : cfg.rollout_max_depth;
                const int K = cfg.rollouts_per_leaf < 1 ? 1 : cfg.rollouts_per_leaf;
                double sum = 0.0;
                for (int k = 0; k < K; ++k) {
                    Stream rs(cfg.seed, tree_index, iteration, 1u + (uint32_t)k);
                    const double v = rollout(s, max_steps, rs);
                    cnt->n_rollouts += 1;
                    if (K == 1) return v;
                    sum += v;
                }
                return sum / (double)K;
            }
        }
    }
    double init_Q(const typename M::State& s, const typename M::Action& a) const {
        if constexpr (M::kActionDoubles == 0) {
            if (cfg.init_q_table) return cfg.init_q_table[mdp.state_index(s) * M::kActions + a];
        }
        return cfg.init_Q;
    }
    int64_t init_N(const typename M::State& s, const typename M::Action& a) const {
        if constexpr (M::kActionDoubles == 0) {
            if (cfg.init_n_table) return cfg.init_n_table[mdp.state_index(s) * M::kActions + a];
        }
        return cfg.init_N;
    }
};

// Export container shared by both planners (mirrors mctsb200_tree_buffers)
struct ExportedTree {
    int64_t state_bytes = 0;
    std::vector<int64_t> total_n;
    std::vector<uint8_t> s_labels;
    std::vector<int64_t> child_offsets, child_ids;
    std::vector<int64_t> n;
    std::vector<double> q;
    std::vector<int32_t> a_labels;
    std::vector<int64_t> n_a_children, trans_offsets, trans_spnode, trans_count;
    std::vector<double> trans_r;
    int64_t action_doubles = 0;
    std::vector<double> a_label_vectors;  // continuous action spaces
    std::vector<int64_t> vis_said, vis_spid, vis_count;  // _vis_stats of a vanilla tree built with enable_tree_vis
};

// ---------------------------------------------------------------------------------------------
// Vanilla UCT  (src/vanilla.jl)
// ---------------------------------------------------------------------------------------------
template <class M, class H>
struct VanillaPlanner {
    using S = typename M::State;
    const M& mdp;
    const mctsb200_solver_cfg& cfg;
    Hooks<M> hooks;
    Counters cnt;
    Stream* rng = nullptr;

    // MCTSTree (src/vanilla.jl:64-94); ids are 1-based, index i-1 in the vectors
    std::unordered_map<S, int64_t, H> state_map;
    std::vector<std::vector<int64_t>> child_ids;
    std::vector<int64_t> total_n;
    std::vector<S> s_labels;
    std::vector<int64_t> n;
    std::vector<double> q;
    std::vector<int32_t> a_labels;
    std::map<std::pair<int64_t, int64_t>, int64_t> vis_stats;  // _vis_stats: (said => spid) => visits, only with enable_tree_vis

    VanillaPlanner(const M& m, const mctsb200_solver_cfg& c, uint64_t tree_index)
        : mdp(m), cfg(c), hooks{m, c, tree_index, 0, &cnt} {}

    // planner.tree across calls (reuse_tree, src/vanilla.jl:213-217): the tree outlives the planner object of one call
    struct Saved {
        std::unordered_map<S, int64_t, H> state_map;
        std::vector<std::vector<int64_t>> child_ids;
        std::vector<int64_t> total_n;
        std::vector<S> s_labels;
        std::vector<int64_t> n;
        std::vector<double> q;
        std::vector<int32_t> a_labels;
        std::map<std::pair<int64_t, int64_t>, int64_t> vis_stats;
    };
    Saved save() { return Saved{std::move(state_map), std::move(child_ids), std::move(total_n), std::move(s_labels), std::move(n), std::move(q), std::move(a_labels), std::move(vis_stats)}; }
    void load(Saved&& t) {
        state_map = std::move(t.state_map); child_ids = std::move(t.child_ids); total_n = std::move(t.total_n); s_labels = std::move(t.s_labels);
        n = std::move(t.n); q = std::move(t.q); a_labels = std::move(t.a_labels); vis_stats = std::move(t.vis_stats);
    }

    int64_t insert_node(const S& s) {
        s_labels.push_back(s);
        state_map[s] = (int64_t)s_labels.size();
        child_ids.emplace_back();
        int64_t tot = 0;
        for (int a = 0; a < mdp.n_actions(s); ++a) {
            const int64_t n0 = hooks.init_N(s, a);
            tot += n0;
            n.push_back(n0);
            q.push_back(hooks.init_Q(s, a));
            a_labels.push_back(a);
            child_ids.back().push_back((int64_t)n.size());
            cnt.n_action_inserts += 1;
        }
        total_n.push_back(tot);
        cnt.n_state_inserts += 1;
        return (int64_t)total_n.size();
    }

    int64_t best_sanode_Q(int64_t snode) const {
        double best_Q = -std::numeric_limits<double>::infinity();
        int64_t best = child_ids[snode - 1].front();
        for (int64_t sa : child_ids[snode - 1]) {
            if (q[sa - 1] > best_Q) {
                best_Q = q[sa - 1];
                best = sa;
            }
        }
        return best;
    }

    int64_t best_sanode_UCB(int64_t snode, double c) const {
        const auto& ch = child_ids[snode - 1];
        if (c == 0) {  // argmax(q, children): first maximal element
            int64_t best = ch.front();
            for (int64_t sa : ch)
                if (q[sa - 1] > q[best - 1]) best = sa;
            return best;
        }
        double best_UCB = -std::numeric_limits<double>::infinity();
        int64_t best = ch.front();
        const int64_t sn = total_n[snode - 1];
        for (int64_t sa : ch) {
            double UCB;
            if (n[sa - 1] == 0) {
                return sa;
            } else {
                UCB = q[sa - 1] + c * std::sqrt(o_log((double)sn) / (double)n[sa - 1]);
            }
            if (UCB > best_UCB) {
                best_UCB = UCB;
                best = sa;
            }
        }
        return best;
    }

    double simulate(int64_t node, int64_t depth) {
        const S s = s_labels[node - 1];
        if (mdp.isterminal(s)) {
            return 0.0;
        } else if (depth == 0) {
            return hooks.estimate_value(s, depth);
        }
        cnt.n_tree_levels += 1;
        const int64_t said = best_sanode_UCB(node, cfg.exploration_constant);
        S sp;
        double r;
        mdp.gen(s, a_labels[said - 1], *rng, sp, r);
        double qv;
        auto it = state_map.find(sp);
        int64_t spid;
        if (it == state_map.end()) {
            spid = insert_node(sp);
            qv = r + mdp.discount() * hooks.estimate_value(sp, depth - 1);
        } else {
            spid = it->second;
            qv = r + mdp.discount() * simulate(spid, depth - 1);
        }
        if (cfg.enable_tree_vis) vis_stats[{said, spid}] += 1;  // record_visit! (src/vanilla.jl:268-270, 328-334)
        total_n[node - 1] += 1;
        n[said - 1] += 1;
        q[said - 1] += (qv - q[said - 1]) / (double)n[said - 1];
        return qv;
    }

    // returns iterations done
    int64_t build_tree(const S& s, int64_t& root_out) {
        auto it = state_map.find(s);
        int64_t root = it == state_map.end() ? insert_node(s) : it->second;
        root_out = root;
        const auto start = std::chrono::steady_clock::now();
        int64_t done = 0;
        for (int64_t i = 0; i < cfg.n_iterations; ++i) {
            Stream descent(cfg.seed, hooks.tree_index, (uint32_t)i, 0);
            rng = &descent;
            hooks.iteration = (uint32_t)i;
            simulate(root, cfg.depth);
            cnt.n_simulations += 1;
            done = i + 1;
            if (std::isfinite(cfg.max_time)) {
                const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
                if (el >= cfg.max_time) break;
            }
        }
        return done;
    }

    void export_tree(ExportedTree& t) const {
        t.state_bytes = sizeof(S);
        t.total_n = total_n;
        t.s_labels.resize(s_labels.size() * sizeof(S));
        if (!s_labels.empty()) std::memcpy(t.s_labels.data(), s_labels.data(), t.s_labels.size());
        t.child_offsets.assign(1, 0);
        for (const auto& ch : child_ids) {
            for (int64_t c : ch) t.child_ids.push_back(c);
            t.child_offsets.push_back((int64_t)t.child_ids.size());
        }
        t.n = n;
        t.q = q;
        t.a_labels = a_labels;
        for (const auto& kv : vis_stats) {
            t.vis_said.push_back(kv.first.first);
            t.vis_spid.push_back(kv.first.second);
            t.vis_count.push_back(kv.second);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// DPW (src/dpw.jl, src/dpw_types.jl)
// ---------------------------------------------------------------------------------------------
template <class M, class H>
struct DPWPlanner {
    using S = typename M::State;
    using Act = typename M::Action;  // an index into actions(mdp), or a vector for a continuous action space
    using ActKey = decltype(M::action_key(std::declval<Act>()));
    const M& mdp;
    const mctsb200_solver_cfg& cfg;
    const bool literal_transitions;
    Hooks<M> hooks;
    Counters cnt;
    Stream* rng = nullptr;

    // DPWTree (src/dpw_types.jl:123-159)
    std::vector<int64_t> total_n;
    std::vector<std::vector<int64_t>> children;
    std::vector<S> s_labels;
    std::unordered_map<S, int64_t, H> s_lookup;
    std::vector<int64_t> n;
    std::vector<double> q;
    std::vector<std::vector<std::pair<int64_t, double>>> transitions;
    std::vector<Act> a_labels;
    std::map<std::pair<int64_t, ActKey>, int64_t> a_lookup;
    std::vector<int64_t> n_a_children;
    std::set<std::pair<int64_t, int64_t>> unique_transitions;

    DPWPlanner(const M& m, const mctsb200_solver_cfg& c, uint64_t tree_index, bool literal)
        : mdp(m), cfg(c), literal_transitions(literal), hooks{m, c, tree_index, 0, &cnt} {}

    // p.tree across calls (keep_tree, src/dpw.jl:31-37)
    struct Saved {
        std::vector<int64_t> total_n;
        std::vector<std::vector<int64_t>> children;
        std::vector<S> s_labels;
        std::unordered_map<S, int64_t, H> s_lookup;
        std::vector<int64_t> n;
        std::vector<double> q;
        std::vector<std::vector<std::pair<int64_t, double>>> transitions;
        std::vector<Act> a_labels;
        std::map<std::pair<int64_t, ActKey>, int64_t> a_lookup;
        std::vector<int64_t> n_a_children;
        std::set<std::pair<int64_t, int64_t>> unique_transitions;
    };
    Saved save() {
        return Saved{std::move(total_n), std::move(children), std::move(s_labels), std::move(s_lookup), std::move(n), std::move(q),
                     std::move(transitions), std::move(a_labels), std::move(a_lookup), std::move(n_a_children), std::move(unique_transitions)};
    }
    void load(Saved&& t) {
        total_n = std::move(t.total_n); children = std::move(t.children); s_labels = std::move(t.s_labels); s_lookup = std::move(t.s_lookup);
        n = std::move(t.n); q = std::move(t.q); transitions = std::move(t.transitions); a_labels = std::move(t.a_labels);
        a_lookup = std::move(t.a_lookup); n_a_children = std::move(t.n_a_children); unique_transitions = std::move(t.unique_transitions);
    }

    int64_t insert_state_node(const S& s, bool maintain_s_lookup) {
        total_n.push_back(0);
        children.emplace_back();
        s_labels.push_back(s);
        const int64_t snode = (int64_t)total_n.size();
        if (maintain_s_lookup) s_lookup[s] = snode;
        cnt.n_state_inserts += 1;
        return snode;
    }

    int64_t insert_action_node(int64_t snode, const Act& a, int64_t n0, double q0, bool maintain_a_lookup) {
        n.push_back(n0);
        q.push_back(q0);
        a_labels.push_back(a);
        transitions.emplace_back();
        const int64_t sanode = (int64_t)n.size();
        children[snode - 1].push_back(sanode);
        n_a_children.push_back(0);
        if (maintain_a_lookup) a_lookup[{snode, M::action_key(a)}] = sanode;
        cnt.n_action_inserts += 1;
        return sanode;
    }

    int64_t best_sanode(int64_t snode) const {
        double best_Q = -std::numeric_limits<double>::infinity();
        int64_t sanode = 0;
        for (int64_t child : children[snode - 1]) {
            if (q[child - 1] > best_Q) {
                best_Q = q[child - 1];
                sanode = child;
            }
        }
        return sanode;
    }

    int64_t best_sanode_UCB(int64_t snode, double c) const {
        double best_UCB = -std::numeric_limits<double>::infinity();
        int64_t sanode = 0;
        const double ltn = o_log((double)total_n[snode - 1]);
        for (int64_t child : children[snode - 1]) {
            const int64_t nn = n[child - 1];
            const double qq = q[child - 1];
            double UCB;
            if ((ltn <= 0 && nn == 0) || c == 0.0) {
                UCB = qq;
            } else {
                UCB = qq + c * std::sqrt(ltn / (double)nn);
            }
            if (UCB > best_UCB) {
                best_UCB = UCB;
                sanode = child;
            }
        }
        return sanode;
    }

    // rand(rng, transitions[sanode])
    std::pair<int64_t, double> sample_transition(int64_t sanode) {
        const auto& tr = transitions[sanode - 1];
        const int j = rng->uniform_int((int)tr.size());
        cnt.n_transition_samples += 1;
        if (literal_transitions) return tr[j];
        // multiset form: distinct spnodes in first-occurrence order, selected by cumulative count
        std::vector<std::pair<int64_t, double>> uniq;
        std::vector<int64_t> count;
        for (const auto& e : tr) {
            size_t k = 0;
            for (; k < uniq.size(); ++k)
                if (uniq[k].first == e.first) break;
            if (k == uniq.size()) {
                uniq.push_back(e);
                count.push_back(1);
            } else {
                count[k] += 1;
            }
        }
        int64_t cum = 0;
        for (size_t k = 0; k < uniq.size(); ++k) {
            cum += count[k];
            if (j < cum) return uniq[k];
        }
        return uniq.back();
    }

    double simulate(int64_t snode, int64_t d) {
        const S s = s_labels[snode - 1];
        if (mdp.isterminal(s)) {
            return 0.0;
        } else if (d == 0) {
            return hooks.estimate_value(s, d);
        }

        // action progressive widening
        if (cfg.enable_action_pw) {
            if ((double)children[snode - 1].size() <= cfg.k_action * std::pow((double)total_n[snode - 1], cfg.alpha_action)) {
                Act a;
                bool tabulated = false;
                if constexpr (M::kActionDoubles == 0) {
                    // next_action(f::Function, mdp, s, snode) = f(mdp, s, snode) (src/action_gen.jl:14; test/options.jl:50) for a state-only f,
                    // tabulated over the dense state space: draws nothing
                    if (cfg.next_action_table && mdp.state_index(s) >= 0) {
                        a = cfg.next_action_table[mdp.state_index(s)];
                        tabulated = true;
                    }
                }
                if (!tabulated) a = mdp.next_action(s, *rng);  // next_action(::RandomActionGenerator) = rand(rng, actions(mdp, s))
                if (!cfg.check_repeat_action || !a_lookup.count({snode, M::action_key(a)})) {
                    const int64_t n0 = hooks.init_N(s, a);
                    insert_action_node(snode, a, n0, hooks.init_Q(s, a), cfg.check_repeat_action != 0);
                    total_n[snode - 1] += n0;
                }
            }
        } else if (children[snode - 1].empty()) {
            if constexpr (M::kActionDoubles == 0) {  // (for a in actions(mdp, s): only a finite action set can be enumerated)
                for (int32_t a = 0; a < mdp.n_actions(s); ++a) {
                    const int64_t n0 = hooks.init_N(s, a);
                    insert_action_node(snode, a, n0, hooks.init_Q(s, a), false);
                    total_n[snode - 1] += n0;
                }
            }
        }

        cnt.n_tree_levels += 1;
        cnt.n_dpw_children_seen += (int64_t)children[snode - 1].size();
        const int64_t sanode = best_sanode_UCB(snode, cfg.exploration_constant);
        const Act a = a_labels[sanode - 1];

        // state progressive widening
        bool new_node = false;
        int64_t spnode;
        double r;
        S sp{};
        if ((cfg.enable_state_pw &&
             (double)n_a_children[sanode - 1] <= cfg.k_state * std::pow((double)n[sanode - 1], cfg.alpha_state)) ||
            n_a_children[sanode - 1] == 0) {
            mdp.gen(s, a, *rng, sp, r);
            auto it = s_lookup.end();
            if (cfg.check_repeat_state && (it = s_lookup.find(sp)) != s_lookup.end()) {
                spnode = it->second;
            } else {
                spnode = insert_state_node(sp, cfg.keep_tree || cfg.check_repeat_state);
                new_node = true;
            }
            transitions[sanode - 1].push_back({spnode, r});
            if (!cfg.check_repeat_state) {
                n_a_children[sanode - 1] += 1;
            } else if (!unique_transitions.count({sanode, spnode})) {
                unique_transitions.insert({sanode, spnode});
                n_a_children[sanode - 1] += 1;
            }
        } else {
            auto pr = sample_transition(sanode);
            spnode = pr.first;
            r = pr.second;
        }

        double qv;
        if (new_node) {
            qv = r + mdp.discount() * hooks.estimate_value(sp, d - 1);
        } else {
            qv = r + mdp.discount() * simulate(spnode, d - 1);
        }

        n[sanode - 1] += 1;
        total_n[snode - 1] += 1;
        q[sanode - 1] += (qv - q[sanode - 1]) / (double)n[sanode - 1];
        return qv;
    }

    // the search block of action_info; returns status
    int32_t search(const S& s, int64_t& root_out, int64_t& done_out) {
        if (mdp.isterminal(s)) return MCTSB200_ERR_TERMINAL_ROOT;
        int64_t snode;
        auto it = s_lookup.find(s);
        if (cfg.keep_tree && !total_n.empty()) {
            snode = it != s_lookup.end() ? it->second : insert_state_node(s, true);
        } else {
            snode = insert_state_node(s, cfg.check_repeat_state != 0);
        }
        root_out = snode;
        const auto start = std::chrono::steady_clock::now();
        done_out = 0;
        for (int64_t i = 0; i < cfg.n_iterations; ++i) {
            Stream descent(cfg.seed, hooks.tree_index, (uint32_t)i, 0);
            rng = &descent;
            hooks.iteration = (uint32_t)i;
            simulate(snode, cfg.depth);
            cnt.n_simulations += 1;
            done_out = i + 1;
            if (std::isfinite(cfg.max_time)) {
                const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
                if (el >= cfg.max_time) break;
            }
        }
        return MCTSB200_OK;
    }

    void export_tree(ExportedTree& t) const {
        t.state_bytes = sizeof(S);
        t.total_n = total_n;
        t.s_labels.resize(s_labels.size() * sizeof(S));
        if (!s_labels.empty()) std::memcpy(t.s_labels.data(), s_labels.data(), t.s_labels.size());
        t.child_offsets.assign(1, 0);
        for (const auto& ch : children) {
            for (int64_t c : ch) t.child_ids.push_back(c);
            t.child_offsets.push_back((int64_t)t.child_ids.size());
        }
        t.n = n;
        t.q = q;
        if constexpr (M::kActionDoubles == 0) {
            t.a_labels = a_labels;
        } else {
            t.a_labels.assign(a_labels.size(), -1);
            t.action_doubles = M::kActionDoubles;
            for (const Act& a : a_labels)
                for (int i = 0; i < M::kActionDoubles; ++i) t.a_label_vectors.push_back(a.v[i]);
        }
        t.n_a_children = n_a_children;
        t.trans_offsets.assign(1, 0);
        for (const auto& tr : transitions) {  // distinct spnodes, first-occurrence order, with counts
            const size_t base = t.trans_spnode.size();
            for (const auto& e : tr) {
                size_t k = base;
                for (; k < t.trans_spnode.size(); ++k)
                    if (t.trans_spnode[k] == e.first) break;
                if (k == t.trans_spnode.size()) {
                    t.trans_spnode.push_back(e.first);
                    t.trans_r.push_back(e.second);
                    t.trans_count.push_back(1);
                } else {
                    t.trans_count[k] += 1;
                }
            }
            t.trans_offsets.push_back((int64_t)t.trans_spnode.size());
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Batch driver
// ---------------------------------------------------------------------------------------------
struct RootResult {
    int32_t action = 0;
    double best_q = 0.0;
    int32_t status = 0;
    int32_t n_children = 0;
    int64_t total_n = 0;
    std::vector<int32_t> child_action;  // action index; for a continuous action space the child's position
    std::vector<int64_t> child_n;
    std::vector<double> child_q;
    std::vector<double> action_vec;     // continuous action space: the chosen action
};

static std::vector<std::unique_ptr<ExportedTree>> g_trees;
static std::vector<RootResult> g_roots;
// trees kept for the next call (reuse_tree / keep_tree): one per root index, valid for one (solver kind, MDP) pair
static std::vector<std::any> g_saved;
static int g_saved_kind = -1, g_saved_mdp = -1;
static std::mutex g_mu;

template <class M, class H>
static void plan_one(const M& mdp, const mctsb200_solver_cfg& cfg, const typename M::State& root, uint64_t tree_index,
                     bool literal, bool keep, RootResult& res, Counters& cnt, std::unique_ptr<ExportedTree>& tree_out,
                     int64_t& iterations_done, std::any* saved) {
    if constexpr (M::kActionDoubles > 0) {
        if (cfg.kind == MCTSB200_KIND_UCT) {  // vanilla UCT enumerates actions(mdp, s) (src/vanilla.jl:297)
            res.status = MCTSB200_ERR_UNSUPPORTED;
            return;
        }
    }
    if (cfg.kind == MCTSB200_KIND_UCT) {
      if constexpr (M::kActionDoubles == 0) {
        VanillaPlanner<M, H> p(mdp, cfg, tree_index);
        using SavedT = typename VanillaPlanner<M, H>::Saved;
        if (cfg.keep_tree && saved && saved->has_value()) p.load(std::move(std::any_cast<SavedT&>(*saved)));
        int64_t rootid = 0;
        iterations_done = p.build_tree(root, rootid);
        const int64_t best = p.best_sanode_Q(rootid);
        res.action = p.a_labels[best - 1];
        res.best_q = p.q[best - 1];
        res.status = MCTSB200_OK;
        res.total_n = p.total_n[rootid - 1];
        for (int64_t sa : p.child_ids[rootid - 1]) {
            res.child_action.push_back(p.a_labels[sa - 1]);
            res.child_n.push_back(p.n[sa - 1]);
            res.child_q.push_back(p.q[sa - 1]);
        }
        res.n_children = (int32_t)res.child_action.size();
        cnt = p.cnt;
        if (keep) {
            tree_out.reset(new ExportedTree());
            p.export_tree(*tree_out);
        }
        if (cfg.keep_tree && saved) *saved = p.save();
      }
    } else {
        DPWPlanner<M, H> p(mdp, cfg, tree_index, literal || M::kLiteralTransitions);
        using SavedT = typename DPWPlanner<M, H>::Saved;
        if (cfg.keep_tree && saved && saved->has_value()) p.load(std::move(std::any_cast<SavedT&>(*saved)));
        int64_t rootid = 0;
        res.status = p.search(root, rootid, iterations_done);
        if (res.status == MCTSB200_OK) {
            const int64_t best = p.best_sanode(rootid);
            if constexpr (M::kActionDoubles == 0) {
                if (best > 0) {
                    res.action = p.a_labels[best - 1];
                    res.best_q = p.q[best - 1];
                }
            } else {
                res.action = -1;
                res.action_vec.assign((size_t)M::kActionDoubles, 0.0);
            }
            res.total_n = p.total_n[rootid - 1];
            int32_t pos = 0;
            for (int64_t sa : p.children[rootid - 1]) {
                if constexpr (M::kActionDoubles == 0) {
                    res.child_action.push_back(p.a_labels[sa - 1]);
                } else {
                    res.child_action.push_back(pos);
                    if (sa == best) {
                        res.action = pos;
                        res.best_q = p.q[sa - 1];
                        for (int i = 0; i < M::kActionDoubles; ++i) res.action_vec[(size_t)i] = p.a_labels[sa - 1].v[i];
                    }
                }
                ++pos;
                res.child_n.push_back(p.n[sa - 1]);
                res.child_q.push_back(p.q[sa - 1]);
            }
            res.n_children = (int32_t)res.child_action.size();
        }
        cnt = p.cnt;
        if (keep) {
            tree_out.reset(new ExportedTree());
            p.export_tree(*tree_out);
        }
        if (cfg.keep_tree && saved) *saved = p.save();
    }
}

static int64_t algorithmic_bytes(const mctsb200_solver_cfg& cfg, const Counters& c, int n_actions) {
    if (cfg.kind == MCTSB200_KIND_UCT) return (80 + 16 * (int64_t)n_actions) * c.n_tree_levels + (48 + 16 * (int64_t)n_actions) * c.n_state_inserts;
    return 144 * c.n_tree_levels + 24 * c.n_dpw_children_seen + 56 * c.n_state_inserts + 72 * c.n_action_inserts;
}

template <class M, class H>
static int32_t plan_batch(const M& mdp, const mctsb200_solver_cfg& cfg, const typename M::State* roots, int64_t n_roots,
                          int64_t base, int n_threads, bool keep, bool literal, bool same_root, int32_t* action_out,
                          double* bestq_out, int32_t* status_out, mctsb200_plan_stats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<RootResult> results((size_t)n_roots);
    std::vector<std::unique_ptr<ExportedTree>> trees((size_t)n_roots);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        const bool compatible = cfg.keep_tree && (int64_t)g_saved.size() == n_roots && g_saved_kind == cfg.kind && g_saved_mdp == M::kMdpId;
        if (!compatible) g_saved.assign(cfg.keep_tree ? (size_t)n_roots : 0, std::any());
        g_saved_kind = cfg.kind;
        g_saved_mdp = M::kMdpId;
    }
    std::vector<Counters> per_thread((size_t)std::max(1, n_threads));
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> min_done{std::numeric_limits<int64_t>::max()};
    auto worker = [&](int tid) {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_roots) break;
            Counters c;
            int64_t done = 0;
            plan_one<M, H>(mdp, cfg, roots[same_root ? 0 : i], (uint64_t)(base + i), literal, keep, results[(size_t)i], c,
                           trees[(size_t)i], done, cfg.keep_tree ? &g_saved[(size_t)i] : nullptr);
            per_thread[(size_t)tid].add(c);
            int64_t cur = min_done.load();
            while (done < cur && !min_done.compare_exchange_weak(cur, done)) {
            }
        }
    };
    if (n_threads <= 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    Counters tot;
    for (const auto& c : per_thread) tot.add(c);
    for (int64_t i = 0; i < n_roots; ++i) {
        if (action_out) action_out[i] = results[(size_t)i].action;
        if (bestq_out) bestq_out[i] = results[(size_t)i].best_q;
        if (status_out) status_out[i] = results[(size_t)i].status;
    }
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->n_trees = n_roots;
        stats->n_simulations = tot.n_simulations;
        stats->n_tree_levels = tot.n_tree_levels;
        stats->n_dpw_children_seen = tot.n_dpw_children_seen;
        stats->n_state_inserts = tot.n_state_inserts;
        stats->n_action_inserts = tot.n_action_inserts;
        stats->n_rollouts = tot.n_rollouts;
        stats->n_rollout_steps = tot.n_rollout_steps;
        stats->n_transition_samples = tot.n_transition_samples;
        stats->algorithmic_bytes = algorithmic_bytes(cfg, tot, mdp.max_actions());
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        stats->lanes_per_tree = 1;
        stats->iterations_done = n_roots ? (int32_t)std::min<int64_t>(min_done.load(), INT32_MAX) : 0;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_trees = std::move(trees);
    g_roots = std::move(results);
    return MCTSB200_OK;
}

}  // namespace oracle

// ---------------------------------------------------------------------------------------------
// C interface for ctypes (tests / bench cpu_baseline only)
// ---------------------------------------------------------------------------------------------
template <class T>
static void copy_out(T* dst, const std::vector<T>& src) {
    if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(T));
}

namespace oracle {
// Episode driver: simulate(RolloutSimulator(max_steps = max_steps), mdp, planner, s0) for n_episodes start states in
// lockstep -- the loop bench/non-terminal_gw.jl:18,31 and bench/non-terminal_gw_dpw.jl time. Per environment step t every
// episode that is still running plans from its current state (plan call t uses the Philox key seed + t; finished episodes
// stay in the batch so that tree i always belongs to episode i) and then the environment moves with the stream
// (env_seed, episode index, t, stream 0). Returns the discounted return and the number of steps of every episode.
template <class M, class H>
static int32_t run_episodes_impl(const M& mdp, const mctsb200_solver_cfg& cfg0, typename M::State* s, int64_t n, int32_t max_steps,
                                 uint64_t env_seed, int64_t base, int n_threads, double* ret_out, int32_t* steps_out,
                                 mctsb200_plan_stats* stats) {
    std::vector<double> disc((size_t)n, 1.0), ret((size_t)n, 0.0);
    std::vector<int32_t> steps((size_t)n, 0), act((size_t)n), status((size_t)n);
    std::vector<char> active((size_t)n, 1);
    mctsb200_plan_stats tot;
    std::memset(&tot, 0, sizeof tot);
    for (int64_t i = 0; i < n; ++i) active[(size_t)i] = !mdp.isterminal(s[i]);
    for (int32_t t = 0; t < max_steps; ++t) {
        bool any = false;
        for (int64_t i = 0; i < n; ++i) any |= active[(size_t)i] != 0;
        if (!any) break;
        mctsb200_solver_cfg cfg = cfg0;
        cfg.seed = cfg0.seed + (uint64_t)t;
        mctsb200_plan_stats st;
        const int32_t rc = plan_batch<M, H>(mdp, cfg, s, n, base, n_threads, false, false, false, act.data(), nullptr, status.data(), &st);
        if (rc != MCTSB200_OK) return rc;
        tot.n_simulations += st.n_simulations; tot.n_tree_levels += st.n_tree_levels; tot.n_dpw_children_seen += st.n_dpw_children_seen;
        tot.n_state_inserts += st.n_state_inserts; tot.n_action_inserts += st.n_action_inserts; tot.n_rollouts += st.n_rollouts;
        tot.n_rollout_steps += st.n_rollout_steps; tot.n_transition_samples += st.n_transition_samples;
        tot.algorithmic_bytes += st.algorithmic_bytes;
        for (int64_t i = 0; i < n; ++i) {
            if (!active[(size_t)i]) continue;
            Stream rng(env_seed, (uint64_t)(base + i), (uint32_t)t, 0);
            typename M::State sp;
            double r;
            if constexpr (M::kActionDoubles > 0) {  // the planner's decision is a vector: a_labels[best_sanode] (src/dpw.jl:65)
                typename M::Action av;
                {
                    std::lock_guard<std::mutex> lk(g_mu);
                    const RootResult& res = g_roots[(size_t)i];
                    for (int k = 0; k < M::kActionDoubles; ++k) av.v[k] = res.action_vec.empty() ? 0.0 : res.action_vec[(size_t)k];
                }
                mdp.gen(s[i], av, rng, sp, r);
            } else {
                mdp.gen(s[i], act[(size_t)i], rng, sp, r);
            }
            ret[(size_t)i] = ret[(size_t)i] + disc[(size_t)i] * r;
            disc[(size_t)i] = disc[(size_t)i] * mdp.discount();
            s[i] = sp;
            steps[(size_t)i] += 1;
            if (mdp.isterminal(sp)) active[(size_t)i] = 0;
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        if (ret_out) ret_out[i] = ret[(size_t)i];
        if (steps_out) steps_out[i] = steps[(size_t)i];
    }
    if (stats) {
        tot.n_trees = n;
        *stats = tot;
    }
    return MCTSB200_OK;
}

}  // namespace oracle

extern "C" {

// clear_tree!(planner) (src/vanilla.jl:148, src/dpw.jl:6-8): forget the trees kept for reuse
int32_t oracle_clear_trees(void) {
    std::lock_guard<std::mutex> lk(oracle::g_mu);
    oracle::g_saved.clear();
    oracle::g_saved_kind = oracle::g_saved_mdp = -1;
    return MCTSB200_OK;
}

int32_t oracle_plan(const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* params, const void* roots, int64_t n_roots,
                    int64_t root_index_base, int32_t n_threads, int32_t keep_trees, int32_t literal_transitions,
                    int32_t same_root, int32_t* action_out, double* bestq_out, int32_t* status_out,
                    mctsb200_plan_stats* stats) {
    using namespace oracle;
    if (!cfg || cfg->struct_size != (int32_t)sizeof(mctsb200_solver_cfg) || !params || (!roots && n_roots > 0)) return MCTSB200_ERR_INVALID_ARG;
    if (mdp_id == MCTSB200_MDP_GRIDWORLD) {
        GridWorld m(*(const mctsb200_gridworld_params*)params);
        return plan_batch<GridWorld, GWPosHash>(m, *cfg, (const GWPos*)roots, n_roots, root_index_base, n_threads, keep_trees != 0,
                                                literal_transitions != 0, same_root != 0, action_out, bestq_out, status_out, stats);
    } else if (mdp_id == MCTSB200_MDP_MOUNTAINCAR) {
        MountainCar m(*(const mctsb200_mountaincar_params*)params);
        return plan_batch<MountainCar, MCStateHash>(m, *cfg, (const MCState*)roots, n_roots, root_index_base, n_threads,
                                                    keep_trees != 0, literal_transitions != 0, same_root != 0, action_out, bestq_out,
                                                    status_out, stats);
    } else if (mdp_id == kOracleMdpNoisyMountainCar) {
        NoisyMountainCar m(*(const NoisyMountainCarParams*)params);
        return plan_batch<NoisyMountainCar, MCStateHash>(m, *cfg, (const MCState*)roots, n_roots, root_index_base, n_threads,
                                                         keep_trees != 0, literal_transitions != 0, same_root != 0, action_out, bestq_out,
                                                         status_out, stats);
    } else if (mdp_id == kOracleMdpInvertedPendulum) {
        InvertedPendulum m(*(const InvertedPendulumParams*)params);
        return plan_batch<InvertedPendulum, MCStateHash>(m, *cfg, (const MCState*)roots, n_roots, root_index_base, n_threads, keep_trees != 0,
                                                         literal_transitions != 0, same_root != 0, action_out, bestq_out, status_out, stats);
    } else if (mdp_id == kOracleMdpContinuousPendulum) {
        ContinuousPendulum m(*(const ContinuousPendulumParams*)params);
        return plan_batch<ContinuousPendulum, MCStateHash>(m, *cfg, (const MCState*)roots, n_roots, root_index_base, n_threads, keep_trees != 0,
                                                           literal_transitions != 0, same_root != 0, action_out, bestq_out, status_out, stats);
    } else if (mdp_id == MCTSB200_MDP_GAUSSIAN_BOX) {
        GaussianBox m(*(const mctsb200_gaussian_box_params*)params);
        return plan_batch<GaussianBox, BoxStateHash>(m, *cfg, (const BoxState*)roots, n_roots, root_index_base, n_threads, keep_trees != 0,
                                                     literal_transitions != 0, same_root != 0, action_out, bestq_out, status_out, stats);
    }
    return MCTSB200_ERR_INVALID_ARG;
}

int32_t oracle_tree_vis_stats(int64_t root_i, int64_t* said, int64_t* spid, int64_t* count, int64_t capacity, int64_t* n_out) {
    using namespace oracle;
    std::lock_guard<std::mutex> lk(g_mu);
    if (root_i < 0 || root_i >= (int64_t)g_trees.size() || !g_trees[(size_t)root_i] || !n_out) return MCTSB200_ERR_STATE;
    const ExportedTree& t = *g_trees[(size_t)root_i];
    *n_out = (int64_t)t.vis_said.size();
    if (said && spid && count) {
        if (capacity < *n_out) return MCTSB200_ERR_INVALID_ARG;
        copy_out(said, t.vis_said);
        copy_out(spid, t.vis_spid);
        copy_out(count, t.vis_count);
    }
    return MCTSB200_OK;
}

int32_t oracle_root_action_vectors(int64_t first_root, int64_t n, double* out) {
    using namespace oracle;
    std::lock_guard<std::mutex> lk(g_mu);
    if (first_root < 0 || n < 0 || first_root + n > (int64_t)g_roots.size() || !out) return MCTSB200_ERR_INVALID_ARG;
    for (int64_t i = 0; i < n; ++i) {
        const RootResult& r = g_roots[(size_t)(first_root + i)];
        if (r.action_vec.empty()) return MCTSB200_ERR_STATE;
        for (size_t k = 0; k < r.action_vec.size(); ++k) out[(size_t)i * r.action_vec.size() + k] = r.action_vec[k];
    }
    return MCTSB200_OK;
}

int32_t oracle_root_stats(int64_t root_i, int32_t* n_children, int32_t* action_idx, int64_t* n, double* q, int64_t* total_n) {
    using namespace oracle;
    std::lock_guard<std::mutex> lk(g_mu);
    if (root_i < 0 || root_i >= (int64_t)g_roots.size()) return MCTSB200_ERR_INVALID_ARG;
    const RootResult& r = g_roots[(size_t)root_i];
    if (n_children) *n_children = r.n_children;
    if (total_n) *total_n = r.total_n;
    for (int i = 0; i < r.n_children; ++i) {
        if (action_idx) action_idx[i] = r.child_action[(size_t)i];
        if (n) n[i] = r.child_n[(size_t)i];
        if (q) q[i] = r.child_q[(size_t)i];
    }
    return MCTSB200_OK;
}

int32_t oracle_tree_sizes_query(int64_t root_i, mctsb200_tree_sizes* sizes) {
    using namespace oracle;
    std::lock_guard<std::mutex> lk(g_mu);
    if (root_i < 0 || root_i >= (int64_t)g_trees.size() || !g_trees[(size_t)root_i]) return MCTSB200_ERR_STATE;
    const ExportedTree& t = *g_trees[(size_t)root_i];
    sizes->n_state_nodes = (int64_t)t.total_n.size();
    sizes->n_action_nodes = (int64_t)t.n.size();
    sizes->n_transitions = (int64_t)t.trans_spnode.size();
    sizes->state_bytes = t.state_bytes;
    sizes->action_doubles = t.action_doubles;
    return MCTSB200_OK;
}

int32_t oracle_tree_export(int64_t root_i, const mctsb200_tree_buffers* b) {
    using namespace oracle;
    std::lock_guard<std::mutex> lk(g_mu);
    if (root_i < 0 || root_i >= (int64_t)g_trees.size() || !g_trees[(size_t)root_i]) return MCTSB200_ERR_STATE;
    const ExportedTree& t = *g_trees[(size_t)root_i];
    copy_out(b->total_n, t.total_n);
    if (b->s_labels && !t.s_labels.empty()) std::memcpy(b->s_labels, t.s_labels.data(), t.s_labels.size());
    copy_out(b->child_offsets, t.child_offsets);
    copy_out(b->child_ids, t.child_ids);
    copy_out(b->n, t.n);
    copy_out(b->q, t.q);
    copy_out(b->a_labels, t.a_labels);
    copy_out(b->n_a_children, t.n_a_children);
    copy_out(b->trans_offsets, t.trans_offsets);
    copy_out(b->trans_spnode, t.trans_spnode);
    copy_out(b->trans_r, t.trans_r);
    copy_out(b->trans_count, t.trans_count);
    copy_out(b->a_label_vectors, t.a_label_vectors);
    return MCTSB200_OK;
}

// Root-parallel ensemble of one root (BASELINE config 5). The reference has no such mode; the contract restated
// here is include/mctsb200.h's: sum_n[a] = sum_t N_t(a); sum_nq[a] = the sum over trees of the once-rounded double products
// N_t(a)*Q_t(a), each taken as a multiple of 2^-64 rounded toward zero, added exactly (128-bit integers) and rounded to double
// once -- so the sums do not depend on the order of the trees or on how a multi-GPU run shards them. Written independently of
// the library's limb arithmetic: frexp + __int128.
static bool fix64_of(double p, __int128& out) {
    if (!std::isfinite(p) || std::fabs(p) >= 1099511627776.0) return false;  // |p| < 2^40
    int e;
    const double m = std::frexp(p, &e);                          // p = m * 2^e, 0.5 <= |m| < 1
    const long long mant = (long long)std::ldexp(m, 53);         // exact 53-bit integer (signed)
    const int shift = e - 53 + 64;                               // p * 2^64 = mant * 2^shift
    const unsigned __int128 mag = (unsigned __int128)(mant < 0 ? -mant : mant);
    const unsigned __int128 fixed = shift >= 0 ? mag << shift : (-shift >= 64 ? (unsigned __int128)0 : mag >> (-shift));
    out = mant < 0 ? -(__int128)fixed : (__int128)fixed;
    return true;
}
int32_t oracle_plan_root_parallel(const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* params, const void* root_state,
                                  int64_t n_trees, int64_t tree_index_base, int32_t n_threads, double* sum_n_out,
                                  double* sum_nq_out, mctsb200_plan_stats* stats) {
    using namespace oracle;
    int32_t rc = oracle_plan(cfg, mdp_id, params, root_state, n_trees, tree_index_base, n_threads, 0, 1, 1, nullptr, nullptr,
                             nullptr, stats);
    if (rc != MCTSB200_OK) return rc;
    const int nA = mdp_id == MCTSB200_MDP_GRIDWORLD ? 4 : 3;
    long long sn[4] = {0, 0, 0, 0};
    __int128 snq[4] = {0, 0, 0, 0};
    bool inexact = false;
    std::lock_guard<std::mutex> lk(g_mu);
    for (const RootResult& r : g_roots) {
        for (int i = 0; i < r.n_children; ++i) {
            const int a = r.child_action[(size_t)i];
            sn[a] += (long long)r.child_n[(size_t)i];
            __int128 f;
            if (fix64_of((double)r.child_n[(size_t)i] * r.child_q[(size_t)i], f)) snq[a] += f;
            else inexact = true;
        }
    }
    for (int a = 0; a < nA; ++a) {
        sum_n_out[a] = (double)sn[a];
        sum_nq_out[a] = inexact ? std::nan("") : std::ldexp((double)snq[a], -64);
    }
    return MCTSB200_OK;
}

int32_t oracle_run_episodes(const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* params, void* states_inout, int64_t n_episodes,
                            int32_t max_steps, uint64_t env_seed, int64_t episode_index_base, int32_t n_threads, double* return_out,
                            int32_t* steps_out, mctsb200_plan_stats* stats) {
    using namespace oracle;
    if (!cfg || cfg->struct_size != (int32_t)sizeof(mctsb200_solver_cfg) || !params || !states_inout || n_episodes <= 0) return MCTSB200_ERR_INVALID_ARG;
    if (mdp_id == MCTSB200_MDP_GRIDWORLD) {
        GridWorld m(*(const mctsb200_gridworld_params*)params);
        return run_episodes_impl<GridWorld, GWPosHash>(m, *cfg, (GWPos*)states_inout, n_episodes, max_steps, env_seed, episode_index_base, n_threads,
                                                       return_out, steps_out, stats);
    } else if (mdp_id == MCTSB200_MDP_MOUNTAINCAR) {
        MountainCar m(*(const mctsb200_mountaincar_params*)params);
        return run_episodes_impl<MountainCar, MCStateHash>(m, *cfg, (MCState*)states_inout, n_episodes, max_steps, env_seed, episode_index_base,
                                                           n_threads, return_out, steps_out, stats);
    } else if (mdp_id == kOracleMdpNoisyMountainCar) {
        NoisyMountainCar m(*(const NoisyMountainCarParams*)params);
        return run_episodes_impl<NoisyMountainCar, MCStateHash>(m, *cfg, (MCState*)states_inout, n_episodes, max_steps, env_seed,
                                                                episode_index_base, n_threads, return_out, steps_out, stats);
    } else if (mdp_id == kOracleMdpInvertedPendulum) {
        InvertedPendulum m(*(const InvertedPendulumParams*)params);
        return run_episodes_impl<InvertedPendulum, MCStateHash>(m, *cfg, (MCState*)states_inout, n_episodes, max_steps, env_seed,
                                                                episode_index_base, n_threads, return_out, steps_out, stats);
    } else if (mdp_id == kOracleMdpContinuousPendulum) {
        ContinuousPendulum m(*(const ContinuousPendulumParams*)params);
        return run_episodes_impl<ContinuousPendulum, MCStateHash>(m, *cfg, (MCState*)states_inout, n_episodes, max_steps, env_seed,
                                                                  episode_index_base, n_threads, return_out, steps_out, stats);
    } else if (mdp_id == MCTSB200_MDP_GAUSSIAN_BOX) {
        GaussianBox m(*(const mctsb200_gaussian_box_params*)params);
        return run_episodes_impl<GaussianBox, BoxStateHash>(m, *cfg, (BoxState*)states_inout, n_episodes, max_steps, env_seed,
                                                            episode_index_base, n_threads, return_out, steps_out, stats);
    }
    return MCTSB200_ERR_INVALID_ARG;
}

// ---- small probes used by the unit tests ----
int32_t oracle_philox(uint64_t seed, uint64_t tree, uint32_t iteration, uint32_t stream_id, int32_t n_blocks, uint32_t* out) {
    oracle::Stream s(seed, tree, iteration, stream_id);
    for (int i = 0; i < 4 * n_blocks; ++i) out[i] = s.next();
    return 0;
}
int32_t oracle_philox_raw(const uint32_t* key, const uint32_t* ctr, uint32_t* out) {
    oracle::philox4x32_10(key, ctr, out);
    return 0;
}
double oracle_log(double x) { return oracle::o_log(x); }
double oracle_cos(double x) { return oracle::o_cos(x); }
double oracle_libm_log(double x) { return std::log(x); }
double oracle_libm_cos(double x) { return std::cos(x); }

// one gen() call with a fresh stream (seed, tree, iteration, stream_id); returns words consumed
int32_t oracle_gen(int32_t mdp_id, const void* params, const void* s, int32_t a, uint64_t seed, uint64_t tree,
                   uint32_t iteration, uint32_t stream_id, void* sp_out, double* r_out) {
    using namespace oracle;
    Stream rng(seed, tree, iteration, stream_id);
    if (mdp_id == MCTSB200_MDP_GRIDWORLD) {
        GridWorld m(*(const mctsb200_gridworld_params*)params);
        GWPos sp;
        m.gen(*(const GWPos*)s, a, rng, sp, *r_out);
        std::memcpy(sp_out, &sp, sizeof sp);
    } else {
        MountainCar m(*(const mctsb200_mountaincar_params*)params);
        MCState sp;
        m.gen(*(const MCState*)s, a, rng, sp, *r_out);
        std::memcpy(sp_out, &sp, sizeof sp);
    }
    return (int32_t)rng.ctr[0] * 4 - rng.have;
}

// one rollout (estimate_value with the configured policy) from s; iteration selects the stream
double oracle_rollout(const mctsb200_solver_cfg* cfg, int32_t mdp_id, const void* params, const void* s, uint64_t tree,
                      uint32_t iteration, int64_t remaining_depth, int64_t* steps_out) {
    using namespace oracle;
    Counters cnt;
    double v;
    if (mdp_id == MCTSB200_MDP_GRIDWORLD) {
        GridWorld m(*(const mctsb200_gridworld_params*)params);
        Hooks<GridWorld> h{m, *cfg, tree, iteration, &cnt};
        v = h.estimate_value(*(const GWPos*)s, remaining_depth);
    } else {
        MountainCar m(*(const mctsb200_mountaincar_params*)params);
        Hooks<MountainCar> h{m, *cfg, tree, iteration, &cnt};
        v = h.estimate_value(*(const MCState*)s, remaining_depth);
    }
    if (steps_out) *steps_out = cnt.n_rollout_steps;
    return v;
}

int32_t oracle_hardware_threads(void) { return (int32_t)std::thread::hardware_concurrency(); }

}  // extern "C"
