// plugins/noisy_mountaincar.cuh -- MountainCar with uniform velocity noise, as a run-time MDP plugin (interface: csrc/custom.cuh;
// registered with max_branch = 0: every (s, a) has a continuum of successors; the action set -- 2 to 4 velocity increments -- is a
// parameter, so the same text serves every action count the library has host code for). One 32-bit word of the caller's Philox stream per
// gen call: v' = clamp(v + a*0.001 + cos(3x)*-0.0025 + (u - 0.5)*noise), u = word * 2^-32. The CPU oracle restates it
// (oracle/mcts_oracle.cpp: NoisyMountainCar) and the GPU parity tests compare searches over it bit-for-bit.
struct NoisyMountainCar {
    struct Params {
        double discount;
        double cost, jackpot, noise;
        double push[4];  // velocity increment of action i (MountainCar: a * 0.001 for a in (-1., 0., 1.))
        int n_actions;   // informational: the action count is the registration's n_actions (2..4)
        int pad;
    };
    __device__ static bool is_terminal(const Params&, const double* s) { return s[0] >= 0.5; }
    __device__ static void gen(const Params& p, const double* s, int ai, mctsb200::Stream& rng, double* sp, double& r) {
        const double u = rng.uniform01();
        double v_ = DM_ADD(DM_ADD(s[1], p.push[ai]), DM_MUL(det_cos(DM_MUL(3.0, s[0])), -0.0025));
        v_ = DM_ADD(v_, DM_MUL(DM_SUB(u, 0.5), p.noise));
        v_ = v_ > 0.07 ? 0.07 : (v_ < -0.07 ? -0.07 : v_);  // clamp (two compares: FP64 min/max are multi-instruction sequences)
        double x_ = DM_ADD(s[0], v_);
        if (x_ < -1.2) {
            x_ = -1.2;
            v_ = 0.0;
        }
        sp[0] = x_;
        sp[1] = v_;
        r = x_ >= 0.5 ? p.jackpot : p.cost;
    }
    __device__ static int policy_action(const Params&, int policy, const int*, const double* s, mctsb200::Stream&) {
        return policy == MCTSB200_POLICY_PLUGIN ? (s[1] < 0.0 ? 0 : 2) : 0;
    }
};
