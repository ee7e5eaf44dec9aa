// plugins/inverted_pendulum.cuh -- POMDPModels.jl's InvertedPendulum (the model of the reference's
// notebooks/Test_Visualization.ipynb cell 4: DPWSolver(n_iterations, depth=10, exploration_constant=10., alpha_state=1/8) on
// InvertedPendulum()) as a run-time MDP plugin (interface: csrc/custom.cuh; register with max_branch = 0: the force noise makes
// every (s, a) a continuum of successors). Restated from the published package -- the reference does not vendor it, so, like
// MountainCar, nothing in /root/reference pins it:
//     state (theta, omega); actions [-50., 0., 50.]; isterminal = |theta| > pi/2; reward = isterminal(sp) ? cost : 0; discount 0.99
//     u = a + (rand(rng) - 0.5) * 10
//     domega/dt = (g sin(theta) - alpha m l omega^2 sin(2 theta) / 2 - alpha cos(theta) u) / (4/3 l - alpha l cos(theta)^2)
//     one Euler step of dt = 0.1
// The sine is det_sin of include/mctsb200_detmath.h (the reproducible twin of det_cos), one Philox word per gen call.
// The CPU oracle restates it (oracle/mcts_oracle.cpp: InvertedPendulum) and the GPU parity tests compare searches bit for bit.
struct InvertedPendulum {
    struct Params {
        double discount;
        double g, m, l, M, alpha, dt, cost;  // 9.81, 2, 8, 8, 1/(m+M), 0.1, -1
        double force[3];                     // [-50., 0., 50.]
        double noise;                        // 10.0
    };
    __device__ static bool is_terminal(const Params&, const double* s) { return fabs(s[0]) > 1.5707963267948966; }
    __device__ static void gen(const Params& p, const double* s, int ai, mctsb200::Stream& rng, double* sp, double& r) {
        const double th = s[0], w = s[1];
        const double u = DM_ADD(p.force[ai], DM_MUL(DM_SUB(rng.uniform01(), 0.5), p.noise));
        double sn, cs;
        det_sincos(th, &sn, &cs);  // (= det_sin(th), det_cos(th): one reduction, no quadrant branch for the lanes of a warp to split over)
        const double s2 = det_sin(DM_MUL(2.0, th));
        const double aml = DM_MUL(DM_MUL(p.alpha, p.m), p.l);
        const double num = DM_SUB(DM_SUB(DM_MUL(p.g, sn), DM_MUL(DM_MUL(DM_MUL(aml, DM_MUL(w, w)), s2), 0.5)), DM_MUL(DM_MUL(p.alpha, cs), u));
        const double den = DM_SUB(DM_MUL(4.0 / 3.0, p.l), DM_MUL(DM_MUL(p.alpha, p.l), DM_MUL(cs, cs)));
        const double acc = DM_DIV(num, den);
        sp[0] = DM_ADD(th, DM_MUL(w, p.dt));
        sp[1] = DM_ADD(w, DM_MUL(acc, p.dt));
        r = fabs(sp[0]) > 1.5707963267948966 ? p.cost : 0.0;
    }
    // policy 100: push against the fall (a bang-bang stabiliser), for rollouts that last
    __device__ static int policy_action(const Params&, int policy, const int*, const double* s, mctsb200::Stream&) {
        return policy == MCTSB200_POLICY_PLUGIN ? (DM_ADD(s[0], DM_MUL(0.3, s[1])) > 0.0 ? 2 : 0) : 1;
    }
};
