// plugins/continuous_pendulum.cuh -- an inverted pendulum with a CONTINUOUS torque: a run-time MDP plugin with a continuous action
// space (interface: csrc/custom.cuh, registered with mctsb200_mdp_register_continuous, state (theta, omega), action (u) in
// Box([-u_max], [u_max])). The dynamics are those of plugins/inverted_pendulum.cuh (POMDPModels.jl's InvertedPendulum, the model of the
// reference's notebooks/Test_Visualization.ipynb cell 4) with the three-valued force replaced by the action itself:
//     u = a + (rand(rng) - 0.5) * noise
//     domega/dt = (g sin(theta) - alpha m l omega^2 sin(2 theta) / 2 - alpha cos(theta) u) / (4/3 l - alpha l cos(theta)^2), Euler step dt
//     isterminal = |theta| > pi/2;  reward = isterminal(sp) ? cost : -effort * a^2
// next_action is the reference's default for DPW's action widening (RandomActionGenerator, src/action_gen.jl:11-13:
// rand(rng, actions(mdp, s))): one uniform draw from the Box. This is the case of test/discussion_514.jl -- actions(m) = Box(...) -- for a
// model the caller brings. The CPU oracle restates it (oracle/mcts_oracle.cpp: ContinuousPendulum, test-only id 1002).
struct ContinuousPendulum {
    struct Params {
        double discount;
        double g, m, l, M, alpha, dt, cost;  // 9.81, 2, 8, 8, 1/(m+M), 0.1, -1
        double u_max;                        // 50.0: actions(mdp) = Box([-u_max], [u_max])
        double noise;                        // 10.0
        double effort;                       // 0.0: reward -effort * a^2 while upright
        double gain;                         // policy 100: u = clamp(gain * (theta + 0.3 omega), -u_max, u_max)
    };
    __device__ static bool is_terminal(const Params&, const double* s) { return fabs(s[0]) > 1.5707963267948966; }
    __device__ static void next_action(const Params& p, const double*, mctsb200::Stream& rng, double* a) {
        a[0] = DM_ADD(DM_SUB(0.0, p.u_max), DM_MUL(DM_MUL(2.0, p.u_max), rng.uniform01()));
    }
    __device__ static void gen(const Params& p, const double* s, const double* a, mctsb200::Stream& rng, double* sp, double& r) {
        const double th = s[0], w = s[1];
        const double u = DM_ADD(a[0], DM_MUL(DM_SUB(rng.uniform01(), 0.5), p.noise));
        double sn, cs;
        det_sincos(th, &sn, &cs);  // (= det_sin(th), det_cos(th): one reduction, no quadrant branch for the lanes of a warp to split over)
        const double s2 = det_sin(DM_MUL(2.0, th));
        const double aml = DM_MUL(DM_MUL(p.alpha, p.m), p.l);
        const double num = DM_SUB(DM_SUB(DM_MUL(p.g, sn), DM_MUL(DM_MUL(DM_MUL(aml, DM_MUL(w, w)), s2), 0.5)), DM_MUL(DM_MUL(p.alpha, cs), u));
        const double den = DM_SUB(DM_MUL(4.0 / 3.0, p.l), DM_MUL(DM_MUL(p.alpha, p.l), DM_MUL(cs, cs)));
        const double acc = DM_DIV(num, den);
        sp[0] = DM_ADD(th, DM_MUL(w, p.dt));
        sp[1] = DM_ADD(w, DM_MUL(acc, p.dt));
        r = fabs(sp[0]) > 1.5707963267948966 ? p.cost : DM_SUB(0.0, DM_MUL(p.effort, DM_MUL(a[0], a[0])));
    }
    // policy 100: a proportional push against the fall (rollouts that last); any other id: no torque
    __device__ static void policy_action(const Params& p, int policy, const int*, const double* s, mctsb200::Stream&, double* a) {
        double u = 0.0;
        if (policy == MCTSB200_POLICY_PLUGIN) {
            u = DM_MUL(p.gain, DM_ADD(s[0], DM_MUL(0.3, s[1])));
            u = u > p.u_max ? p.u_max : (u < DM_SUB(0.0, p.u_max) ? DM_SUB(0.0, p.u_max) : u);
        }
        a[0] = u;
    }
};
