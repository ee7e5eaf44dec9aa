// plugins/mountaincar.cuh -- POMDPModels.MountainCar written as a run-time MDP plugin (interface: csrc/custom.cuh; registered with
// mctsb200_mdp_register(name, <this text>, "MountainCarPlugin", /*state doubles*/ 2, /*actions*/ 3, /*max_branch*/ 1, &id)).
// Same arithmetic as the built-in MountainCarDev (csrc/models.cuh), so a search over this plugin is bit-identical to one over the
// built-in model: the parity anchor of the registration path (tests/test_gpu_parity.py::test_plugin_*), and the template to copy
// for a new model. State s = (x, v); actions [-1., 0., 1.] as indices 0..2.
struct MountainCarPlugin {
    struct Params {
        double discount;  // (first member: the library reads it)
        double cost, jackpot;
    };
    // isterminal(mdp, s)
    __device__ static bool is_terminal(const Params&, const double* s) { return s[0] >= 0.5; }
    // sp, r = @gen(:sp, :r)(mdp, s, a, rng):  v' = clamp(v + a*0.001 + cos(3x)*-0.0025, -0.07, 0.07); x' = x + v'; inelastic wall at -1.2.
    // DM_ADD / DM_MUL / det_cos: individually rounded operations, identical on every device and in the CPU oracle.
    __device__ static void gen(const Params& p, const double* s, int ai, mctsb200::Stream&, double* sp, double& r) {
        const double push = ai == 0 ? -0.001 : (ai == 1 ? 0.0 : 0.001);  // a * 0.001 for a in (-1., 0., 1.): exact
        double v_ = DM_ADD(DM_ADD(s[1], push), DM_MUL(det_cos(DM_MUL(3.0, s[0])), -0.0025));
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
    // rollout policies beyond RANDOM / CONSTANT: id 100 (MCTSB200_POLICY_PLUGIN) = push in the direction of motion
    __device__ static int policy_action(const Params&, int policy, const int*, const double* s, mctsb200::Stream&) {
        return policy == MCTSB200_POLICY_PLUGIN ? (s[1] < 0.0 ? 0 : 2) : 0;
    }
};
