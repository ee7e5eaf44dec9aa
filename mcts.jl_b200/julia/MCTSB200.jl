# MCTSB200.jl -- the glue a maintainer adds to MCTS.jl so that `solve(solver, mdp)` returns a planner whose
# `action` / `action_info` run the search loop on a B200 through libmctsb200.so (include/mctsb200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia. The identical call sequence is exercised by
# the Python ctypes mirror (mcts.jl_b200/solvers.py, tests/test_api_mirror.py); this file is written against the
# same header and kept field-for-field in sync with mcts.jl_b200/_abi.py.
#
# What it replaces (paths in JuliaPOMDP/MCTS.jl):
#   src/vanilla.jl:209-237  build_tree          -> mctsb200_plan (kind = UCT)
#   src/dpw.jl:29-67        the DPW search block -> mctsb200_plan (kind = DPW)
# What stays in Julia: the solver structs, `solve`, `action = first(action_info(...))`, the DPW try/catch ->
# default_action wrapper (src/dpw.jl:21,68-71), info assembly, `value`, `clear_tree!`, visualisation.
module MCTSB200

using POMDPs, POMDPTools, POMDPModels, MCTS, Random
import POMDPs: solve, action
import POMDPTools: action_info

const LIB = get(ENV, "MCTSB200_LIB", "libmctsb200")

# ---- status codes / enums (include/mctsb200.h) ----------------------------------------------------------
const OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_TERMINAL_ROOT, ERR_CAPACITY, ERR_CUDA, ERR_STATE = 0, -1, -2, -3, -4, -5, -6
const MDP_GRIDWORLD, MDP_MOUNTAINCAR = Int32(0), Int32(1)
const KIND_UCT, KIND_DPW = Int32(0), Int32(1)
const EST_ROLLOUT, EST_CONSTANT, EST_TABLE = Int32(0), Int32(1), Int32(2)
const POLICY_RANDOM, POLICY_CONSTANT, POLICY_GW_SEEK, POLICY_MC_ENERGY = Int32(0), Int32(1), Int32(2), Int32(3)
const GW_MAX_SPECIAL = 32

last_error() = unsafe_string(ccall((:mctsb200_last_error, LIB), Cstring, ()))
# every non-zero status becomes a Julia exception, so DPW's `catch ex -> default_action` keeps working
check(status) = status == OK ? nothing : error("libmctsb200 [$status]: $(last_error())")

# ---- flat structs, same layout as the C header ----------------------------------------------------------
struct SolverCfg                       # mctsb200_solver_cfg
    struct_size::Int32; kind::Int32
    n_iterations::Int64
    depth::Int32; rollouts_per_leaf::Int32
    exploration_constant::Float64; max_time::Float64
    k_action::Float64; alpha_action::Float64; k_state::Float64; alpha_state::Float64
    enable_action_pw::Int32; enable_state_pw::Int32; check_repeat_state::Int32; check_repeat_action::Int32
    keep_tree::Int32; lanes_per_tree::Int32
    init_Q::Float64; init_N::Int64
    estimate_kind::Int32; rollout_policy::Int32
    estimate_const::Float64
    rollout_policy_param::NTuple{4,Int32}
    rollout_max_depth::Int32; keep_tree_calls::Int32
    rollout_eps::Float64
    seed::UInt64
    init_q_table::Ptr{Float64}; init_n_table::Ptr{Int64}; value_table::Ptr{Float64}; policy_table::Ptr{Int32}
    enable_tree_vis::Int32; reserved1::Int32
end

mutable struct PlanStats               # mctsb200_plan_stats
    n_trees::Int64; n_simulations::Int64; n_tree_levels::Int64; n_dpw_children_seen::Int64
    n_state_inserts::Int64; n_action_inserts::Int64; n_rollouts::Int64; n_rollout_steps::Int64
    n_transition_samples::Int64; n_select_exact::Int64; algorithmic_bytes::Int64
    kernel_ms::Float64; total_ms::Float64; h2d_bytes::Int64; d2h_bytes::Int64
    lanes_per_tree::Int32; n_kernel_launches::Int32; iterations_done::Int32; kernel_variant::Int32
    PlanStats() = new(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.0, 0.0, 0, 0, 0, 0, 0, 0)
end

struct GridWorldParams                 # mctsb200_gridworld_params
    nx::Int32; ny::Int32
    tprob::Float64; discount::Float64
    n_rewards::Int32; n_terminate::Int32
    reward_x::NTuple{GW_MAX_SPECIAL,Int32}; reward_y::NTuple{GW_MAX_SPECIAL,Int32}
    reward_v::NTuple{GW_MAX_SPECIAL,Float64}
    term_x::NTuple{GW_MAX_SPECIAL,Int32}; term_y::NTuple{GW_MAX_SPECIAL,Int32}
end

struct MountainCarParams; discount::Float64; cost::Float64; jackpot::Float64; end

# ---- device MDP plugins: an MDP opts in by defining `device_mdp` ------------------------------------------
pad(v, T) = ntuple(i -> i <= length(v) ? T(v[i]) : zero(T), GW_MAX_SPECIAL)
function device_mdp(m::SimpleGridWorld)
    rk = collect(keys(m.rewards)); tk = collect(m.terminate_from)
    length(rk) <= GW_MAX_SPECIAL && length(tk) <= GW_MAX_SPECIAL || error("too many reward/terminate cells for the device plugin")
    p = GridWorldParams(m.size[1], m.size[2], m.tprob, m.discount, length(rk), length(tk),
                        pad(first.(rk), Int32), pad(last.(rk), Int32), pad([m.rewards[k] for k in rk], Float64),
                        pad(first.(tk), Int32), pad(last.(tk), Int32))
    return MDP_GRIDWORLD, Ref(p), sizeof(GridWorldParams)
end
device_mdp(m::MountainCar) = (MDP_MOUNTAINCAR, Ref(MountainCarParams(m.discount, m.cost, m.jackpot)), sizeof(MountainCarParams))
device_mdp(m) = error("no device plugin registered for $(typeof(m)); libmctsb200 has no CPU fallback")

# Run-time registration of a device plugin for a new MDP type (include/mctsb200.h: mctsb200_mdp_register). `source` is CUDA text
# defining `struct <plugin_struct>` (csrc/custom.cuh); afterwards add a `device_mdp(::YourMDP)` method returning
# (id, Ref(params_pod), sizeof(params_pod)) with a POD that mirrors the plugin's `Params` (`discount` first).
function register_mdp(name::AbstractString, source::AbstractString, plugin_struct::AbstractString;
                      state_doubles::Integer, n_actions::Integer, max_branch::Integer = 0)
    id = Ref{Int32}(-1)
    check(ccall((:mctsb200_mdp_register, LIB), Int32, (Cstring, Cstring, Cstring, Int32, Int32, Int32, Ref{Int32}),
                name, source, plugin_struct, state_doubles, n_actions, max_branch, id))
    return id[]
end

# ---- context --------------------------------------------------------------------------------------------
mutable struct Ctx
    h::Ptr{Cvoid}
    function Ctx(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mctsb200_ctx_create, LIB), Int32, (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, r))
        c = new(r[])
        finalizer(c -> ccall((:mctsb200_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), c)
    end
end
function configure!(c::Ctx, mdp)
    id, p, n = device_mdp(mdp)
    check(ccall((:mctsb200_mdp_configure, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t), c.h, id, p, n))
    return id
end

# ---- solver struct -> flat config; host closures have no device form and are refused ------------------------
unsupported(what) = error("libmctsb200 [$ERR_UNSUPPORTED]: $what cannot run on the device and there is no CPU fallback")
numeric(x::Number, _) = x
numeric(x, name) = unsupported("a Function/object-valued `$name`")

function rollout_fields(ev, mdp)
    ev isa Number && return (EST_CONSTANT, Float64(ev), POLICY_RANDOM, (Int32(0), Int32(0), Int32(0), Int32(0)), Int32(50), 0.0)
    ev isa MCTS.RolloutEstimator || unsupported("a Function/object-valued `estimate_value`")
    s = ev.solver
    pol, par = if s isa RandomSolver || s isa RandomPolicy
        POLICY_RANDOM, (Int32(0), Int32(0), Int32(0), Int32(0))
    else
        unsupported("rollout policy $(typeof(s)) (register it as a device rollout-policy plugin)")
    end
    return (EST_ROLLOUT, 0.0, pol, par, Int32(ev.max_depth), Float64(ev.eps))
end

seed_of(rng::AbstractRNG) = rand(rng, UInt64)   # Philox key; the MersenneTwister stream itself is not reproduced

function cfg_from(s::MCTSSolver, mdp)
    s.enable_tree_vis && unsupported("`enable_tree_vis`")
    ek, ec, pol, par, rmd, eps = rollout_fields(s.estimate_value, mdp)
    SolverCfg(sizeof(SolverCfg), KIND_UCT, s.n_iterations, s.depth, 1, s.exploration_constant, s.max_time,
              0.0, 0.0, 0.0, 0.0, 0, 0, 1, 1, s.reuse_tree, 0, numeric(s.init_Q, "init_Q"), numeric(s.init_N, "init_N"),
              ek, pol, ec, par, rmd, 0, eps, seed_of(s.rng), C_NULL, C_NULL, C_NULL)
end
function cfg_from(s::DPWSolver, mdp)
    s.show_progress && unsupported("`show_progress`")
    s.next_action isa MCTS.RandomActionGenerator || unsupported("a custom `next_action`")
    ek, ec, pol, par, rmd, eps = rollout_fields(s.estimate_value, mdp)
    SolverCfg(sizeof(SolverCfg), KIND_DPW, s.n_iterations, s.depth, 1, s.exploration_constant, s.max_time,
              s.k_action, s.alpha_action, s.k_state, s.alpha_state, s.enable_action_pw, s.enable_state_pw,
              s.check_repeat_state, s.check_repeat_action, s.keep_tree, 0, numeric(s.init_Q, "init_Q"), numeric(s.init_N, "init_N"),
              ek, pol, ec, par, rmd, 0, eps, seed_of(s.rng), C_NULL, C_NULL, C_NULL)
end

# ---- planners --------------------------------------------------------------------------------------------
"`B200(solver)` marks a solver for the device: `solve(B200(MCTSSolver(...)), mdp)`."
struct B200{S<:MCTS.AbstractMCTSSolver} <: Solver; solver::S; device::Int; end
B200(s) = B200(s, 0)

mutable struct B200Planner{S,P<:Union{MDP,POMDP}} <: MCTS.AbstractMCTSPlanner{P}
    solver::S; mdp::P; ctx::Ctx; mdp_id::Int32; stats::PlanStats
end
function solve(b::B200, mdp::Union{MDP,POMDP})   # "no computation is done in solve" (src/vanilla.jl:151)
    cfg_from(b.solver, mdp)                       # refuse unsupported options up front
    c = Ctx(b.device)
    B200Planner(b.solver, mdp, c, configure!(c, mdp), PlanStats())
end

state_bytes(p::B200Planner) = sizeof(statetype(p.mdp))   # GWPos = SVector{2,Int} and MountainCar's SVector{2,Float64}: 16 B

"Batch extension: one tree per root state, all planned in one call. Returns (actions, best_Q)."
function action_info_batch(p::B200Planner, states::AbstractVector; root_index_base::Integer = 0)
    cfg = Ref(cfg_from(p.solver, p.mdp))
    n = length(states)
    roots = collect(states)                       # Vector{statetype}: contiguous isbits values
    act = Vector{Int32}(undef, n); bq = Vector{Float64}(undef, n)
    GC.@preserve roots act bq begin
        check(ccall((:mctsb200_plan, LIB), Int32,
                    (Ptr{Cvoid}, Ref{SolverCfg}, Int32, Ptr{Cvoid}, Int64, Csize_t, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ref{PlanStats}),
                    p.ctx.h, cfg, p.mdp_id, pointer(roots), n, state_bytes(p), root_index_base, act, bq, C_NULL, p.stats))
    end
    as = collect(actions(p.mdp))                  # device returns 0-based indices into actions(mdp) iteration order
    return [as[i + 1] for i in act], bq
end

# vanilla: `action_info(p, s)` returns `a, (tree=..., best_Q=...)` (src/vanilla.jl:158-162)
function action_info(p::B200Planner{<:MCTSSolver}, s)
    a, q = action_info_batch(p, [s])
    return a[1], (tree = DeviceTree(p, 0), best_Q = q[1])
end
# DPW: same try/catch -> default_action contract as src/dpw.jl:18-74
function action_info(p::B200Planner{<:DPWSolver}, s; tree_in_info = false)
    local a
    info = Dict{Symbol,Any}()
    try
        isterminal(p.mdp, s) && error("MCTS cannot handle terminal states. action was called with\ns = $s")
        t0 = time()
        as, q = action_info_batch(p, [s])
        a = as[1]
        info[:search_time] = time() - t0
        info[:search_time_us] = 1e6 * info[:search_time]
        info[:tree_queries] = Int(p.stats.iterations_done)
        info[:best_Q] = q[1]
        (p.solver.tree_in_info || tree_in_info) && (info[:tree] = DeviceTree(p, 0))
    catch ex
        a = convert(actiontype(p.mdp), MCTS.default_action(p.solver.default_action, p.mdp, s, ex))
        info[:exception] = ex
    end
    return a, info
end
action(p::B200Planner, s) = first(action_info(p, s))
MCTS.clear_tree!(p::B200Planner) = check(ccall((:mctsb200_clear_trees, LIB), Int32, (Ptr{Cvoid},), p.ctx.h))

# ---- MCTSSolver.callback / solver.timer: host functions run between launches ------------------------------------
# callback(planner, i) after every iteration (src/vanilla.jl:231) = a progress hook with every = 1; a coarser `every`
# trades fidelity for fewer launches. The planner is passed through the user pointer.
function _progress_thunk(user::Ptr{Cvoid}, done::Int64, total::Int64)::Int32
    p = unsafe_pointer_to_objref(user)::B200Planner
    p.solver.callback(p, Int(done))
    return Int32(0)
end
_timer_thunk(user::Ptr{Cvoid})::Float64 = (unsafe_pointer_to_objref(user)::B200Planner).solver.timer()
function install_hooks!(p::B200Planner; every::Integer = 1)   # p must be a mutable struct kept alive (GC.@preserve) during plan
    if p.solver isa MCTSSolver
        check(ccall((:mctsb200_set_progress_callback, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), p.ctx.h,
                    @cfunction(_progress_thunk, Int32, (Ptr{Cvoid}, Int64, Int64)), pointer_from_objref(p), every))
    end
    check(ccall((:mctsb200_set_timer, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), p.ctx.h,
                @cfunction(_timer_thunk, Float64, (Ptr{Cvoid},)), pointer_from_objref(p)))
end
function clear_hooks!(p::B200Planner)
    check(ccall((:mctsb200_set_progress_callback, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), p.ctx.h, C_NULL, C_NULL, 0))
    check(ccall((:mctsb200_set_timer, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), p.ctx.h, C_NULL, C_NULL))
end

# ---- simulate(RolloutSimulator(max_steps = T), mdp, planner, s0) for a batch of start states, on the device ----------
"Closed-loop episodes (bench/non-terminal_gw.jl:18,31): returns (discounted returns, steps taken, final states)."
function simulate_batch(sim::RolloutSimulator, p::B200Planner, starts::AbstractVector; env_seed::Integer = 0, episode_index_base::Integer = 0)
    sim.max_steps === nothing && error("RolloutSimulator needs max_steps for the device episode loop")
    cfg = Ref(cfg_from(p.solver, p.mdp))
    n = length(starts)
    s0 = collect(starts); sT = similar(s0)
    ret = Vector{Float64}(undef, n); steps = Vector{Int32}(undef, n)
    GC.@preserve s0 sT ret steps begin
        check(ccall((:mctsb200_run_episodes, LIB), Int32,
                    (Ptr{Cvoid}, Ref{SolverCfg}, Int32, Ptr{Cvoid}, Int64, Csize_t, Int32, UInt64, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Cvoid}, Ref{PlanStats}),
                    p.ctx.h, cfg, p.mdp_id, pointer(s0), n, state_bytes(p), sim.max_steps, env_seed, episode_index_base, ret, steps, pointer(sT), p.stats))
    end
    return ret, steps, sT
end
POMDPs.simulate(sim::RolloutSimulator, mdp::MDP, p::B200Planner, s0) = first(simulate_batch(sim, p, [s0]))[1]

# ---- info[:tree]: the MCTSTree / DPWTree vectors, exported on demand (two-call pattern) -------------------
struct DeviceTree; planner::B200Planner; root_i::Int; end
struct TreeSizes; n_state_nodes::Int64; n_action_nodes::Int64; n_transitions::Int64; state_bytes::Int64; end
mutable struct TreeBuffers
    total_n::Ptr{Int64}; s_labels::Ptr{Cvoid}; child_offsets::Ptr{Int64}; child_ids::Ptr{Int64}
    n::Ptr{Int64}; q::Ptr{Float64}; a_labels::Ptr{Int32}
    n_a_children::Ptr{Int64}; trans_offsets::Ptr{Int64}; trans_spnode::Ptr{Int64}; trans_r::Ptr{Float64}; trans_count::Ptr{Int64}
end
function export_tree(t::DeviceTree)
    p = t.planner
    sz = Ref(TreeSizes(0, 0, 0, 0))
    check(ccall((:mctsb200_tree_sizes_query, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{TreeSizes}), p.ctx.h, t.root_i, sz))
    ns, na, nt = sz[].n_state_nodes, sz[].n_action_nodes, sz[].n_transitions
    S = statetype(p.mdp)
    out = (total_n = zeros(Int64, ns), s_labels = Vector{S}(undef, ns), child_offsets = zeros(Int64, ns + 1), child_ids = zeros(Int64, na),
           n = zeros(Int64, na), q = zeros(Float64, na), a_labels = zeros(Int32, na), n_a_children = zeros(Int64, na),
           trans_offsets = zeros(Int64, na + 1), trans_spnode = zeros(Int64, nt), trans_r = zeros(Float64, nt), trans_count = zeros(Int64, nt))
    GC.@preserve out begin
        b = TreeBuffers(pointer(out.total_n), pointer(out.s_labels), pointer(out.child_offsets), pointer(out.child_ids), pointer(out.n),
                        pointer(out.q), pointer(out.a_labels), pointer(out.n_a_children), pointer(out.trans_offsets), pointer(out.trans_spnode),
                        pointer(out.trans_r), pointer(out.trans_count))
        check(ccall((:mctsb200_tree_export, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{TreeBuffers}), p.ctx.h, t.root_i, b))
    end
    return out   # ids are 1-based in insertion order, exactly the MCTSTree / DPWTree vectors
end

end # module
