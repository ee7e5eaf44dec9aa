# MCTSB200.jl -- the glue a maintainer adds to MCTS.jl so that `solve(B200(solver), mdp)` returns a planner whose
# `action` / `action_info` / `value` / `ranked_actions` run the search loop on B200s through libmctsb200.so (include/mctsb200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia. The identical call sequence is exercised by the Python ctypes
# mirror (mcts.jl_b200/solvers.py, tests/test_api_mirror.py); this file is written against the same header and kept
# field-for-field in sync with mcts.jl_b200/_abi.py. Every function below names the Python function that runs the same calls.
#
# What it replaces (paths in JuliaPOMDP/MCTS.jl):
#   src/vanilla.jl:209-237  build_tree          -> mctsb200_plan (kind = UCT)
#   src/dpw.jl:29-67        the DPW search block -> mctsb200_plan (kind = DPW)
# What stays in Julia: the solver structs, `solve`, `action = first(action_info(...))`, the DPW try/catch -> default_action wrapper
# (src/dpw.jl:21,68-71), info assembly, `value`, `ranked_actions`, `clear_tree!`, visualisation -- all of them on REAL
# MCTSTree / DPWTree structs rebuilt from the device export, so `D3Tree(info[:tree])` (src/visualization.jl:52,180) works unchanged.
module MCTSB200

using POMDPs, POMDPTools, POMDPModels, MCTS, Random
import POMDPs: solve, action, value
import POMDPTools: action_info
import MCTS: ranked_actions, clear_tree!

const LIB = get(ENV, "MCTSB200_LIB", "libmctsb200")

# ---- status codes / enums (include/mctsb200.h) ----------------------------------------------------------
const OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_TERMINAL_ROOT, ERR_CAPACITY, ERR_CUDA, ERR_STATE = 0, -1, -2, -3, -4, -5, -6
const MDP_GRIDWORLD, MDP_MOUNTAINCAR, MDP_GAUSSIAN_BOX = Int32(0), Int32(1), Int32(2)
const KIND_UCT, KIND_DPW = Int32(0), Int32(1)
const EST_ROLLOUT, EST_CONSTANT, EST_TABLE = Int32(0), Int32(1), Int32(2)
const POLICY_RANDOM, POLICY_CONSTANT, POLICY_GW_SEEK, POLICY_MC_ENERGY, POLICY_TABLE = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
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
    next_action_table::Ptr{Int32}      # DPW's next_action as a state-only function, tabulated (NULL = RandomActionGenerator)
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
struct GaussianBoxParams               # mctsb200_gaussian_box_params (test/discussion_514.jl's model)
    discount::Float64; a_lo::NTuple{2,Float64}; a_hi::NTuple{2,Float64}; var::NTuple{3,Float64}; drift::Float64; cost::Float64
end

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
# dense state index of a state (rollout-policy / init tables): GridWorld's is the library's own, (x-1) + (y-1)*nx, terminal last
n_dense_states(m::SimpleGridWorld) = m.size[1] * m.size[2] + 1
dense_state(m::SimpleGridWorld, i) = i == m.size[1] * m.size[2] ? GWPos(-1, -1) : GWPos(i % m.size[1] + 1, i ÷ m.size[1] + 1)
n_dense_states(m) = 0

# A model whose actions(mdp) is a continuum (a Box: test/discussion_514.jl) says how many Float64 coordinates an action has; the library
# then returns action VECTORS (mctsb200_root_action_vectors, tree_buffers.a_label_vectors) and `action` converts them to actiontype(mdp).
action_doubles(m) = 0
action_from_vector(m, v::AbstractVector{Float64}) = convert(actiontype(m), v)   # e.g. SVector{2,Float64}(v), Vector{Float64}

# Run-time registration of a device plugin for a new MDP type (include/mctsb200.h: mctsb200_mdp_register). `source` is CUDA text
# defining `struct <plugin_struct>` (csrc/custom.cuh); afterwards add a `device_mdp(::YourMDP)` method returning
# (id, Ref(params_pod), sizeof(params_pod)) with a POD that mirrors the plugin's `Params` (`discount` first). max_branch = 0
# (unbounded successors per (s, a)) is always correct; 1 is an explicit promise of deterministic transitions.
function register_mdp(name::AbstractString, source::AbstractString, plugin_struct::AbstractString;
                      state_doubles::Integer, n_actions::Integer, max_branch::Integer = 0)
    id = Ref{Int32}(-1)
    check(ccall((:mctsb200_mdp_register, LIB), Int32, (Cstring, Cstring, Cstring, Int32, Int32, Int32, Ref{Int32}),
                name, source, plugin_struct, state_doubles, n_actions, max_branch, id))
    return id[]
end

# The same for a model whose actions(mdp) is a Box (test/discussion_514.jl): the struct defines next_action / gen over action vectors
# (csrc/custom.cuh), only DPWSolver plans for it and `action` returns the Vector{Float64} (see `action_vector` below).
function register_mdp_continuous(name::AbstractString, source::AbstractString, plugin_struct::AbstractString;
                                 state_doubles::Integer, action_doubles::Integer)
    id = Ref{Int32}(-1)
    check(ccall((:mctsb200_mdp_register_continuous, LIB), Int32, (Cstring, Cstring, Cstring, Int32, Int32, Ref{Int32}),
                name, source, plugin_struct, state_doubles, action_doubles, id))
    return id[]
end

# ---- context: one device, or several devices of this process (mctsb200_ctx_create_multi) ---------------------------------------
mutable struct Ctx
    h::Ptr{Cvoid}
    function Ctx(devices::AbstractVector{<:Integer})
        r = Ref{Ptr{Cvoid}}(C_NULL)
        if length(devices) == 1
            check(ccall((:mctsb200_ctx_create, LIB), Int32, (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), devices[1], C_NULL, r))
        else   # roots / episodes / ensemble trees are sharded inside the library; one ncclAllReduce per root-parallel decision
            ids = Int32.(devices)
            check(ccall((:mctsb200_ctx_create_multi, LIB), Int32, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), ids, length(ids), r))
        end
        c = new(r[])
        finalizer(c -> ccall((:mctsb200_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), c)
    end
end
Ctx(device::Integer = 0) = Ctx([device])
function configure!(c::Ctx, mdp)
    id, p, n = device_mdp(mdp)
    check(ccall((:mctsb200_mdp_configure, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t), c.h, id, p, n))
    return id
end

# ---- solver struct -> flat config; host closures have no device form and are refused ------------------------
unsupported(what) = error("libmctsb200 [$ERR_UNSUPPORTED]: $what cannot run on the device and there is no CPU fallback")
numeric(x::Number, _) = x
numeric(x, name) = unsupported("a Function/object-valued `$name`")
action_index(mdp, a) = Int32(findfirst(isequal(a), collect(actions(mdp))) - 1)

# -> (estimate_kind, estimate_const, rollout_policy, params, max_depth, eps, keep-alive table or nothing)
# (Python: solvers.build_cfg, the `ev` branch)
function rollout_fields(ev, mdp)
    zero4 = (Int32(0), Int32(0), Int32(0), Int32(0))
    ev isa Number && return (EST_CONSTANT, Float64(ev), POLICY_RANDOM, zero4, Int32(50), 0.0, nothing)
    ev isa MCTS.RolloutEstimator || unsupported("a Function/object-valued `estimate_value`")
    s = ev.solver
    if s isa RandomSolver || s isa RandomPolicy
        return (EST_ROLLOUT, 0.0, POLICY_RANDOM, zero4, Int32(ev.max_depth), Float64(ev.eps), nothing)
    elseif s isa FunctionPolicy || s isa Function
        # FunctionPolicy(f) / RolloutEstimator(x -> :up) (src/domain_knowledge.jl:57, test/options.jl:29): a state-only host function is
        # tabulated once over the dense state space (MCTSB200_POLICY_TABLE); a constant table is sent as POLICY_CONSTANT
        f = s isa FunctionPolicy ? s.f : s
        n = n_dense_states(mdp)
        n > 0 || unsupported("a Function-valued rollout policy on an MDP without a dense state index (register a device rollout-policy plugin)")
        tab = Int32[action_index(mdp, f(dense_state(mdp, i))) for i in 0:n-1]
        all(==(tab[1]), tab) && return (EST_ROLLOUT, 0.0, POLICY_CONSTANT, (tab[1], Int32(0), Int32(0), Int32(0)), Int32(ev.max_depth), Float64(ev.eps), nothing)
        return (EST_ROLLOUT, 0.0, POLICY_TABLE, zero4, Int32(ev.max_depth), Float64(ev.eps), tab)
    end
    unsupported("rollout policy $(typeof(s)) (register it as a device rollout-policy plugin)")
end

seed_of(rng::AbstractRNG) = rand(rng, UInt64)   # Philox key; the MersenneTwister stream itself is not reproduced

# -> (SolverCfg, objects the pointers inside it refer to: keep them alive across the ccall with GC.@preserve)
function cfg_from(s::MCTSSolver, mdp)
    ek, ec, pol, par, rmd, eps, tab = rollout_fields(s.estimate_value, mdp)
    cfg = SolverCfg(sizeof(SolverCfg), KIND_UCT, s.n_iterations, s.depth, 1, s.exploration_constant, s.max_time,
                    0.0, 0.0, 0.0, 0.0, 0, 0, 1, 1, s.reuse_tree, 0, numeric(s.init_Q, "init_Q"), numeric(s.init_N, "init_N"),
                    ek, pol, ec, par, rmd, 0, eps, seed_of(s.rng), C_NULL, C_NULL, C_NULL, tab === nothing ? C_NULL : pointer(tab),
                    s.enable_tree_vis, 0, C_NULL)   # enable_tree_vis: _vis_stats recorded on the device (src/vanilla.jl:268-270)
    return cfg, tab
end
function cfg_from(s::DPWSolver, mdp)
    # next_action(f::Function, mdp, s, snode) = f(mdp, s, snode) (src/action_gen.jl:14; test/options.jl:50): a function that looks at the
    # state only is a table over a dense state space (snode is passed as `nothing`: a generator that inspects the node is refused)
    natab = nothing
    if !(s.next_action isa MCTS.RandomActionGenerator)
        n = n_dense_states(mdp)
        (s.next_action isa Function && n > 0) || unsupported("a custom `next_action` object, or one on an MDP without a dense state index")
        natab = try
            Int32[action_index(mdp, s.next_action(mdp, dense_state(mdp, i), nothing)) for i in 0:n-1]
        catch
            unsupported("a `next_action` that inspects the tree node")
        end
    end
    ek, ec, pol, par, rmd, eps, tab = rollout_fields(s.estimate_value, mdp)
    cfg = SolverCfg(sizeof(SolverCfg), KIND_DPW, s.n_iterations, s.depth, 1, s.exploration_constant, s.max_time,
                    s.k_action, s.alpha_action, s.k_state, s.alpha_state, s.enable_action_pw, s.enable_state_pw,
                    s.check_repeat_state, s.check_repeat_action, s.keep_tree, 0, numeric(s.init_Q, "init_Q"), numeric(s.init_N, "init_N"),
                    ek, pol, ec, par, rmd, 0, eps, seed_of(s.rng), C_NULL, C_NULL, C_NULL, tab === nothing ? C_NULL : pointer(tab), 0, 0,
                    natab === nothing ? C_NULL : pointer(natab))
    return cfg, (tab, natab)
end

# ---- planners --------------------------------------------------------------------------------------------
"`B200(solver)` marks a solver for the device; `B200(solver, [0, 1, 2, 3])` for several devices of this process."
struct B200{S<:MCTS.AbstractMCTSSolver} <: Solver; solver::S; devices::Vector{Int}; end
B200(s) = B200(s, [0])
B200(s, device::Integer) = B200(s, [device])

# `tree` is what the reference's planners hold: a real MCTSTree / DPWTree (of the last single-state call), or nothing
mutable struct B200Planner{S,P<:Union{MDP,POMDP}} <: MCTS.AbstractMCTSPlanner{P}
    solver::S; mdp::P; ctx::Ctx; mdp_id::Int32; stats::PlanStats
    tree::Union{Nothing,MCTS.MCTSTree,MCTS.DPWTree}
end
function solve(b::B200, mdp::Union{MDP,POMDP})   # "no computation is done in solve" (src/vanilla.jl:151)
    cfg_from(b.solver, mdp)                       # refuse unsupported options up front
    c = Ctx(b.devices)
    B200Planner(b.solver, mdp, c, configure!(c, mdp), PlanStats(), nothing)
end

state_bytes(p::B200Planner) = sizeof(statetype(p.mdp))   # GWPos = SVector{2,Int} and MountainCar's SVector{2,Float64}: 16 B

"Batch extension: one tree per root state, all planned in one call. Returns (actions, best_Q). (Python: _Planner._plan)"
function action_info_batch(p::B200Planner, states::AbstractVector; root_index_base::Integer = 0)
    cfgv, keep = cfg_from(p.solver, p.mdp)
    cfg = Ref(cfgv)
    n = length(states)
    roots = collect(states)                       # Vector{statetype}: contiguous isbits values
    act = Vector{Int32}(undef, n); bq = Vector{Float64}(undef, n)
    GC.@preserve roots act bq keep begin
        check(ccall((:mctsb200_plan, LIB), Int32,
                    (Ptr{Cvoid}, Ref{SolverCfg}, Int32, Ptr{Cvoid}, Int64, Csize_t, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ref{PlanStats}),
                    p.ctx.h, cfg, p.mdp_id, pointer(roots), n, state_bytes(p), root_index_base, act, bq, C_NULL, p.stats))
    end
    ad = action_doubles(p.mdp)
    if ad > 0                                     # a continuous action space: a_labels[sanode] of src/dpw.jl:65 as vectors
        vec = Matrix{Float64}(undef, ad, n)
        check(ccall((:mctsb200_root_action_vectors, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), p.ctx.h, 0, n, vec))
        return [action_from_vector(p.mdp, vec[:, i]) for i in 1:n], bq
    end
    as = collect(actions(p.mdp))                  # device returns 0-based indices into actions(mdp) iteration order
    return [as[i + 1] for i in act], bq
end

# vanilla: `action_info(p, s)` returns `a, (tree=..., best_Q=...)` (src/vanilla.jl:158-162)
function action_info(p::B200Planner{<:MCTSSolver}, s)
    a, q = action_info_batch(p, [s])
    p.tree = export_tree(p, 0)
    return a[1], (tree = p.tree, best_Q = q[1])
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
        if p.solver.keep_tree || p.solver.tree_in_info || tree_in_info
            p.tree = export_tree(p, 0)           # src/dpw.jl:61-63
            (p.solver.tree_in_info || tree_in_info) && (info[:tree] = p.tree)
        end
    catch ex
        a = convert(actiontype(p.mdp), MCTS.default_action(p.solver.default_action, p.mdp, s, ex))
        info[:exception] = ex
    end
    return a, info
end
action(p::B200Planner, s) = first(action_info(p, s))
function clear_tree!(p::B200Planner)             # src/vanilla.jl:148, src/dpw.jl:6-8
    p.tree = nothing
    check(ccall((:mctsb200_clear_trees, LIB), Int32, (Ptr{Cvoid},), p.ctx.h))
end

# value(planner, s) / value(planner, s, a) (src/vanilla.jl:169-197): plans first if there is no tree, then reads the tree
# (Python: MCTSPlanner.value)
function value(p::B200Planner{<:MCTSSolver}, s)
    p.tree === nothing && action_info(p, s)
    return value(p.tree, s)                       # MCTS.jl's own method on the exported MCTSTree
end
function value(p::B200Planner{<:MCTSSolver}, s, a)
    p.tree === nothing && action_info(p, s)
    return value(p.tree, s, a)
end
# ranked_actions (src/util.jl:8-24) works through AbstractMCTSPlanner's method once `p.tree` is a real tree:
#   ranked_actions(p, s) == ranked_actions(p.tree, s)   (test/other.jl:15,29)

# ---- MCTSSolver.callback / solver.timer: host functions run between launches ------------------------------------
# callback(planner, i) after every iteration (src/vanilla.jl:231) = a progress hook with every = 1; a coarser `every`
# trades fidelity for fewer launches. The planner is passed through the user pointer.
function _progress_thunk(user::Ptr{Cvoid}, done::Int64, total::Int64)::Int32
    p = unsafe_pointer_to_objref(user)::B200Planner
    if p.solver isa MCTSSolver
        p.solver.callback(p, Int(done))
    else   # DPWSolver.show_progress (src/dpw.jl:45-53: a ProgressMeter advanced per iteration), as a coarse host hook
        print(stderr, "\rMCTS (DPW) ", done, "/", total, " iterations", done >= total ? "\n" : "")
    end
    return Int32(0)
end
_timer_thunk(user::Ptr{Cvoid})::Float64 = (unsafe_pointer_to_objref(user)::B200Planner).solver.timer()
function install_hooks!(p::B200Planner; every::Integer = 1)   # p must be a mutable struct kept alive (GC.@preserve) during plan
    if p.solver isa MCTSSolver || p.solver.show_progress
        step = p.solver isa MCTSSolver ? every : max(1, p.solver.n_iterations ÷ 20)
        check(ccall((:mctsb200_set_progress_callback, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), p.ctx.h,
                    @cfunction(_progress_thunk, Int32, (Ptr{Cvoid}, Int64, Int64)), pointer_from_objref(p), step))
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
    cfgv, keep = cfg_from(p.solver, p.mdp)
    cfg = Ref(cfgv)
    n = length(starts)
    s0 = collect(starts); sT = similar(s0)
    ret = Vector{Float64}(undef, n); steps = Vector{Int32}(undef, n)
    GC.@preserve s0 sT ret steps keep begin
        check(ccall((:mctsb200_run_episodes, LIB), Int32,
                    (Ptr{Cvoid}, Ref{SolverCfg}, Int32, Ptr{Cvoid}, Int64, Csize_t, Int32, UInt64, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Cvoid}, Ref{PlanStats}),
                    p.ctx.h, cfg, p.mdp_id, pointer(s0), n, state_bytes(p), sim.max_steps, env_seed, episode_index_base, ret, steps, pointer(sT), p.stats))
    end
    return ret, steps, sT
end
POMDPs.simulate(sim::RolloutSimulator, mdp::MDP, p::B200Planner, s0) = first(simulate_batch(sim, p, [s0]))[1]

# ---- root-parallel search of one root (BASELINE config 5): n_trees independent trees, merged on the device(s) -------------------
# With a multi-device Ctx the library splits the trees over its devices and merges their integer partial sums with ONE
# ncclAllReduce; the result does not depend on the number of devices. (Python: distributed.plan_root_parallel)
function action_root_parallel(p::B200Planner{<:MCTSSolver}, s, n_trees::Integer; tree_index_base::Integer = 0)
    cfgv, keep = cfg_from(p.solver, p.mdp)
    cfg = Ref(cfgv)
    na = length(actions(p.mdp))
    sn = zeros(Float64, na); snq = zeros(Float64, na); q = zeros(Float64, na)
    root = [s]
    GC.@preserve root keep begin
        check(ccall((:mctsb200_plan_root_parallel, LIB), Int32,
                    (Ptr{Cvoid}, Ref{SolverCfg}, Int32, Ptr{Cvoid}, Csize_t, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{PlanStats}),
                    p.ctx.h, cfg, p.mdp_id, pointer(root), state_bytes(p), n_trees, tree_index_base, sn, snq, C_NULL, p.stats))
    end
    a = Ref{Int32}(0); bq = Ref{Float64}(0.0)
    check(ccall((:mctsb200_root_parallel_decide, LIB), Int32, (Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ref{Int32}, Ref{Float64}, Ptr{Float64}),
                sn, snq, na, Float64(p.solver.init_Q), a, bq, q))
    return collect(actions(p.mdp))[a[] + 1], (best_Q = bq[], q = q, sum_n = sn)
end

# ---- info[:tree]: REAL MCTSTree / DPWTree structs rebuilt from the device export (two-call pattern) -----------------------------
struct TreeSizes; n_state_nodes::Int64; n_action_nodes::Int64; n_transitions::Int64; state_bytes::Int64; action_doubles::Int64; end
mutable struct TreeBuffers
    total_n::Ptr{Int64}; s_labels::Ptr{Cvoid}; child_offsets::Ptr{Int64}; child_ids::Ptr{Int64}
    n::Ptr{Int64}; q::Ptr{Float64}; a_labels::Ptr{Int32}
    n_a_children::Ptr{Int64}; trans_offsets::Ptr{Int64}; trans_spnode::Ptr{Int64}; trans_r::Ptr{Float64}; trans_count::Ptr{Int64}
    a_label_vectors::Ptr{Float64}
end
"The vectors of tree `root_i` of the last plan call; ids are 1-based in insertion order. (Python: Context.tree_export)"
function export_vectors(p::B200Planner, root_i::Integer)
    sz = Ref(TreeSizes(0, 0, 0, 0, 0))
    check(ccall((:mctsb200_tree_sizes_query, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{TreeSizes}), p.ctx.h, root_i, sz))
    ns, na, nt = sz[].n_state_nodes, sz[].n_action_nodes, sz[].n_transitions
    S = statetype(p.mdp)
    ad = Int(sz[].action_doubles)                 # > 0: the labels are vectors (a_labels then holds -1)
    out = (total_n = zeros(Int64, ns), s_labels = Vector{S}(undef, ns), child_offsets = zeros(Int64, ns + 1), child_ids = zeros(Int64, na),
           n = zeros(Int64, na), q = zeros(Float64, na), a_labels = zeros(Int32, na), n_a_children = zeros(Int64, na),
           trans_offsets = zeros(Int64, na + 1), trans_spnode = zeros(Int64, nt), trans_r = zeros(Float64, nt), trans_count = zeros(Int64, nt),
           a_label_vectors = zeros(Float64, max(ad, 1), na))
    GC.@preserve out begin
        b = TreeBuffers(pointer(out.total_n), pointer(out.s_labels), pointer(out.child_offsets), pointer(out.child_ids), pointer(out.n),
                        pointer(out.q), pointer(out.a_labels), pointer(out.n_a_children), pointer(out.trans_offsets), pointer(out.trans_spnode),
                        pointer(out.trans_r), pointer(out.trans_count), ad > 0 ? pointer(out.a_label_vectors) : Ptr{Float64}(C_NULL))
        check(ccall((:mctsb200_tree_export, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{TreeBuffers}), p.ctx.h, root_i, b))
    end
    return out
end
function vis_stats(p::B200Planner, root_i::Integer)   # tree._vis_stats (src/vanilla.jl:75, 328-334)
    n = Ref{Int64}(0)
    check(ccall((:mctsb200_tree_vis_stats, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64, Ref{Int64}),
                p.ctx.h, root_i, C_NULL, C_NULL, C_NULL, 0, n))
    said = zeros(Int64, n[]); spid = zeros(Int64, n[]); cnt = zeros(Int64, n[])
    check(ccall((:mctsb200_tree_vis_stats, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64, Ref{Int64}),
                p.ctx.h, root_i, said, spid, cnt, n[], n))
    return Dict{Pair{Int,Int},Int}((Int(said[i]) => Int(spid[i])) => Int(cnt[i]) for i in 1:n[])
end

function export_tree(p::B200Planner{<:MCTSSolver}, root_i::Integer)
    v = export_vectors(p, root_i)
    S, A = statetype(p.mdp), actiontype(p.mdp)
    as = collect(actions(p.mdp))
    t = MCTS.MCTSTree{S,A}(length(v.total_n))
    for (i, s) in enumerate(v.s_labels)            # src/vanilla.jl:64-94, field for field
        t.state_map[s] = i
        push!(t.s_labels, s); push!(t.total_n, v.total_n[i])
        push!(t.child_ids, Int.(v.child_ids[v.child_offsets[i]+1:v.child_offsets[i+1]]))
    end
    append!(t.n, v.n); append!(t.q, v.q); append!(t.a_labels, [as[k + 1] for k in v.a_labels])
    p.solver.enable_tree_vis && merge!(t._vis_stats, vis_stats(p, root_i))
    return t
end
function export_tree(p::B200Planner{<:DPWSolver}, root_i::Integer)
    v = export_vectors(p, root_i)
    S, A = statetype(p.mdp), actiontype(p.mdp)
    # the label of action node sa: an element of actions(mdp), or (continuous action spaces) the vector next_action drew
    label = if action_doubles(p.mdp) > 0
        sa -> action_from_vector(p.mdp, v.a_label_vectors[:, sa])
    else
        as = collect(actions(p.mdp))
        sa -> as[v.a_labels[sa] + 1]
    end
    t = MCTS.DPWTree{S,A}(length(v.total_n))
    for (i, s) in enumerate(v.s_labels)            # src/dpw_types.jl:123-159, field for field
        push!(t.total_n, v.total_n[i]); push!(t.s_labels, s)
        kids = Int.(v.child_ids[v.child_offsets[i]+1:v.child_offsets[i+1]])
        push!(t.children, kids)
        t.s_lookup[s] = i                          # (the newest node of a repeated state, as `s_lookup[s] = snode` leaves it)
        for sa in kids
            t.a_lookup[(i, label(sa))] = sa
        end
    end
    for sa in 1:length(v.n)
        push!(t.n, v.n[sa]); push!(t.q, v.q[sa]); push!(t.a_labels, label(sa)); push!(t.n_a_children, v.n_a_children[sa])
        tr = Tuple{Int,Float64}[]
        for k in v.trans_offsets[sa]+1:v.trans_offsets[sa+1], _ in 1:v.trans_count[k]   # transitions[sanode] as a multiset: one entry per push
            push!(tr, (Int(v.trans_spnode[k]), v.trans_r[k]))
            push!(t.unique_transitions, (sa, Int(v.trans_spnode[k])))
        end
        push!(t.transitions, tr)
    end
    return t
end

end # module
