# TrajOptB200.jl — the reference-side binding of libtrajopt_b200.so (include/trajopt_b200.h).
#
# This is the stub a TrajectoryOptimization.jl (v0.1.1) maintainer adds to route the iLQR / AL-iLQR /
# ALTRO hot path to the B200 engine.  It keeps the package's API:
#
#     solve!(prob, B200(ALTROSolverOptions{Float64}()))          # one Problem, in place, returns a solver-like object
#     solve!(probs::Vector{<:Problem}, B200(opts))               # NEW: batched solve over same-shape problems
#     TrajOptB200.enable!()                                      # optional: make solve!(prob, opts) itself use the GPU
#
# and replaces (file:line in the reference tree)
#     solve!(prob, opts::AbstractSolverOptions)                  src/solvers.jl:91-94
#     solve!(prob, opts::ALTROSolverOptions)                     src/solvers/altro/altro_methods.jl:2-53
#     solve!(prob, opts::AugmentedLagrangianSolverOptions)       src/solvers/augmented_lagrangian/augmented_lagrangian_methods.jl:33-36
#
# NOTE: there is no Julia toolchain in the build image, so this file is syntax-reviewed only; the
# same C ABI is exercised end to end from Python (ctypes) by tests/ and bench.py.
module TrajOptB200

using TrajectoryOptimization
using LinearAlgebra
const TO = TrajectoryOptimization

const LIB = get(ENV, "TRAJOPT_B200_LIB", "libtrajopt_b200")

# ---- POD mirrors of include/trajopt_b200.h -----------------------------------------------------
struct TOConstraintRow
    kind::Int32; equality::Int32; var::Int32; is_bound::Int32
    sign::Float64; a::Float64; b::Float64; c::Float64; r::Float64
end

struct TOProblemDesc
    model::Int32; integrator::Int32; n::Int32; m::Int32; N::Int32; reserved0::Int32
    dt::Float64; tf::Float64
    Q::Ptr{Float64}; R::Ptr{Float64}; H::Ptr{Float64}; q::Ptr{Float64}; r::Ptr{Float64}; c::Float64
    Qf::Ptr{Float64}; qf::Ptr{Float64}; cf::Float64
    n_classes::Int32; reserved1::Int32
    class_of_knot::Ptr{Int32}; class_row_start::Ptr{Int32}; rows::Ptr{TOConstraintRow}
end

struct TOiLQROptions
    cost_tolerance::Float64; gradient_norm_tolerance::Float64
    iterations::Int32; dJ_counter_limit::Int32; square_root::Int32; iterations_linesearch::Int32
    line_search_lower_bound::Float64; line_search_upper_bound::Float64
    bp_reg_increase_factor::Float64; bp_reg_max::Float64; bp_reg_min::Float64; bp_reg_fp::Float64
    max_cost_value::Float64; max_state_value::Float64; max_control_value::Float64
    bp_reg_type::Int32; gradient_type::Int32
end
struct TOALOptions
    opts_uncon::TOiLQROptions
    cost_tolerance::Float64; cost_tolerance_intermediate::Float64
    gradient_norm_tolerance::Float64; gradient_norm_tolerance_intermediate::Float64
    constraint_tolerance::Float64
    iterations::Int32; kickout_max_penalty::Int32
    dual_min::Float64; dual_max::Float64; penalty_max::Float64; penalty_initial::Float64; penalty_scaling::Float64
end
struct TOALTROOptions
    opts_al::TOALOptions
    R_inf::Float64
    dynamically_feasible_projection::Int32; resolve_feasible_problem::Int32
    R_minimum_time::Float64; dt_max::Float64; dt_min::Float64
    projected_newton::Int32; pn_n_steps::Int32
    projected_newton_tolerance::Float64; pn_feasibility_tolerance::Float64; pn_active_set_tolerance::Float64
end
struct TOResult
    J::Float64; c_max::Float64
    iterations_total::Int32; iterations_outer::Int32; status::Int32; steps::Int32
end
struct TOIterRecord
    cost::Float64; dJ::Float64; gradient::Float64; expected::Float64; z::Float64; alpha::Float64; rho::Float64
    outer::Int32; iter::Int32
end
struct TOOuterRecord
    cost::Float64; c_max::Float64; penalty_max::Float64
    iterations_inner::Int32; pad::Int32
end

# ---- model / integrator recognition --------------------------------------------------------------
# User closures cannot run on the device: only the package's own Dynamics models map to a device id.
const MODEL_IDS = IdDict{Any,Int32}()
function __init__()
    D = TO.Dynamics
    MODEL_IDS[D.doubleintegrator.f] = 0
    MODEL_IDS[D.pendulum.f] = 1
    MODEL_IDS[D.car.f] = 2
    MODEL_IDS[D.cartpole.f] = 3
    MODEL_IDS[D.quadrotor.f] = 4
    MODEL_IDS[D.acrobot_model.f] = 5
    MODEL_IDS[D.doublependulum.f] = 6
end
const INTEG_IDS = Dict(:rk3 => Int32(0), :rk4 => Int32(1), :midpoint => Int32(2))

function model_ids(model)
    haskey(model.info, :fc) || error("TrajOptB200: the Problem needs a model discretized with rk3/rk4/midpoint")
    fc = model.info[:fc]
    haskey(MODEL_IDS, fc) || error("TrajOptB200: only TrajectoryOptimization.Dynamics models run on the device")
    integ = get(INTEG_IDS, model.info[:integration], nothing)
    integ === nothing && error("TrajOptB200: integrator $(model.info[:integration]) is not built into the engine")
    MODEL_IDS[fc], integ
end

# ---- constraints -> rows (src/constraints.jl:155-237,299-304; src/utils.jl:140-156) ----------------
function rows_of(con::TO.BoundConstraint, n, m, terminal::Bool)
    rows = TOConstraintRow[]
    for i = 1:n; isfinite(con.x_max[i]) && push!(rows, TOConstraintRow(0, 0, i - 1, 1, 1.0, con.x_max[i], 0, 0, 0)); end
    if !terminal
        for i = 1:m; isfinite(con.u_max[i]) && push!(rows, TOConstraintRow(0, 0, n + i - 1, 1, 1.0, con.u_max[i], 0, 0, 0)); end
    end
    for i = 1:n; isfinite(con.x_min[i]) && push!(rows, TOConstraintRow(0, 0, i - 1, 1, -1.0, con.x_min[i], 0, 0, 0)); end
    if !terminal
        for i = 1:m; isfinite(con.u_min[i]) && push!(rows, TOConstraintRow(0, 0, n + i - 1, 1, -1.0, con.u_min[i], 0, 0, 0)); end
    end
    rows
end

"""
Obstacle lists are closures in the reference (`problems/car_escape.jl:36-41`), so the caller
registers their data: `register_circles!(con, [(x, y, r), ...])` / `register_spheres!`.
A goal constraint built by `goal_constraint(xf)` is recognised by its label and evaluated at 0.
"""
const CIRCLES = IdDict{Any,Vector{NTuple{3,Float64}}}()
const SPHERES = IdDict{Any,Vector{NTuple{4,Float64}}}()
register_circles!(con, list) = (CIRCLES[con] = [Float64.(c) for c in list]; con)
register_spheres!(con, list) = (SPHERES[con] = [Float64.(s) for s in list]; con)

function rows_of(con::TO.Constraint, n, m, terminal::Bool)
    if con.label == :goal
        terminal || return TOConstraintRow[]
        v = zeros(con.p); con.c(v, zeros(n))                  # c(x) = x - xf  =>  xf = -c(0)
        return [TOConstraintRow(0, 1, i - 1, 0, 1.0, -v[i], 0, 0, 0) for i = 1:con.p]
    elseif haskey(CIRCLES, con)
        terminal && return TOConstraintRow[]
        return [TOConstraintRow(1, 0, 0, 0, 1.0, c[1], c[2], 0, c[3]) for c in CIRCLES[con]]
    elseif haskey(SPHERES, con)
        terminal && return TOConstraintRow[]
        return [TOConstraintRow(2, 0, 0, 0, 1.0, s[1], s[2], s[3], s[4]) for s in SPHERES[con]]
    end
    error("TrajOptB200: constraint :$(con.label) is a user closure; register its data or solve on the CPU")
end

# ---- Problem -> descriptor ---------------------------------------------------------------------------
struct Marshalled
    desc::TOProblemDesc
    keep::Vector{Any}   # arrays the descriptor points into
end

function marshal(prob::TO.Problem)
    model_id, integ = model_ids(prob.model)
    n, m, N = prob.model.n, prob.model.m, prob.N
    st, tm = prob.obj[1], prob.obj[N]
    (st isa TO.QuadraticCost && tm isa TO.QuadraticCost) || error("TrajOptB200: only QuadraticCost / LQRObjective objectives")
    for k = 2:N-1
        prob.obj[k] === st || error("TrajOptB200: the stage cost must be the same at every knot")
    end
    Q, R, H = Matrix{Float64}(st.Q), Matrix{Float64}(st.R), Matrix{Float64}(st.H)
    q, r = Vector{Float64}(st.q), Vector{Float64}(st.r)
    Qf, qf = Matrix{Float64}(tm.Q), Vector{Float64}(tm.q)
    classes = Vector{Vector{TOConstraintRow}}()
    class_of_knot = fill(Int32(-1), N)
    for k = 1:N
        rows = TOConstraintRow[]
        for con in prob.constraints[k]
            append!(rows, rows_of(con, n, m, k == N))
        end
        isempty(rows) && continue
        idx = findfirst(c -> c == rows, classes)
        if idx === nothing
            push!(classes, rows); idx = length(classes)
        end
        class_of_knot[k] = idx - 1
    end
    starts = Int32[0]
    flat = TOConstraintRow[]
    for rows in classes
        append!(flat, rows); push!(starts, length(flat))
    end
    isempty(flat) && push!(flat, TOConstraintRow(0, 0, 0, 0, 0, 0, 0, 0, 0))
    desc = TOProblemDesc(model_id, integ, n, m, N, 0, prob.dt, prob.tf,
                         pointer(Q), pointer(R), pointer(H), pointer(q), pointer(r), st.c,
                         pointer(Qf), pointer(qf), tm.c,
                         length(classes), 0, pointer(class_of_knot), pointer(starts), pointer(flat))
    Marshalled(desc, Any[Q, R, H, q, r, Qf, qf, class_of_knot, starts, flat])
end

# ---- options ------------------------------------------------------------------------------------------
# bp_reg_type / gradient_type (ilqr_solver.jl:47-52,76-80) as the ABI's enums
const REG_TYPES = Dict(:control => Int32(0), :state => Int32(1))
const GRAD_TYPES = Dict(:todorov => Int32(0), :feedforward => Int32(1), :ℓ2 => Int32(2), :ℓinf => Int32(3))
function c_opts(o::TO.iLQRSolverOptions)
    haskey(REG_TYPES, o.bp_reg_type) || error("TrajOptB200: unknown bp_reg_type $(o.bp_reg_type)")
    haskey(GRAD_TYPES, o.gradient_type) || error("TrajOptB200: unknown gradient_type $(o.gradient_type)")
    (o.bp_reg_type == :state && o.square_root) &&
        error("TrajOptB200: bp_reg_type=:state with square_root=true is not on the device path")
    TOiLQROptions(o.cost_tolerance, o.gradient_norm_tolerance, o.iterations, o.dJ_counter_limit,
        o.square_root, o.iterations_linesearch, o.line_search_lower_bound, o.line_search_upper_bound, o.bp_reg_increase_factor,
        o.bp_reg_max, o.bp_reg_min, o.bp_reg_fp, o.max_cost_value, o.max_state_value, o.max_control_value,
        REG_TYPES[o.bp_reg_type], GRAD_TYPES[o.gradient_type])
end
c_opts(o::TO.AugmentedLagrangianSolverOptions) = TOALOptions(c_opts(o.opts_uncon), o.cost_tolerance, o.cost_tolerance_intermediate,
    o.gradient_norm_tolerance, o.gradient_norm_tolerance_intermediate, o.constraint_tolerance, o.iterations,
    o.kickout_max_penalty, o.dual_min, o.dual_max, o.penalty_max, o.penalty_initial, o.penalty_scaling)
function c_opts(o::TO.ALTROSolverOptions)
    pn = o.opts_pn
    (!o.projected_newton || pn.solve_type == :feasible) ||
        error("TrajOptB200: ProjectedNewtonSolverOptions.solve_type=$(pn.solve_type) is not on the device path (only :feasible)")
    TOALTROOptions(c_opts(o.opts_al), o.R_inf, o.dynamically_feasible_projection, o.resolve_feasible_problem,
                   o.R_minimum_time, o.dt_max, o.dt_min, o.projected_newton, pn.n_steps,
                   o.projected_newton_tolerance, pn.feasibility_tolerance, pn.active_set_tolerance)
end

# ---- the solver-like object returned to the caller (docs/src/solvers.md:27-51) -----------------------
struct B200{O<:TO.AbstractSolverOptions}
    opts::O
    device::Int
end
B200(opts) = B200(opts, 0)

mutable struct B200Solver{T} <: TO.AbstractSolver{T}
    opts::Any
    stats::Dict{Symbol,Any}
end

check(rc, h) = rc == 0 || error("trajopt_b200: ", unsafe_string(ccall((:to_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

function solve_batch!(probs::Vector{<:TO.Problem}, opts, device::Int)
    p0 = probs[1]
    B, n, m, N = length(probs), p0.model.n, p0.model.m, p0.N
    M = marshal(p0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve M begin
        rc = ccall((:to_create, LIB), Cint, (Ref{TOProblemDesc}, Int32, Int32, Ref{Ptr{Cvoid}}), M.desc, B, device, h)
        rc == 0 || error("trajopt_b200: ", unsafe_string(ccall((:to_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    try
        x0 = hcat([p.x0 for p in probs]...)                               # n × B (column-major == [b][i])
        U0 = cat([hcat(p.U...) for p in probs]..., dims=3)                # m × (N-1) × B
        infeasible = !all(isnan, p0.X[1])                                 # altro_methods.jl:102
        X0 = infeasible ? cat([hcat(p.X...) for p in probs]..., dims=3) : nothing
        check(ccall((:to_set_batch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    h[], x0, U0, infeasible ? X0 : C_NULL), h[])
        outer_cap = opts isa TO.iLQRSolverOptions ? 0 : 2 * ((opts isa TO.ALTROSolverOptions ? opts.opts_al : opts).iterations + 2)
        check(ccall((:to_set_trace, LIB), Cint, (Ptr{Cvoid}, Int32, Int32), h[], 4096, outer_cap), h[])
        if opts isa TO.iLQRSolverOptions
            check(ccall((:to_solve_ilqr, LIB), Cint, (Ptr{Cvoid}, Ref{TOiLQROptions}), h[], c_opts(opts)), h[])
        elseif opts isa TO.AugmentedLagrangianSolverOptions
            check(ccall((:to_solve_al, LIB), Cint, (Ptr{Cvoid}, Ref{TOALOptions}), h[], c_opts(opts)), h[])
        else
            check(ccall((:to_solve_altro, LIB), Cint, (Ptr{Cvoid}, Ref{TOALTROOptions}), h[], c_opts(opts)), h[])
        end
        X = Array{Float64}(undef, n, N, B); U = Array{Float64}(undef, m, N - 1, B); dts = Array{Float64}(undef, N - 1, B)
        check(ccall((:to_get_solution, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h[], X, U, dts), h[])
        res = Vector{TOResult}(undef, B)
        check(ccall((:to_get_results, LIB), Cint, (Ptr{Cvoid}, Ptr{TOResult}), h[], res), h[])
        inner = Matrix{TOIterRecord}(undef, 4096, B); n_inner = Vector{Int32}(undef, B)
        outer = Matrix{TOOuterRecord}(undef, max(outer_cap, 1), B); n_outer = Vector{Int32}(undef, B)
        check(ccall((:to_get_trace, LIB), Cint, (Ptr{Cvoid}, Ptr{TOIterRecord}, Ptr{Int32}, Ptr{TOOuterRecord}, Ptr{Int32}),
                    h[], inner, n_inner, outer, n_outer), h[])
        solvers = B200Solver{Float64}[]
        for (b, p) in enumerate(probs)
            for k = 1:N; p.X[k] = X[:, k, b]; end
            for k = 1:N-1
                # minimum time: the reference leaves [u; ū] in prob.U (altro_methods.jl:81-86, SURVEY Q18);
                # the engine returns the first m controls and the time steps separately
                p.U[k] = U[:, k, b]
            end
            r = res[b]
            stats = Dict{Symbol,Any}(:status => r.status, :iterations_total => Int(r.iterations_total), :dt => dts[:, b])
            if r.iterations_outer > 0
                o = outer[1:n_outer[b], b]
                stats[:iterations] = Int(r.iterations_outer)
                stats[:cost] = [x.cost for x in o]; stats[:c_max] = [x.c_max for x in o]
                stats[:penalty_max] = [x.penalty_max for x in o]; stats[:iterations_inner] = [Int(x.iterations_inner) for x in o]
            else
                i = inner[1:n_inner[b], b]
                stats[:iterations] = Int(r.iterations_total)
                stats[:cost] = [x.cost for x in i]; stats[:dJ] = [x.dJ for x in i]; stats[:gradient] = [x.gradient for x in i]
            end
            push!(solvers, B200Solver{Float64}(opts, stats))
        end
        return solvers
    finally
        ccall((:to_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

TO.solve!(prob::TO.Problem, b::B200) = solve_batch!([prob], b.opts, b.device)[1]
TO.solve!(probs::Vector{<:TO.Problem}, b::B200) = solve_batch!(probs, b.opts, b.device)

# the copying variants (src/solvers.jl:104-108): the caller's problem stays untouched, `(prob_solved, solver)` comes back
function TO.solve(prob::TO.Problem, b::B200)
    p = copy(prob)
    solver = TO.solve!(p, b)
    return p, solver
end
function TO.solve(probs::Vector{<:TO.Problem}, b::B200)
    ps = [copy(p) for p in probs]
    solvers = TO.solve!(ps, b)
    return ps, solvers
end

"""
    print_log(solver; io=stderr)

The OuterLoop table of the reference's SolverLogger (src/logger.jl:10-38; columns iter / total / c_max as
augmented_lagrangian_methods.jl:93-96 logs them), rebuilt from the histories the engine returned.
"""
function print_log(solver::B200Solver; io::IO=stderr)
    st = solver.stats
    haskey(st, :c_max) || return
    logger = TO.SolverLogger(TO.OuterLoop, io=io)
    TO.add_level!(logger, TO.OuterLoop, [:iter, :total, :c_max, :info], [6, 7, 12, 50], print_color=:yellow, indent=0)
    total = 0
    for i = 1:length(st[:c_max])
        total += st[:iterations_inner][i]
        d = logger[TO.OuterLoop].data
        d[:iter] = i; d[:total] = total; d[:c_max] = st[:c_max][i]
        println(logger, TO.OuterLoop)
    end
end

"Route the package's own entry points to the GPU (drop-in): `solve!(prob, ALTROSolverOptions{Float64}())` etc."
function enable!(; device::Int=0)
    @eval TO.solve!(prob::TO.Problem, opts::TO.ALTROSolverOptions) = TrajOptB200.solve_batch!([prob], opts, $device)[1]
    @eval TO.solve!(prob::TO.Problem, opts::TO.AugmentedLagrangianSolverOptions) = TrajOptB200.solve_batch!([prob], opts, $device)[1]
    @eval TO.solve!(prob::TO.Problem, opts::TO.iLQRSolverOptions) = TrajOptB200.solve_batch!([prob], opts, $device)[1]
    nothing
end

end # module
