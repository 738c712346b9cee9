# baseline/run_reference.jl -- the TRUE reference arm: TrajectoryOptimization.jl v0.1.1 itself, all host threads.
#
# There is no Julia in the build image or on the GPU box, so this script cannot run there; `bench.py --impl reference` times the
# C++ port of the algorithm (oracle/) instead.  Anyone with Julia 1.1 and the package's Manifest can run this file to get
#   (1) the reference's own solves/s on the SAME synthetic batch bench.py uses (problem b seeded by splitmix64(3000 + b)), and
#   (2) golden per-iteration traces under tests/golden/julia_*.json, which tests/test_julia_golden.py consumes when present
#       (they pin the pieces of the oracle no committed reference artefact pins: the quadrotor model, regularisation restarts,
#       the square-root pass inside a full solve).
#
#   JULIA_NUM_THREADS=16 julia --project=/path/to/TrajectoryOptimization.jl baseline/run_reference.jl [B] [golden_dir]
#
# Thread safety: the reference's model closures share mutable scratch (src/model.jl:461-466,492-503) and eval counters
# (:59-60,173), so every thread builds its OWN Model / Problem from the dynamics function; nothing is shared between threads.
using TrajectoryOptimization
using LinearAlgebra
using Printf
const TO = TrajectoryOptimization

# ---- synthetic inputs: identical to trajopt_b200/problems.py (SURVEY 8d) ---------------------------------------------------
function splitmix_uniform(seed::Integer, count::Integer)
    out = zeros(Float64, count)
    for i = 1:count
        z = UInt64(seed) + UInt64(i) * 0x9E3779B97F4A7C15          # wraps modulo 2^64
        z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
        z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
        z = z ⊻ (z >> 31)
        out[i] = Float64(z >> 11) * (1.0 / 9007199254740992.0)
    end
    out
end
u_(r, a, b) = a .+ (b - a) .* r

function quadrotor_x0(b::Integer)
    r = splitmix_uniform(3000 + b, 13)
    x0 = zeros(13)
    x0[1:3] = [0.0, 0.0, 10.0] .+ u_(r[1:3], -2, 2)
    q = [1.0, 0, 0, 0] .+ 0.1 .* u_(r[4:7], -1, 1)
    x0[4:7] = q ./ norm(q)
    x0[8:13] = u_(r[8:13], -0.5, 0.5)
    x0
end

# ---- problems/quadrotor.jl, rebuilt per call so that no closure is shared between threads ----------------------------------
function quadrotor_problem(x0::Vector{Float64})
    model = Model(TO.Dynamics.quadrotor_dynamics!, 13, 4, TO.Dynamics.quad_params)
    model_d = rk3(model)
    n, m, N = 13, 4, 101
    q0 = [1.0; 0.0; 0.0; 0.0]
    xf = zeros(n); xf[1:3] = [0.0; 60.0; 10.0]; xf[4:7] = q0
    Q = Matrix(1.0e-3 * I, n, n); Q[4:7, 4:7] = Matrix(1.0e-2 * I, 4, 4)
    R = Matrix(1.0e-4 * I, m, m)
    Qf = Matrix(1000.0 * I, n, n)
    bnd3 = BoundConstraint(n, m, u_min=0.0)
    xU = copy(xf); xL = copy(xf)
    xU[4:7] .= Inf; xL[4:7] .= -Inf
    bnd_xf = BoundConstraint(n, m, x_min=xL, x_max=xU)
    dt = 5.0 / (N - 1)
    obj = LQRObjective(Diagonal(Q), Diagonal(R), Diagonal(Qf), xf, N)
    prob = Problem(model_d, obj, x0=x0, xf=xf, N=N, dt=dt)
    initial_controls!(prob, [0.5 * 9.81 / 4.0 * ones(m) for k = 1:N-1])
    for k = 1:N-1
        prob.constraints[k] += bnd3
    end
    prob.constraints[N] += bnd_xf
    prob
end

# AL phase of benchmark/quadrotor_benchmarks.jl:12-34, projected Newton off (what bench.py measures)
function quadrotor_options(; projected_newton=false)
    opts_ilqr = iLQRSolverOptions{Float64}(verbose=false, iterations=300)
    opts_al = AugmentedLagrangianSolverOptions{Float64}(verbose=false, opts_uncon=opts_ilqr, iterations=40,
        cost_tolerance=1.0e-5, cost_tolerance_intermediate=1.0e-4, constraint_tolerance=1.0e-3,
        penalty_scaling=10.0, penalty_initial=1.0)
    ALTROSolverOptions{Float64}(verbose=false, opts_al=opts_al, R_inf=1.0e-8, resolve_feasible_problem=false,
        projected_newton=projected_newton, projected_newton_tolerance=1.0e-3)
end

# ---- tiny JSON writer (no JSON.jl dependency in the package's Manifest) ------------------------------------------------------
jnum(x::AbstractFloat) = isfinite(x) ? @sprintf("%.17g", x) : (isnan(x) ? "\"nan\"" : (x > 0 ? "\"inf\"" : "\"-inf\""))
jnum(x::Integer) = string(x)
jvec(v) = "[" * join((jnum(x) for x in v), ",") * "]"
jmat(V) = "[" * join((jvec(v) for v in V), ",") * "]"

function trace_json(b, prob, solver)
    st = solver.solver_al.stats
    su = solver.solver_al.stats_uncon   # one copy of the iLQR stats per outer iteration (augmented_lagrangian_methods.jl:89)
    inner = "[" * join(("{\"cost\":" * jvec(s[:cost]) * ",\"dJ\":" * jvec(s[:dJ]) * ",\"gradient\":" * jvec(s[:gradient]) *
                        ",\"iterations\":" * jnum(s[:iterations]) * "}" for s in su), ",") * "]"
    "{\"problem\":$(b),\"iterations_outer\":$(st[:iterations]),\"iterations_total\":$(st[:iterations_total])," *
    "\"iterations_inner\":" * jvec(st[:iterations_inner]) * ",\"cost\":" * jvec(st[:cost]) * ",\"c_max\":" * jvec(st[:c_max]) *
    ",\"penalty_max\":" * jvec(st[:penalty_max]) * ",\"inner\":" * inner *
    ",\"X\":" * jmat(prob.X) * ",\"U\":" * jmat(prob.U) * ",\"max_violation\":" * jnum(max_violation(prob)) * "}"
end

function main()
    B = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 512
    golden = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden")
    nt = Threads.nthreads()
    # warm-up (compilation) on every thread
    Threads.@threads for t = 1:nt
        p = quadrotor_problem(quadrotor_x0(t - 1))
        solve!(p, quadrotor_options())
    end
    iters = zeros(Int, B)
    t0 = time()
    Threads.@threads for b = 0:B-1
        p = quadrotor_problem(quadrotor_x0(b))
        s = solve!(p, quadrotor_options())
        iters[b+1] = s.solver_al.stats[:iterations_total]
    end
    dt = time() - t0
    @printf("{\"impl\": \"reference-julia\", \"metric\": \"ALTRO solves/s (batched quadrotor N=101)\", \"value\": %.4f, \"unit\": \"solves/s\", \"threads\": %d, \"batch\": %d, \"seconds\": %.3f, \"ilqr_iters_per_s\": %.2f, \"julia\": \"%s\"}\n",
            B / dt, nt, B, dt, sum(iters) / dt, string(VERSION))
    # ---- golden traces: problems 0-7 and 6324-6331 of the synthetic batch (6326 is the REG_DIVERGED one: it never returns in the
    # reference -- backward_pass.jl:52-63 restarts forever -- so it is skipped here and documented in DESIGN.md section 2)
    mkpath(golden)
    open(joinpath(golden, "julia_quadrotor_traces.json"), "w") do f
        rows = String[]
        for b in vcat(0:7, [6324, 6325, 6327, 6328, 6329, 6330, 6331])
            p = quadrotor_problem(quadrotor_x0(b))
            s = solve!(p, quadrotor_options())
            push!(rows, trace_json(b, p, s))
        end
        write(f, "{\"source\": \"baseline/run_reference.jl\", \"julia\": \"$(VERSION)\", \"traces\": [" * join(rows, ",\n") * "]}\n")
    end
    # ---- square-root backward pass inside a full solve (test/sqrt_bp_tests.jl only checks one pass): pendulum ALTRO, square_root=true
    open(joinpath(golden, "julia_sqrt_traces.json"), "w") do f
        rows = String[]
        p = copy(TO.Problems.pendulum)
        o = quadrotor_options()
        o.opts_al.opts_uncon.square_root = true
        s = solve!(p, o)
        push!(rows, trace_json(0, p, s))
        write(f, "{\"source\": \"baseline/run_reference.jl\", \"case\": \"Problems.pendulum, ALTRO, square_root=true\", \"traces\": [" * join(rows, ",\n") * "]}\n")
    end
    # ---- projected-Newton polish (altro_methods.jl:31-39): final feasibility of problems 0-3 with projected_newton=true
    open(joinpath(golden, "julia_quadrotor_pn.json"), "w") do f
        rows = String[]
        for b in 0:3
            p = quadrotor_problem(quadrotor_x0(b))
            s = solve!(p, quadrotor_options(projected_newton=true))
            push!(rows, "{\"problem\":$(b),\"max_violation\":" * jnum(max_violation(p)) * ",\"cost\":" * jnum(TO.cost(p)) *
                        ",\"X\":" * jmat(p.X) * ",\"U\":" * jmat(p.U) * "}")
        end
        write(f, "{\"source\": \"baseline/run_reference.jl\", \"traces\": [" * join(rows, ",\n") * "]}\n")
    end
end

main()
