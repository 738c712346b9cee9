"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA engine, called through the C ABI,
against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): iteration counts, status and active-set masks BIT-EXACT; per-iteration
cost J, final X/U and constraint violation within 1e-8 relative.  The engine and the oracle share one
arithmetic contract (sequential FMA chains for inner products, no other contraction, a common
sin/cos), so in practice every float below is bit-identical too — the tests assert the stated
tolerance and additionally report exact equality for the integer observables.
"""
import numpy as np
import pytest

from cases import CASES

pytestmark = pytest.mark.gpu
RTOL = 1e-8  # north_star tolerance for floating-point observables


def _solve_gpu(to, prob, opts, x0, X0, B, inner_cap=2048, outer_cap=96):
    bs = to.api.BatchSolver(prob, B, 0, inner_cap, outer_cap)
    try:
        bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), X0)
        bs.solve(opts)
        X, U, dts = bs.solution()
        inner, outer = bs.trace()
        lam, mu, act = bs.duals()
        return dict(results=bs.results(), X=X, U=U, dts=dts, inner=inner, outer=outer, lam=lam, mu=mu, act=act,
                    ms=bs.kernel_ms(), launches=bs.launches())
    finally:
        bs.close()


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_nan | both_inf | (np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + 1e-300)
    return bool(np.all(ok))


def _compare(ref, gpu, B, atol_xu=1e-10):
    for f in ("iterations_total", "iterations_outer", "status", "steps"):
        assert np.array_equal(ref["results"][f], gpu["results"][f]), (f, ref["results"][f], gpu["results"][f])
    assert _close(ref["results"]["J"], gpu["results"]["J"]) and _close(ref["results"]["c_max"], gpu["results"]["c_max"])
    fin = np.isfinite(ref["X"]).all(axis=(1, 2))
    assert np.allclose(ref["X"][fin], gpu["X"][fin], rtol=RTOL, atol=atol_xu)
    assert np.allclose(ref["U"][fin], gpu["U"][fin], rtol=RTOL, atol=atol_xu)
    assert np.allclose(ref["dts"][fin], gpu["dts"][fin], rtol=RTOL, atol=0)
    for b in range(B):
        ri, gi = ref["inner"][b], gpu["inner"][b]
        assert len(ri) == len(gi)
        for f in ("iter", "outer"):
            assert np.array_equal(ri[f], gi[f])
        assert np.array_equal(ri["alpha"], gi["alpha"])          # accepted step sizes: bit-exact
        assert np.array_equal(ri["rho"] > 0, gi["rho"] > 0)
        for f in ("cost", "dJ", "gradient", "expected", "z", "rho"):
            assert _close(ri[f], gi[f], 1e-8 if f in ("cost", "rho") else 1e-6), (b, f)
        ro, go = ref["outer"][b], gpu["outer"][b]
        assert len(ro) == len(go) and np.array_equal(ro["iterations_inner"], go["iterations_inner"])
        assert _close(ro["cost"], go["cost"]) and _close(ro["c_max"], go["c_max"]) and _close(ro["penalty_max"], go["penalty_max"])


@pytest.mark.parametrize("name", [c for c in CASES if c not in ("escape_altro",)])
def test_parity_small_batches(to, oracle, name):
    """every parity case on the default engine.  Few live problems: after the init kernel the CTA-per-problem resident kernel
    (resident.cuh) runs every problem to completion (the square-root cases stay on the lockstep tick)"""
    B = 8
    prob, opts, x0, X0 = CASES[name](B)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    gpu = _solve_gpu(to, prob, opts, x0, X0, B)
    _compare(ref, gpu, B)


@pytest.mark.parametrize("name", [c for c in CASES if c not in ("escape_altro",)])
def test_parity_lockstep_tail(to, oracle, name, monkeypatch):
    """the lockstep tick in tail mode (all step sizes in one launch, CTA-per-problem backward pass) with the resident kernel off"""
    monkeypatch.setenv("TRAJOPT_B200_RESIDENT_THRESHOLD", "0")
    B = 8
    prob, opts, x0, X0 = CASES[name](B)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    gpu = _solve_gpu(to, prob, opts, x0, X0, B)
    _compare(ref, gpu, B)


@pytest.mark.parametrize("name", ["quad_altro", "quad_regdiv", "cart_altro", "escape_notebook", "park_inf_altro", "pend_mintime",
                                  "acrobot_sqrt_al"])
def test_parity_grouped_line_search(to, oracle, name, monkeypatch):
    """the same engine with the bulk kernels of large batches forced on: step sizes tried 8 at a time with retry lists, and the
    lane-group backward pass (ls_bp_kernel) instead of the CTA-per-problem one that small batches get by default"""
    monkeypatch.setenv("TRAJOPT_B200_TAIL_THRESHOLD", "0")
    monkeypatch.setenv("TRAJOPT_B200_BP_CTA_THRESHOLD", "0")
    monkeypatch.setenv("TRAJOPT_B200_RESIDENT_THRESHOLD", "0")
    B = 8
    prob, opts, x0, X0 = CASES[name](B)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    gpu = _solve_gpu(to, prob, opts, x0, X0, B)
    _compare(ref, gpu, B)


@pytest.mark.parametrize("name", [c for c in CASES if c not in ("escape_altro",)])
def test_parity_lane_group_backward_pass(to, oracle, name, monkeypatch):
    """every case through ls_bp_kernel (16 lanes per problem, the bulk backward pass) with the tail-mode line search"""
    monkeypatch.setenv("TRAJOPT_B200_BP_CTA_THRESHOLD", "0")
    monkeypatch.setenv("TRAJOPT_B200_RESIDENT_THRESHOLD", "0")
    B = 8
    prob, opts, x0, X0 = CASES[name](B)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    gpu = _solve_gpu(to, prob, opts, x0, X0, B)
    _compare(ref, gpu, B)


def test_backward_pass_kernels_agree_bitwise(to, monkeypatch):
    """the two backward-pass kernels (lane group per problem / CTA per problem with the knot-parallel expansion) and the
    restart hand-over between them, on a batch that has regularisation restarts and a REG_DIVERGED problem: identical result
    records, X and U"""
    B = 256
    prob, opts, x0, _ = CASES["quad_altro"](B)
    x0 = x0.copy()
    pr, _, xr, _ = CASES["quad_regdiv"](8)
    x0[:8] = xr
    monkeypatch.setenv("TRAJOPT_B200_RESIDENT_THRESHOLD", "0")
    monkeypatch.setenv("TRAJOPT_B200_BP_CTA_THRESHOLD", "0")
    monkeypatch.setenv("TRAJOPT_B200_BP_DEFER_RESTARTS", "0")   # lane-group kernel serves every restart itself
    a = _solve_gpu(to, prob, opts, x0, None, B, inner_cap=0, outer_cap=0)
    monkeypatch.setenv("TRAJOPT_B200_BP_DEFER_RESTARTS", "1")   # ... hands every regularisation increase to the CTA kernel
    monkeypatch.setenv("TRAJOPT_B200_BP_INLINE_RESTARTS", "0")
    c = _solve_gpu(to, prob, opts, x0, None, B, inner_cap=0, outer_cap=0)
    monkeypatch.setenv("TRAJOPT_B200_BP_CTA_THRESHOLD", "100000000")
    b = _solve_gpu(to, prob, opts, x0, None, B, inner_cap=0, outer_cap=0)
    for other in (b, c):
        assert a["results"].tobytes() == other["results"].tobytes()
        assert np.array_equal(a["X"], other["X"], equal_nan=True) and np.array_equal(a["U"], other["U"], equal_nan=True)


def test_split_line_search_agrees_bitwise(to, oracle, monkeypatch):
    """car_escape (177 constraint rows per knot) on the lockstep tick in tail mode: the cost evaluated ON the state chain (one
    kernel) and OFF it (chains, then the costs of all (knot, step size), then the pick; resident.cuh) -- identical records, X, U and
    iteration traces, and equal to the oracle"""
    monkeypatch.setenv("TRAJOPT_B200_RESIDENT_THRESHOLD", "0")
    B = 4
    prob, opts, x0, X0 = CASES["escape_notebook"](B)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    split = _solve_gpu(to, prob, opts, x0, X0, B)
    monkeypatch.setenv("TRAJOPT_B200_SPLIT_LINESEARCH", "0")
    fused = _solve_gpu(to, prob, opts, x0, X0, B)
    assert split["launches"] > fused["launches"]          # the split path really ran (two more kernels per tick)
    assert split["results"].tobytes() == fused["results"].tobytes()
    assert np.array_equal(split["X"], fused["X"], equal_nan=True) and np.array_equal(split["U"], fused["U"], equal_nan=True)
    for b in range(B):
        for f in ("cost", "dJ", "gradient", "expected", "z", "rho", "alpha"):
            assert np.array_equal(split["inner"][b][f], fused["inner"][b][f], equal_nan=True), (b, f)
    _compare(ref, split, B)


@pytest.mark.parametrize("name", ["pend_sqrt_altro", "dp_sqrt_ilqr", "acrobot_sqrt_al", "quad_sqrt_ilqr", "acrobot_sqrt_mintime",
                                  "dp_sqrt_mintime"])
def test_square_root_pass_kernels_agree_bitwise(to, oracle, name, monkeypatch):
    """the two kernels of the square-root backward pass (sqrt_bp.cuh): thread per problem (full batches, and the small models at
    any size) and warp per problem (few live problems of a model too large for one thread's registers) -- the same records, X and
    U bit for bit, and both equal to the oracle"""
    B = 8
    prob, opts, x0, X0 = CASES[name](B)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    monkeypatch.setenv("TRAJOPT_B200_SQRT_WARP_THRESHOLD", "1000000")
    warp = _solve_gpu(to, prob, opts, x0, X0, B)
    monkeypatch.setenv("TRAJOPT_B200_SQRT_WARP_THRESHOLD", "0")
    thread = _solve_gpu(to, prob, opts, x0, X0, B)
    assert warp["results"].tobytes() == thread["results"].tobytes()
    assert np.array_equal(warp["X"], thread["X"], equal_nan=True) and np.array_equal(warp["U"], thread["U"], equal_nan=True)
    for b in range(B):
        for f in ("cost", "dJ", "gradient", "expected", "z", "rho", "alpha"):
            assert np.array_equal(warp["inner"][b][f], thread["inner"][b][f], equal_nan=True), (b, f)
    _compare(ref, thread, B)


@pytest.mark.parametrize("name", ["di_altro", "quad_altro", "quad_regdiv", "cart_ilqr", "escape_notebook", "park_inf_altro", "pend_mintime"])
def test_parity_persistent_engine(to, oracle, name, monkeypatch):
    """the second, independent CUDA implementation (one warp-resident kernel per solve, engine.cuh) against the same oracle"""
    monkeypatch.setenv("TRAJOPT_B200_ENGINE", "persistent")
    B = 8
    prob, opts, x0, X0 = CASES[name](B)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    gpu = _solve_gpu(to, prob, opts, x0, X0, B)
    _compare(ref, gpu, B)


def test_status_bits_match_the_oracle(to, oracle):
    """per-problem status instead of exceptions: REG_DIVERGED (problem 6326 of the synthetic quadrotor batch) and
    NOT_PD_SQRT (square-root + minimum time) are reported identically by oracle and engine"""
    prob, opts, x0, X0 = CASES["quad_regdiv"](8)
    gpu = _solve_gpu(to, prob, opts, x0, X0, 8, inner_cap=0, outer_cap=0)
    assert gpu["results"]["status"][2] == 32 and np.count_nonzero(gpu["results"]["status"]) == 1
    prob, opts, x0, X0 = CASES["acrobot_sqrt_mintime"](8)
    gpu = _solve_gpu(to, prob, opts, x0, X0, 8, inner_cap=0, outer_cap=0)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=8, inner_cap=0, outer_cap=0)
    assert np.array_equal(gpu["results"]["status"], ref["results"]["status"]) and np.all(gpu["results"]["status"] & 4)


def test_parity_car_escape_170_circles(to, oracle):
    B = 2  # 177 constraint rows per knot: slow on one warp, keep it small
    prob, opts, x0, X0 = CASES["escape_altro"](B)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    gpu = _solve_gpu(to, prob, opts, x0, X0, B)
    _compare(ref, gpu, B)


def test_duals_and_active_set_bit_exact(to, oracle):
    """active-set masks bit-exact; λ, μ to 1e-8 (quadrotor ALTRO and the README block move)."""
    for name, B in (("quad_altro", 4), ("di_altro", 1), ("cart_altro", 4)):
        prob, opts, x0, X0 = CASES[name](B)
        ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=0, outer_cap=0, want_duals=True)
        gpu = _solve_gpu(to, prob, opts, x0, X0, B)
        assert ref["act"].shape == gpu["act"].shape and np.array_equal(ref["act"], gpu["act"])
        assert _close(ref["lam"], gpu["lam"]) and _close(ref["mu"], gpu["mu"])


def test_readme_block_move(to):
    """README.md:29-67 through the reference-style API: solve!(prob, ALTROSolverOptions{Float64}())."""
    prob = to.problems.doubleintegrator()
    solver = to.solve_b(prob, to.ALTROSolverOptions())
    assert solver.status == 0 and to.max_violation(prob) < 1e-3
    assert np.all(np.abs(prob.U) <= 1.5 + 1e-3) and np.linalg.norm(prob.X[-1] - prob.xf) < 1e-3
    assert solver.stats["iterations"] == len(solver.stats["cost"]) == len(solver.stats["c_max"])


def test_copying_solve_and_logger_tables(to):
    """solve(prob, opts) (src/solvers.jl:104-108) leaves the caller's problem alone; the verbose tables of src/logger.jl are
    rebuilt from the histories (one OuterLoop row per recorded outer iteration, one InnerLoop row per inner record)."""
    prob = to.problems.doubleintegrator()
    U_before, X_before = prob.U.copy(), prob.X.copy()
    solved, solver = to.solve(prob, to.ALTROSolverOptions())
    assert np.array_equal(prob.U, U_before) and np.array_equal(np.isnan(prob.X), np.isnan(X_before))
    ref = to.problems.doubleintegrator()
    to.solve_b(ref, to.ALTROSolverOptions())
    assert np.array_equal(solved.X, ref.X) and np.array_equal(solved.U, ref.U)
    log = to.format_log(solver)
    rows = [ln for ln in log.splitlines() if ln[:1].isdigit()]
    assert len(rows) == len(solver.stats["c_max"]) and "c_max" in log and "expected" in log
    inner_rows = [ln for ln in log.splitlines() if ln.startswith("    ") and ln.strip()[:1].isdigit()]
    assert len(inner_rows) == len(solver.inner)
    both, solvers = to.solve([prob, prob], to.ALTROSolverOptions())
    assert np.array_equal(both[0].X, solved.X) and np.array_equal(both[1].U, solved.U) and len(solvers) == 2


def test_reference_integration_inequalities(to):
    """The reference's own solve-level tests, re-expressed as batch-of-1 solves:
    test/quadrotor_tests.jl:1-60 (rk4; iLQR reaches xf; AL with goal; AL with goal + control bounds),
    test/car_tests.jl:31-32 (parallel park), test/minimum_time_tests.jl:38-46 (total time shrinks)."""
    from helpers import quadrotor_test_problem
    p, il, al = quadrotor_test_problem(to, "none")
    to.solve_b(p, il)
    assert np.linalg.norm(p.X[-1] - p.xf) < 5e-3
    p, il, al = quadrotor_test_problem(to, "goal")
    to.solve_b(p, al)
    assert np.abs(p.X[-1] - p.xf).max() < al.constraint_tolerance and to.max_violation(p) < al.constraint_tolerance
    p, il, al = quadrotor_test_problem(to, "goal+bounds")
    to.solve_b(p, al)
    assert np.linalg.norm(p.X[-1] - p.xf) < al.constraint_tolerance and to.max_violation(p) < al.constraint_tolerance
    p = to.problems.quadrotor()
    s = to.solve_b(p, to.problems.quadrotor_bench_options())
    assert to.max_violation(p) < 1e-3 and s.c_max < 1e-3
    p = to.problems.parallel_park()
    p.constraints = to.Constraints(p.N)
    to.solve_b(p, to.iLQRSolverOptions())
    assert np.linalg.norm(p.X[-1] - p.xf) < 1e-3
    # test/minimum_time_tests.jl:36-63: tt_mt < 0.5*tt, tt_mt < 1.0, goal reached, constraints satisfied
    from helpers import pendulum_mintime_test
    make, o = pendulum_mintime_test(to)
    p = make(0.15)
    to.solve_b(p, o)
    tt = to.total_time(p)
    p_mt = make(0.075, tf=0.0, U0=p.U)
    to.solve_b(p_mt, o)
    tt_mt = to.total_time(p_mt)
    assert tt_mt < 0.5 * tt and tt_mt < 1.0
    assert np.abs(p_mt.X[-1] - p_mt.xf).max() < 1e-3 and to.max_violation(p_mt) < o.opts_al.constraint_tolerance


def test_mpc_warm_start_shift(to, oracle):
    """to_warm_start_shift: the re-solve from the device-side shifted controls equals (bit for bit) a fresh solve that is handed
    the same shifted controls and predicted states through to_set_batch, matches the oracle on them, and needs fewer iLQR
    iterations than the cold solve did."""
    B = 16
    prob, opts, x0, _ = CASES["cart_altro"](B)
    N, n, m = prob.N, prob.model.n, prob.model.m
    bs = to.api.BatchSolver(prob, B, 0, 0, 0)
    try:
        bs.set_batch(x0, np.broadcast_to(prob.U, (B, N - 1, m)))
        bs.solve(opts)
        cold = bs.results().copy()
        X, U, _ = bs.solution()
        X, U = X.copy(), U.copy()
        bs.warm_start_shift(None, 1)
        bs.solve(opts)
        warm = bs.results().copy()
        Xw, Uw, _ = bs.solution()
        Xw, Uw = Xw.copy(), Uw.copy()
    finally:
        bs.close()
    U0 = np.concatenate([U[:, 1:], U[:, -1:]], axis=1)
    x1 = X[:, 1]
    fresh = to.api.BatchSolver(prob, B, 0, 0, 0)
    try:
        fresh.set_batch(x1, U0)
        fresh.solve(opts)
        rf = fresh.results().copy()
        Xf, Uf, _ = fresh.solution()
        assert rf.tobytes() == warm.tobytes() and np.array_equal(Xf, Xw) and np.array_equal(Uf, Uw)
    finally:
        fresh.close()
    p2 = prob.copy()
    ref = oracle.solve(p2, opts, x0=x1[:4], U0=U0[:4], B=4, inner_cap=0, outer_cap=0)
    assert ref["results"].tobytes() == warm[:4].tobytes() and np.array_equal(ref["X"], Xw[:4])
    assert warm["steps"].sum() < cold["steps"].sum()


def test_batched_solve_matches_singles_and_is_order_independent(to):
    """The new batched solve!: B problems at once == B single solves (queue order must not matter)."""
    B = 64
    prob, opts, x0, _ = CASES["cart_altro"](B)
    big = _solve_gpu(to, prob, opts, x0, None, B, inner_cap=0, outer_cap=0)
    perm = np.random.default_rng(0).permutation(B)
    shuf = _solve_gpu(to, prob, opts, x0[perm], None, B, inner_cap=0, outer_cap=0)
    assert big["results"][perm].tobytes() == shuf["results"].tobytes()
    assert np.array_equal(big["X"][perm], shuf["X"]) and np.array_equal(big["U"][perm], shuf["U"])
    one = _solve_gpu(to, prob, opts, x0[5:6], None, 1, inner_cap=0, outer_cap=0)
    assert one["results"].tobytes() == big["results"][5:6].tobytes() and np.array_equal(one["X"][0], big["X"][5])


def test_full_size_properties_quadrotor(to, oracle):
    """A larger batch (size-independent properties): every problem converges with c_max < tol, the
    returned trajectories are dynamically feasible under the oracle's rk3 (X[k+1] == fd(X[k],U[k])
    bit for bit), controls respect u >= -tol, and a sampled subset matches the oracle exactly."""
    B = 2048
    prob, opts, x0, _ = CASES["quad_altro"](B)
    gpu = _solve_gpu(to, prob, opts, x0, None, B, inner_cap=0, outer_cap=0)
    r = gpu["results"]
    assert np.all(r["status"] == 0) and np.all(r["c_max"] < 1e-3) and np.all(r["steps"] == r["iterations_total"] - r["iterations_outer"] + 1)
    assert gpu["U"].min() > -1e-3
    for b in (0, 777, 2047):
        for k in (0, 50, 99):
            xn = oracle.discrete(4, 0, gpu["X"][b, k], gpu["U"][b, k], prob.dt)
            assert np.array_equal(xn, gpu["X"][b, k + 1])
    idx = np.array([0, 1, 1023, 2047])
    ref = oracle.solve(prob, opts, x0=x0[idx], B=len(idx), inner_cap=0, outer_cap=0)
    assert ref["results"].tobytes() == r[idx].tobytes() and np.array_equal(ref["X"], gpu["X"][idx])


def test_errors_are_reported_not_thrown(to):
    import ctypes as C
    prob = to.problems.doubleintegrator()
    bs = to.api.BatchSolver(prob, 1, 0)
    try:
        with pytest.raises(RuntimeError, match="to_set_batch"):
            bs.solve(to.ALTROSolverOptions())
        bs.set_batch(prob.x0[None], prob.U[None])
        o = to.ALTROSolverOptions()
        o.opts_al.opts_uncon.iterations_linesearch = 40
        with pytest.raises(RuntimeError, match="iterations_linesearch"):
            bs.solve(o)
    finally:
        bs.close()
