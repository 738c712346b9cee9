"""Known-answer vectors from the reference's own unit tests, replayed on the CPU oracle."""
import numpy as np
import pytest

from trajopt_b200 import abi, api, problems


def _car_problem(cons_stage=None, cons_term=None, N=5, X0=None, tf=float("nan"), dt=0.1):
    model = api.rk3(api.Dynamics.car)
    n, m = 3, 2
    Q, R, Qf = np.diag([1.0, 2.0, 3.0]), np.diag([0.5, 0.25]), np.diag([10.0, 20.0, 30.0])
    obj = api.LQRObjective(Q, R, Qf, np.array([0.0, 1.0, 0.0]), N)
    cons = api.Constraints(N)
    for k in range(N - 1):
        for c in (cons_stage or []):
            cons.add(k, c)
    for c in (cons_term or []):
        cons.add(N - 1, c)
    kw = dict(dt=dt) if np.isnan(tf) else dict(tf=tf, dt=dt)
    return api.Problem(model, obj, constraints=cons, x0=np.zeros(n), xf=np.array([0.0, 1.0, 0.0]), N=N, X0=X0, **kw)


def test_bound_constraint_values_and_jacobian(oracle):
    """test/constraint_tests.jl:91-108"""
    n, m = 3, 2
    x, u = np.array([1.0, 2.0, 3.0]), np.array([-5.0, 5.0])
    bnd = api.BoundConstraint(n, m, x_max=[5, 5, np.inf], x_min=[-10, -5, 0.0], u_max=0.0, u_min=-10.0)
    prob = _car_problem([bnd], [bnd])
    sp = oracle.Spec(prob)
    c, jac, eq = sp.constraints(0, x, u)
    assert c.tolist() == [-4, -3, -5, 5, -11, -7, -3, -5, -15]
    eye = np.eye(n + m)
    assert np.array_equal(jac, np.vstack([eye[[0, 1, 3, 4]], -eye]))
    assert not eq.any()
    cN, jacN, _ = sp.constraints(prob.N - 1, x)
    assert cN.tolist() == [-4, -3, -11, -7, -3]
    assert np.array_equal(jacN[:, :n], np.vstack([np.eye(n)[[0, 1]], -np.eye(n)]))
    assert prob.constraints.num_constraints() == [9, 9, 9, 9, 5]


def test_infeasible_constraint_and_slack_model(oracle):
    """test/constraint_tests.jl:189-202 and test/modified_model_tests.jl:107-128"""
    n, m = 3, 2
    x, u = np.array([1.0, 2.0, 3.0]), np.array([-5.0, 5.0])
    bnd = api.BoundConstraint(n, m, u_max=0.0, u_min=-10.0)
    X0 = np.zeros((5, 3))
    prob = _car_problem([bnd], [api.goal_constraint([0.0, 1.0, 0.0])], X0=X0)
    sp0 = oracle.Spec(prob)
    sp = oracle.Spec(prob, infeasible=True)
    assert (sp.n, sp.mc) == (n, m + n)
    u_inf = np.concatenate([u, [5.0, -5.0, 10.0]])
    c, jac, eq = sp.constraints(1, x, u_inf)
    c0, _, _ = sp0.constraints(1, x, u)
    assert c.tolist() == c0.tolist() + [5, -5, 10]          # v_inf == [v_stage; 5; -5; 10]
    assert eq.tolist() == [0] * len(c0) + [1, 1, 1]
    assert np.array_equal(jac[-3:, n + m:], np.eye(n)) and not jac[-3:, : n + m].any()
    # slack model: x_inf == x_d + u_slack, Jacobian block == I, inner blocks unchanged
    xn0, A0, B0 = sp0.dynamics(x, u)
    xn, A, B = sp.dynamics(x, u_inf)
    assert np.array_equal(xn, xn0 + u_inf[m:])
    assert np.array_equal(A, A0) and np.array_equal(B[:, :m], B0) and np.array_equal(B[:, m:], np.eye(n))


def test_min_time_model_and_cost(oracle):
    """test/modified_model_tests.jl:131-158, test/cost_tests.jl:167-182 (MinTimeCost expansion)."""
    n, m = 3, 2
    bnd = api.BoundConstraint(n, m, u_max=2.0, u_min=-2.0)
    prob = _car_problem([bnd], [api.goal_constraint([0.0, 1.0, 0.0])], tf=0.0, dt=0.1)
    opts = api.ALTROSolverOptions(R_minimum_time=15.0, dt_max=0.2, dt_min=1e-3)
    sp = oracle.Spec(prob, min_time=True, opts=opts)
    sp0 = oracle.Spec(prob)
    assert (sp.n, sp.mc) == (n + 1, m + 1)
    x, u, h = np.array([0.3, -0.2, 0.7, 0.31]), np.array([0.5, -0.4, 0.25]), 0.25
    xn, A, B = sp.dynamics(x, u)
    assert xn[-1] == u[-1]                                   # x⁺[end] == u[end]
    m0 = api.Marshalled(prob)
    Z = oracle.discrete_jacobian(abi.MODEL_CAR, abi.INTEG_RK3, x[:n], u[:m], h * h)
    assert np.array_equal(xn[:n], oracle.discrete(abi.MODEL_CAR, abi.INTEG_RK3, x[:n], u[:m], h * h))
    assert np.array_equal(A[:n, :n], Z[:, :n]) and not A[n].any() and not A[:, n].any()
    assert np.array_equal(B[:n, :m], Z[:, n:n + m])
    assert np.array_equal(B[:n, m], Z[:, n + m] * (2 * h)) and B[n, m] == 1.0 and not B[n, :m].any()
    # constraints: [u_max(2), sqrt(dt_max), u_min(2), sqrt(dt_min)] then the min-time equality for 1<k<N
    c1, _, eq1 = sp.constraints(0, x, u)
    c2, jac2, eq2 = sp.constraints(1, x, u)
    assert len(c1) == 6 and len(c2) == 7 and eq2.tolist() == [0] * 6 + [1]
    assert c2[2] == u[2] - np.sqrt(0.2) and c2[5] == np.sqrt(1e-3) - u[2] and c2[6] == u[2] - x[3]
    assert jac2[6, n] == -1.0 and jac2[6, n + 1 + m] == 1.0
    # cost: ℓ(x,u)τ² + R τ² ; expansion entries as written in minimum_time.jl:161-191 (Q12)
    l1 = sp0.stage_cost(x[:n], u[:m]) / prob.dt
    assert np.isclose(sp.stage_cost(x, u), l1 * h * h + 15.0 * h * h, rtol=1e-15)
    Qx, Qu, Qxx, Quu, Qux = sp.stage_expansion(x, u)
    assert np.isclose(Qu[-1], h * (2 * l1 + 15.0), rtol=1e-15) and np.isclose(Quu[-1, -1], 2 * l1 + 15.0, rtol=1e-15)
    assert Qx[-1] == 15.0 * x[-1] and Qxx[-1, -1] == 15.0
    st = prob.obj.stage
    Qu_in = st.R @ u[:m] + st.r + st.H @ x[:n]
    assert np.allclose(Quu[:m, -1], 2 * h * Qu_in, rtol=1e-14) and np.allclose(Quu[-1, :m], 2 * h * Qu_in, rtol=1e-14)
    assert np.allclose(Qxx[:n, :n], st.Q * h * h, rtol=1e-15)


def test_circle_sphere_pos(oracle):
    """test/test_utils.jl:3-5,82-94"""
    prob = _car_problem([api.CircleConstraints([(1.0, 0.0, 1.0), (1.0, 0.0, 0.5)]),
                         api.SphereConstraints([(1.0, 0.0, 0.0, 1.0), (1.0, 0.0, 0.0, 0.5)])])
    sp = oracle.Spec(prob)
    c, jac, _ = sp.constraints(0, np.zeros(3), np.zeros(2))
    assert c[0] == 0 and c[1] < 0 and c[2] == 0 and c[3] < 0
    c, jac, _ = sp.constraints(0, np.array([0.75, 0, 0]), np.zeros(2))
    assert c[1] > 0 and c[3] > 0
    assert jac[1].tolist() == [-(2 * (0.75 - 1.0)), -0.0, 0, 0, 0]
    assert max(0, 10) == 10 and max(0, -1e-4) == 0  # pos()


def test_cost_identities(oracle):
    """test/cost_tests.jl:64-89: stage cost, LQR terminal, expansions."""
    rng = np.random.default_rng(7)
    prob = _car_problem()
    sp = oracle.Spec(prob)
    st, tm = prob.obj.stage, prob.obj.terminal
    x, u, dt = rng.random(3), rng.random(2), prob.dt
    xf = prob.xf
    ref = (0.5 * x @ st.Q @ x + 0.5 * u @ st.R @ u + st.q @ x + st.c) * dt
    assert np.isclose(sp.stage_cost(x, u), ref, rtol=1e-14)
    assert np.isclose(sp.term_cost(x), 0.5 * (x - xf) @ tm.Q @ (x - xf), rtol=1e-13)
    Qx, Qu, Qxx, Quu, Qux = sp.stage_expansion(x, u)
    assert np.allclose(Qx, (st.Q @ x + st.q) * dt, rtol=1e-15) and np.array_equal(Qxx, st.Q * dt)
    assert np.allclose(Qu, st.R @ u * dt, rtol=1e-15) and np.array_equal(Quu, st.R * dt) and not Qux.any()
    Sx, Sxx = sp.term_expansion(x)
    assert np.allclose(Sx, tm.Q @ (x - xf), rtol=1e-14) and np.array_equal(Sxx, tm.Q)


@pytest.mark.parametrize("model", range(7))
def test_dual_jacobian_matches_finite_differences(oracle, model):
    """test/model_tests.jl:57-65,137-148: the ForwardDiff Jacobian of the discrete model."""
    n, m = abi.MODEL_DIMS[model]
    rng = np.random.default_rng(model)
    x, u, dt = rng.normal(size=n) * 0.5, rng.normal(size=m) * 0.5 + 1.0, 0.05
    if model == abi.MODEL_QUADROTOR:
        x[3:7] = [0.9, 0.1, -0.2, 0.3]
    Z = oracle.discrete_jacobian(model, abi.INTEG_RK3, x, u, dt)
    s = np.concatenate([x, u, [dt]])
    f = lambda s: oracle.discrete(model, abi.INTEG_RK3, s[:n], s[n:n + m], s[-1])
    for j in range(n + m + 1):
        e = np.zeros_like(s)
        e[j] = 1e-6
        fd = (f(s + e) - f(s - e)) / 2e-6
        assert np.allclose(Z[:, j], fd, rtol=2e-6, atol=2e-7), (model, j)


def test_rk3_matches_hand_formula(oracle):
    """src/integration.jl:149-158 on the double integrator (exact small integers)."""
    x, u, dt = np.array([1.0, 2.0]), np.array([3.0]), 0.5
    f = lambda x: np.array([x[1], u[0]])
    k1 = f(x) * dt
    k2 = f(x + k1 / 2) * dt
    k3 = f(x - k1 + 2 * k2) * dt
    assert np.array_equal(oracle.discrete(0, abi.INTEG_RK3, x, u, dt), x + (k1 + 4 * k2 + k3) / 6)


def test_sqrt_backward_pass_equals_standard(oracle):
    """test/sqrt_bp_tests.jl:30-44: ΔV, K, d, S.x agree and S.xx == S√'S√ (isapprox, rtol √eps)."""
    import ctypes as C
    rng = np.random.default_rng(3)
    n, m, N = 3, 2, 12
    A = np.stack([np.eye(n) + 0.1 * rng.normal(size=(n, n)) for _ in range(N - 1)])
    B = np.stack([0.3 * rng.normal(size=(n, m)) for _ in range(N - 1)])
    Qx, Qu = rng.normal(size=(N - 1, n)), rng.normal(size=(N - 1, m))
    def spd(k):
        M = rng.normal(size=(k, k))
        return M @ M.T + k * np.eye(k)
    Qxx, Quu = np.stack([spd(n) for _ in range(N - 1)]), np.stack([spd(m) for _ in range(N - 1)])
    Qux = np.zeros((N - 1, m, n))
    QNx, QNxx = rng.normal(size=n), spd(n)
    cm = lambda M: np.ascontiguousarray(np.swapaxes(M, -1, -2))  # to column-major blocks
    outs = []
    for sq in (0, 1):
        K, d = np.zeros((N - 1, n, m)), np.zeros((N - 1, m))
        S1, s1, dV = np.zeros((n, n)), np.zeros(n), np.zeros(2)
        args = [cm(A), cm(B), Qx.copy(), Qu.copy(), cm(Qxx), cm(Quu), cm(Qux), QNx.copy(), cm(QNxx)]
        rc = oracle.lib().oracle_backwardpass(n, m, N, sq, 0.0, *[a.ctypes.data for a in args], K.ctypes.data, d.ctypes.data,
                                              S1.ctypes.data, s1.ctypes.data, dV.ctypes.data)
        assert rc == 0
        outs.append((np.swapaxes(K, 1, 2), d, S1.T, s1, dV))
    (K0, d0, S0, s0, dV0), (K1, d1, Sq, s1, dV1) = outs
    tol = dict(rtol=1.5e-8, atol=1e-10)
    assert np.allclose(K0, K1, **tol) and np.allclose(d0, d1, **tol) and np.allclose(dV0, dV1, **tol)
    assert np.allclose(s0, s1, **tol) and np.allclose(S0, Sq.T @ Sq, **tol)


def test_sincos_within_one_ulp(oracle):
    import math
    xs = np.concatenate([np.linspace(-8, 8, 4001), np.random.default_rng(0).uniform(-1e5, 1e5, 4000), [0.0, 1e-9, -1e-9]])
    for x in xs:
        s, c = oracle.sincos(float(x))
        assert abs(s - math.sin(x)) <= np.spacing(abs(math.sin(x))) and abs(c - math.cos(x)) <= np.spacing(abs(math.cos(x)))
    assert all(np.isnan(v) for v in oracle.sincos(float("inf")))


def test_oracle_reports_status_instead_of_hanging_or_throwing(oracle):
    """(a) problem 6326 of the synthetic quadrotor batch: the reference's in-place Q accumulation blows
    Quu up to NaN and its backward pass would restart forever; the restatement stops with TO_STATUS_REG_DIVERGED once
    rho is no longer finite.  (b) square-root + minimum time: cost_expansion_sqrt! hits a non-PD stage Hessian
     -> TO_STATUS_NOT_PD_SQRT, every other problem of the batch unaffected."""
    from cases import CASES
    prob, opts, x0, X0 = CASES["quad_regdiv"](4)
    r = oracle.solve(prob, opts, x0=x0, X0=X0, B=4, inner_cap=0, outer_cap=0)
    assert r["results"]["status"].tolist() == [0, 0, 32, 0]
    prob, opts, x0, X0 = CASES["acrobot_sqrt_mintime"](4)
    r = oracle.solve(prob, opts, x0=x0, X0=X0, B=4, inner_cap=0, outer_cap=0)
    assert np.all(r["results"]["status"] == 4)
    prob, opts, x0, X0 = CASES["pend_sqrt_altro"](4)
    r = oracle.solve(prob, opts, x0=x0, X0=X0, B=4, inner_cap=0, outer_cap=0)
    assert np.all(r["results"]["status"] == 0) and np.all(r["results"]["c_max"] < 1e-3)


def test_reference_minimum_time_test_inequalities(oracle):
    """test/minimum_time_tests.jl:36-63 replayed on the oracle: the minimum-time pendulum swing-up takes less than half
    the fixed-time solution and less than 1 s, reaches the goal and satisfies the constraints.  This is the only
    reference artefact that pins the minimum-time path."""
    import trajopt_b200 as to
    from helpers import pendulum_mintime_test
    make, o = pendulum_mintime_test(to)
    p = make(0.15)
    r = oracle.solve(p, o, x0=p.x0[None], B=1)
    tt = r["dts"].sum()
    assert abs(tt - 4.5) < 1e-12 and r["results"]["c_max"][0] < 1e-3
    p_mt = make(0.075, tf=0.0, U0=r["U"][0])
    r2 = oracle.solve(p_mt, o, x0=p_mt.x0[None], B=1)
    tt_mt = r2["dts"].sum()
    assert tt_mt < 0.5 * tt and tt_mt < 1.0
    assert np.abs(r2["X"][0, -1] - p_mt.xf).max() < 1e-3 and r2["results"]["c_max"][0] < 1e-3 and r2["results"]["status"][0] == 0


def test_reference_quadrotor_test_inequalities(oracle):
    """test/quadrotor_tests.jl:1-60 replayed on the oracle (rk4 quadrotor, 50 m flight): iLQR reaches xf to 5e-3;
    AL with the goal constraint and AL with goal + control bounds meet the constraint tolerance."""
    import trajopt_b200 as to
    from helpers import quadrotor_test_problem
    p, il, al = quadrotor_test_problem(to, "none")
    r = oracle.solve(p, il, x0=p.x0[None], B=1)
    assert np.linalg.norm(r["X"][0, -1] - p.xf) < 5e-3
    p, il, al = quadrotor_test_problem(to, "goal")
    r = oracle.solve(p, al, x0=p.x0[None], B=1)
    assert np.abs(r["X"][0, -1] - p.xf).max() < 1e-3 and r["results"]["c_max"][0] < 1e-3
    p, il, al = quadrotor_test_problem(to, "goal+bounds")
    r = oracle.solve(p, al, x0=p.x0[None], B=1)
    assert np.linalg.norm(r["X"][0, -1] - p.xf) < 1e-3 and r["results"]["c_max"][0] < 1e-3
    assert r["U"].min() > -1e-3 and r["U"].max() < 15.0 + 1e-3


def test_reference_pendulum_integrator_sweep(oracle):
    """test/pendulum_tests.jl:9-27 (explicit schemes): ALTRO on the pendulum with midpoint / rk3 / rk4 meets the
    constraint tolerance."""
    from cases import CASES
    for name in ("pend_altro", "pend_rk4_altro", "pend_midpoint_altro"):
        prob, opts, x0, X0 = CASES[name](1)
        r = oracle.solve(prob, opts, x0=prob.x0[None], B=1)
        assert r["results"]["c_max"][0] < 1e-3 and r["results"]["status"][0] == 0, name
