"""Shared problem builders for the tests (the notebook variants are defined inline in the
reference's notebooks, not in problems/*.jl)."""
import numpy as np

import trajopt_b200 as to
from trajopt_b200 import api, problems


def acrobot_notebook():
    """examples/acrobot/Acrobot.ipynb cells 5-17"""
    p = problems.acrobot(N=251, dt=0.01, Qs=1e-3, Rs=1e-3, Qfs=1000.0)
    il = api.iLQRSolverOptions(cost_tolerance=1e-6)
    al = api.AugmentedLagrangianSolverOptions(opts_uncon=il, iterations=30, penalty_scaling=10.0, cost_tolerance=1e-6,
                                              cost_tolerance_intermediate=1e-5, constraint_tolerance=1e-4)
    return p, al


def car_escape_notebook():
    """examples/car/Car Escape.ipynb cells 3-23: trap at every k<N, no bounds, R_inf=1e-3."""
    model = api.rk3(api.Dynamics.car)
    n, m, N, tf = 3, 2, 101, 3.0
    x0, xf = np.array([2.5, 2.5, 0.0]), np.array([7.5, 2.5, 0.0])
    obj = api.LQRObjective(1e-3 * np.eye(n), 1e-2 * np.eye(m), 100 * np.eye(n), xf, N)
    cons = api.Constraints(N)
    trap = api.CircleConstraints(problems.escape_circles(), "trap")
    for k in range(N - 1):
        cons.add(k, trap)
    cons.add(N - 1, api.goal_constraint(xf))
    Xg = [[2.5, 2.5, 0.0], [4.0, 5.0, 0.785], [5.0, 6.25, 0.0], [7.5, 6.25, -0.261], [9, 5.0, -1.57], [7.5, 2.5, 0.0]]
    X0 = problems.natural_spline_rows(N, tf, Xg)
    p = api.Problem(model, obj, constraints=cons, x0=x0, xf=xf, N=N, tf=tf, U0=np.ones((N - 1, m)), X0=X0)
    al = api.AugmentedLagrangianSolverOptions(cost_tolerance=1e-4, cost_tolerance_intermediate=1e-2,
                                              constraint_tolerance=1e-3, penalty_scaling=50.0, penalty_initial=10.0)
    return p, api.ALTROSolverOptions(opts_al=al, R_inf=1e-3)


def printed_close(value, printed, extra_rel=0.0):
    """True if `value` rounds to the printed string (allowing one unit in the last printed digit
    plus an optional relative slack for late-iteration drift)."""
    s = printed.lower()
    ref = float(s)
    mant = s.split("e")[0]
    decimals = len(mant.split(".")[1]) if "." in mant else 0
    expo = int(s.split("e")[1]) if "e" in s else 0
    ulp = 10.0 ** (expo - decimals)
    return abs(value - ref) <= 1.01 * ulp + extra_rel * abs(ref)


def quadrotor_test_problem(to, kind):
    """test/quadrotor_tests.jl:1-60: rk4 quadrotor, Q=R=1e-2 I, Qf=1000 I, fly 50 m in y, N=101, dt=0.05,
    hover controls; `kind` in {"none", "goal", "goal+bounds"}."""
    import numpy as np
    n, m, N, dt = 13, 4, 101, 0.05
    x0 = np.zeros(n)
    x0[3] = 1.0
    xf = x0.copy()
    xf[1] = 50.0
    obj = to.LQRObjective(1e-2 * np.eye(n), 1e-2 * np.eye(m), 1000.0 * np.eye(n), xf, N)
    cons = to.Constraints(N)
    if "bounds" in kind:
        for k in range(N - 1):
            cons.add(k, to.BoundConstraint(n, m, u_min=0.0, u_max=15.0))
    if "goal" in kind:
        cons.add(N - 1, to.goal_constraint(xf))
    p = to.Problem(to.rk4(to.Dynamics.quadrotor), obj, constraints=cons, x0=x0, xf=xf, N=N, dt=dt,
                   U0=np.full((N - 1, m), 0.5 * 9.81 / 4.0))
    il = to.iLQRSolverOptions(cost_tolerance=1e-5)
    al = to.AugmentedLagrangianSolverOptions(opts_uncon=il, constraint_tolerance=1e-3, cost_tolerance=1e-5,
                                             cost_tolerance_intermediate=1e-4)
    return p, il, al


def pendulum_mintime_test(to):
    """test/minimum_time_tests.jl:1-63: pendulum rk3, Q=R=Qf=1e-3 I, N=31, bounds +-5 and goal; first a fixed-time ALTRO
    solve at dt=0.15, then the minimum-time solve (tf=:min, dt=0.075) warm-started with its controls.
    Returns (make_problem(dt, tf, U0), options)."""
    import numpy as np
    n, m, N = 2, 1, 31
    xf = np.array([np.pi, 0.0])

    def make(dt, tf=float("nan"), U0=None):
        obj = to.LQRObjective(1e-3 * np.eye(n), 1e-3 * np.eye(m), 1e-3 * np.eye(n), xf, N)
        cons = to.Constraints(N)
        bnd = to.BoundConstraint(n, m, u_min=-5.0, u_max=5.0)
        for k in range(N - 1):
            cons.add(k, bnd)
        cons.add(N - 1, to.goal_constraint(xf))
        return to.Problem(to.rk3(to.Dynamics.pendulum), obj, constraints=cons, x0=np.zeros(n), xf=xf, N=N, dt=dt, tf=tf,
                          U0=np.ones((N - 1, m)) if U0 is None else U0)
    al = to.AugmentedLagrangianSolverOptions(iterations=50, penalty_scaling=10.0)
    return make, to.ALTROSolverOptions(opts_al=al, R_minimum_time=15.0, dt_max=0.15, dt_min=1e-3)
