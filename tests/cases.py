"""The parity cases: BASELINE.json's configs at sizes the CPU oracle finishes in seconds.
Each entry returns (problem, options, x0[B,n], X0 or None)."""
import numpy as np

from trajopt_b200 import api, problems
from helpers import acrobot_notebook, car_escape_notebook


def _di(B):
    p = problems.doubleintegrator()
    return p, api.ALTROSolverOptions(), np.broadcast_to(p.x0, (B, 2)).copy(), None


def _pend_ilqr(B):
    p = problems.pendulum(N=101, dt=0.045, constrained=False)
    return p, api.iLQRSolverOptions(), problems.batch_x0("pendulum", B), None


def _cart_ilqr(B):
    p = problems.cartpole(constrained=False)
    return p, api.iLQRSolverOptions(), problems.batch_x0("cartpole", B), None


def _cart_altro(B):
    p = problems.cartpole(constrained=True)
    return p, api.ALTROSolverOptions(), problems.batch_x0("cartpole", B), None


def _pend_altro(B):
    p = problems.pendulum()
    return p, api.ALTROSolverOptions(), problems.batch_x0("pendulum", B) * 0.2, None


def _pend_integrator(integ):
    """test/pendulum_tests.jl:9-27: the pendulum ALTRO solve over the explicit integration schemes"""
    def make(B):
        p = problems.pendulum()
        p.model = {"rk4": api.rk4, "midpoint": api.midpoint}[integ](api.Dynamics.pendulum)
        al = api.AugmentedLagrangianSolverOptions(iterations=50, penalty_scaling=10.0)
        o = api.ALTROSolverOptions(opts_al=al, R_minimum_time=15.0, dt_max=0.15, dt_min=1e-3)
        x0 = problems.batch_x0("pendulum", B) * 0.2
        x0[0] = 0.0
        return p, o, x0, None
    return make


def _quad(B):
    p = problems.quadrotor()
    return p, problems.quadrotor_bench_options(), problems.batch_x0("quadrotor", B), None


def _quad_regdiv(B):
    """problems 6324.. of the synthetic quadrotor batch: #6326 makes the reference's backward-pass restart
    loop diverge (Q1 double accumulation -> NaN Quu -> endless restarts); the engine reports
    TO_STATUS_REG_DIVERGED instead of hanging."""
    p = problems.quadrotor()
    return p, problems.quadrotor_bench_options(), problems.batch_x0("quadrotor", B, offset=6324), None


def _quad_ilqr(B):
    p = problems.quadrotor()
    return p, api.iLQRSolverOptions(), problems.batch_x0("quadrotor", B), None


def _acrobot_al(B):
    p, al = acrobot_notebook()
    x0 = problems.batch_x0("acrobot", B)
    x0[0] = 0.0
    return p, al, x0, None


def _dp_ilqr(B):
    p = problems.doublependulum()
    return p, api.iLQRSolverOptions(), problems.batch_x0("doublependulum", B), None


def escape_options():
    """examples/IROS_2019/car_escape.jl:12-32 with projected Newton off"""
    al = api.AugmentedLagrangianSolverOptions(cost_tolerance=1e-6, cost_tolerance_intermediate=1e-2,
                                              constraint_tolerance=1e-3, penalty_scaling=50.0, penalty_initial=10.0)
    return api.ALTROSolverOptions(opts_al=al, R_inf=0.1, resolve_feasible_problem=False)


def _escape(B):
    p = problems.car_escape()
    x0 = problems.batch_x0("car_escape", B)
    x0[0] = p.x0
    X0 = np.broadcast_to(p.X, (B,) + p.X.shape).copy()
    X0[:, 0, :] = x0
    return p, escape_options(), x0, X0


def _escape_notebook(B):
    p, o = car_escape_notebook()
    x0 = np.broadcast_to(p.x0, (B, 3)).copy()
    X0 = np.broadcast_to(p.X, (B,) + p.X.shape).copy()
    return p, o, x0, X0


def park_options():
    """benchmark/car_benchmarks.jl:18-30 with projected Newton off"""
    al = api.AugmentedLagrangianSolverOptions(iterations=30, penalty_scaling=10.0, constraint_tolerance=1e-3)
    return api.ALTROSolverOptions(opts_al=al)


def _park_inf(B):
    p = problems.parallel_park(infeasible=True)
    x0 = problems.batch_x0("parallel_park", B)
    x0[0] = 0.0
    X0 = np.stack([problems.line_trajectory(x0[b], p.xf, p.N) for b in range(B)])
    return p, park_options(), x0, X0


def _park(B):
    p = problems.parallel_park(infeasible=False)
    x0 = problems.batch_x0("parallel_park", B)
    return p, park_options(), x0, None


def _car_3obs(B):
    """problems/car_3obs.jl with the ALTRO options of examples/IROS_2019/car_3obs.jl:10-27 (projected Newton off: the AL phase
    stops at the tolerance altro_methods.jl:6-9 would hand it, 1e-3)"""
    p = problems.car_3obs()
    al = api.AugmentedLagrangianSolverOptions(cost_tolerance=1e-4, cost_tolerance_intermediate=1e-2, constraint_tolerance=1e-3,
                                              penalty_scaling=50.0, penalty_initial=10.0)
    return p, api.ALTROSolverOptions(opts_al=al), problems.batch_x0("car_3obs", B), None


def _quad_obs(B):
    """problems/quad_obs.jl with the AL options of examples/quadrotor/quad_obs.jl:85-95 (cylinders = circle rows, spheres =
    sphere rows, state + control bounds with infinite entries trimmed)"""
    p = problems.quad_obs()
    il = api.iLQRSolverOptions(iterations=300)
    al = api.AugmentedLagrangianSolverOptions(opts_uncon=il, iterations=25, cost_tolerance=1e-5, cost_tolerance_intermediate=1e-3,
                                              constraint_tolerance=1e-4, penalty_scaling=10.0, penalty_initial=0.1)
    x0 = problems.batch_x0("quadrotor", B)
    x0[0] = p.x0
    return p, al, x0, None


def _quad_maze(B):
    """problems/quadrotor_maze.jl (44 cylinders, state + control bounds, terminal box) from the hover controls, with the AL options
    of benchmark/quadrotor_benchmarks.jl:12-34 the maze benchmark shares with the line problem"""
    p = problems.quadrotor_maze()
    x0 = problems.batch_x0("quadrotor", B)
    x0[0] = p.x0
    o = problems.quadrotor_bench_options()
    o.opts_al.iterations = 5  # from hover the straight line runs into the middle wall: keep the (diverging) solve short
    o.opts_al.opts_uncon.iterations = 60
    return p, o, x0, None


def _pend_mintime(B):
    """test/minimum_time_tests.jl:17-19,38-46 (pendulum, R_minimum_time=15, dt_max=0.15)"""
    p = problems.pendulum()
    p.tf = 0.0
    al = api.AugmentedLagrangianSolverOptions(iterations=50, penalty_scaling=10.0)
    o = api.ALTROSolverOptions(opts_al=al, R_minimum_time=15.0, dt_max=0.15, dt_min=1e-3)
    return p, o, np.broadcast_to(p.x0, (B, 2)).copy() + 0.05 * problems.batch_x0("pendulum", B), None


def _inf_mintime(resolve):
    """infeasible start AND minimum time in ONE ALTRO solve (altro_methods.jl:98-124; the wish-list at test/runtests.jl:65-90 names
    it as untested): parallel park with a straight-line state guess and tf = 0.  The reference composes the two transforms so
    that the sqrt(dt) bounds of knots that own a BoundConstraint land on the first slack control (see build_variant / build_spec):
    the solve runs into the penalty ceiling instead of converging -- reproduced, and bit-compared, as it is."""
    def make(B):
        p = problems.parallel_park(infeasible=True)
        p.tf = 0.0
        x0 = problems.batch_x0("parallel_park", B)
        x0[0] = 0.0
        X0 = np.stack([problems.line_trajectory(x0[b], p.xf, p.N) for b in range(B)])
        al = api.AugmentedLagrangianSolverOptions(iterations=8, penalty_scaling=10.0, constraint_tolerance=1e-3)
        o = api.ALTROSolverOptions(opts_al=al, R_minimum_time=10.0, dt_max=0.2, dt_min=1e-3, resolve_feasible_problem=resolve)
        return p, o, x0, X0
    return make


def _sqrt_opts(**al_kw):
    il = api.iLQRSolverOptions(square_root=True)
    return api.AugmentedLagrangianSolverOptions(opts_uncon=il, **al_kw)


def _pend_sqrt_altro(B):
    """square-root backward pass (backward_pass.jl:87-169) in a full ALTRO solve: pendulum with bounds + goal"""
    p = problems.pendulum()
    return p, api.ALTROSolverOptions(opts_al=_sqrt_opts()), problems.batch_x0("pendulum", B) * 0.2, None


def _dp_sqrt_ilqr(B):
    p = problems.doublependulum()
    return p, api.iLQRSolverOptions(square_root=True), problems.batch_x0("doublependulum", B), None


def _quad_sqrt_ilqr(B):
    """the square-root pass on the largest model (n = 13, m = 4): unconstrained quadrotor iLQR"""
    p = problems.quadrotor()
    p.constraints = api.Constraints(p.N)
    return p, api.iLQRSolverOptions(square_root=True, iterations=40), problems.batch_x0("quadrotor", B), None


def _acrobot_sqrt_al(B):
    p = problems.acrobot()
    x0 = problems.batch_x0("acrobot", B)
    x0[0] = 0.0
    return p, _sqrt_opts(iterations=30, penalty_scaling=10.0), x0, None


def _sqrt_mintime(which):
    """BASELINE config 5: square-root backward pass + minimum time: +-20 torque bounds (needed by
    minimum_time.jl:6), tf = 0, R_minimum_time = 15, dt_max = 0.02, dt_min = 1e-3.  Per-problem NOT_PD statuses are
    expected; parity = same status."""
    def make(B):
        p = problems.acrobot() if which == "acrobot" else problems.doublependulum()
        n, m = p.model.n, p.model.m
        bnd = api.BoundConstraint(n, m, u_min=-20.0, u_max=20.0)
        for k in range(p.N - 1):
            p.constraints.add(k, bnd)
        p.tf = 0.0
        o = api.ALTROSolverOptions(opts_al=_sqrt_opts(), R_minimum_time=15.0, dt_max=0.02, dt_min=1e-3)
        x0 = problems.batch_x0(which if which == "acrobot" else "doublependulum", B)
        return p, o, x0, None
    return make


CASES = {
    "di_altro": _di,
    "pend_ilqr": _pend_ilqr,
    "cart_ilqr": _cart_ilqr,
    "pend_altro": _pend_altro,
    "cart_altro": _cart_altro,
    "quad_ilqr": _quad_ilqr,
    "quad_altro": _quad,
    "quad_regdiv": _quad_regdiv,
    "acrobot_al": _acrobot_al,
    "dp_ilqr": _dp_ilqr,
    "escape_altro": _escape,
    "escape_notebook": _escape_notebook,
    "park_altro": _park,
    "park_inf_altro": _park_inf,
    "car_3obs_altro": _car_3obs,
    "quad_obs_al": _quad_obs,
    "quad_maze_altro": _quad_maze,
    "pend_mintime": _pend_mintime,
    "park_inf_mintime": _inf_mintime(False),
    "park_inf_mintime_resolve": _inf_mintime(True),
    "pend_rk4_altro": _pend_integrator("rk4"),
    "pend_midpoint_altro": _pend_integrator("midpoint"),
    "pend_sqrt_altro": _pend_sqrt_altro,
    "dp_sqrt_ilqr": _dp_sqrt_ilqr,
    "acrobot_sqrt_al": _acrobot_sqrt_al,
    "quad_sqrt_ilqr": _quad_sqrt_ilqr,
    "acrobot_sqrt_mintime": _sqrt_mintime("acrobot"),
    "dp_sqrt_mintime": _sqrt_mintime("doublependulum"),
}
