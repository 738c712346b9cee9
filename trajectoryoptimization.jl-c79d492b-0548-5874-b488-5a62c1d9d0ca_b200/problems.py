"""The reference's problem zoo (`Problems` submodule, problems/*.jl) for the hot-path configs,
plus the seeded synthetic batch generators of SURVEY.md §8(d).  Data only — no solver code."""
import numpy as np

from . import api


def _eye(n, s=1.0):
    return s * np.eye(n)


def doubleintegrator(U0=None):
    """problems/doubleintegrator.jl == README.md:29-67 (block move)."""
    model = api.rk3(api.Dynamics.doubleintegrator)
    n, m, N, dt = 2, 1, 21, 0.1
    xf = np.array([1.0, 0.0])
    obj = api.LQRObjective(_eye(n), _eye(m, 0.1), _eye(n), xf, N)
    cons = api.Constraints(N)
    bnd = api.BoundConstraint(n, m, u_max=1.5, u_min=-1.5)
    for k in range(N - 1):
        cons.add(k, bnd)
    cons.add(N - 1, api.goal_constraint(xf))
    prob = api.Problem(model, obj, constraints=cons, x0=np.zeros(n), xf=xf, N=N, dt=dt)
    if U0 is None:
        U0 = 0.01 * splitmix_uniform(1, (N - 1) * m).reshape(N - 1, m)
    api.initial_controls_b(prob, U0)
    return prob


def pendulum(N=31, dt=0.15, constrained=True):
    """problems/pendulum.jl"""
    model = api.rk3(api.Dynamics.pendulum)
    n, m = 2, 1
    xf = np.array([np.pi, 0.0])
    Q = _eye(n, 1e-3)
    obj = api.LQRObjective(Q, _eye(m, 1e-3), Q, xf, N)
    cons = api.Constraints(N)
    if constrained:
        bnd = api.BoundConstraint(n, m, u_min=-3.0, u_max=3.0)
        for k in range(N - 1):
            cons.add(k, bnd)
        cons.add(N - 1, api.goal_constraint(xf))
    return api.Problem(model, obj, constraints=cons, x0=np.zeros(n), xf=xf, N=N, dt=dt, U0=np.ones((N - 1, m)))


def cartpole(constrained=True):
    """problems/cartpole.jl"""
    model = api.rk3(api.Dynamics.cartpole)
    n, m, N = 4, 1, 101
    dt = 5.0 / (N - 1)
    xf = np.array([0.0, np.pi, 0.0, 0.0])
    obj = api.LQRObjective(_eye(n, 1e-2), _eye(m, 1e-1), _eye(n, 100.0), xf, N)
    cons = api.Constraints(N)
    if constrained:
        bnd = api.BoundConstraint(n, m, u_min=-3.0, u_max=3.0)
        for k in range(N - 1):
            cons.add(k, bnd)
        cons.add(N - 1, api.goal_constraint(xf))
    return api.Problem(model, obj, constraints=cons, x0=np.zeros(n), xf=xf, N=N, dt=dt, U0=0.01 * np.ones((N - 1, m)))


def quadrotor():
    """problems/quadrotor.jl: u >= 0 at k<N, terminal box on x[1:3], x[8:13]."""
    model = api.rk3(api.Dynamics.quadrotor)
    n, m, N = 13, 4, 101
    dt = 5.0 / (N - 1)
    x0 = np.zeros(n)
    x0[0:3] = [0.0, 0.0, 10.0]
    x0[3] = 1.0
    xf = np.zeros(n)
    xf[0:3] = [0.0, 60.0, 10.0]
    xf[3] = 1.0
    Q = _eye(n, 1e-3)
    Q[3:7, 3:7] = _eye(4, 1e-2)
    obj = api.LQRObjective(Q, _eye(m, 1e-4), _eye(n, 1000.0), xf, N)
    cons = api.Constraints(N)
    bnd3 = api.BoundConstraint(n, m, u_min=0.0)
    xU, xL = xf.copy(), xf.copy()
    xU[3:7], xL[3:7] = np.inf, -np.inf
    xU[7:10], xL[7:10] = 0.0, 0.0
    bnd_xf = api.BoundConstraint(n, m, x_min=xL, x_max=xU)
    for k in range(N - 1):
        cons.add(k, bnd3)
    cons.add(N - 1, bnd_xf)
    U0 = np.full((N - 1, m), 0.5 * 9.81 / 4.0)
    return api.Problem(model, obj, constraints=cons, x0=x0, xf=xf, N=N, dt=dt, U0=U0)


def escape_circles():
    """problems/car_escape.jl:18-40"""
    r = 0.5
    s1, s2, s3 = 30, 50, 15
    c = []
    for i in np.linspace(0, 5, s1):
        c.append((0.0, i, r))
    for i in np.linspace(0, 5, s1):
        c.append((5.0, i, r))
    for i in np.linspace(0, 5, s1):
        c.append((10.0, i, r))
    for i in np.linspace(0, 10, s2):
        c.append((i, 0.0, r))
    for i in np.linspace(0, 3, s3):
        c.append((i, 5.0, r))
    for i in np.linspace(5, 8, s3):
        c.append((i, 5.0, r))
    return c


def natural_spline_rows(N, tf, Xg):
    """interp_rows (src/utils.jl:5-15): Interpolations.CubicSplineInterpolation == natural cubic
    spline through the guess columns."""
    from scipy.interpolate import CubicSpline
    Xg = np.asarray(Xg, dtype=np.float64)
    t1 = np.linspace(0, tf, Xg.shape[0])
    t2 = np.linspace(0, tf, N)
    return np.stack([CubicSpline(t1, Xg[:, i], bc_type="natural")(t2) for i in range(Xg.shape[1])], axis=1)


def car_escape():
    """problems/car_escape.jl (infeasible start through the X0 guess)."""
    model = api.rk3(api.Dynamics.car)
    n, m, N, tf = 3, 2, 101, 3.0
    x0 = np.array([2.5, 2.5, 0.0])
    xf = np.array([7.5, 2.5, 0.0])
    obj = api.LQRObjective(_eye(n, 1e-3), _eye(m, 1e-2), _eye(n, 100.0), xf, N)
    trap = api.CircleConstraints(escape_circles(), "trap")
    bnd = api.BoundConstraint(n, m, u_min=-5.0, u_max=5.0)
    cons = api.Constraints(N)
    cons.add(0, bnd)
    for k in range(1, N - 1):
        cons.add(k, trap)
        cons.add(k, bnd)
    cons.add(N - 1, api.goal_constraint(xf))
    Xg = [[2.5, 2.5, 0.0], [4.0, 5.0, 0.785], [5.0, 6.25, 0.0], [7.5, 6.25, -0.261], [9, 5.0, -1.57], [7.5, 2.5, 0.0]]
    X0 = natural_spline_rows(N, tf, Xg)
    return api.Problem(model, obj, constraints=cons, x0=x0, xf=xf, N=N, tf=tf, U0=np.ones((N - 1, m)), X0=X0)


def line_trajectory(x0, xf, N):
    """altro/infeasible.jl:82-89 (slope = (xf-x0)/N, t = range(0,N,length=N))."""
    x0, xf = np.asarray(x0, dtype=np.float64), np.asarray(xf, dtype=np.float64)
    t = np.linspace(0, N, N)
    slope = (xf - x0) / N
    X = np.stack([slope * t[k] for k in range(N)])
    X[0], X[-1] = x0, xf
    return X


def parallel_park(infeasible=False):
    """problems/parallel_park.jl"""
    model = api.rk3(api.Dynamics.car)
    n, m, N, dt = 3, 2, 51, 0.06
    x0, xf = np.zeros(n), np.array([0.0, 1.0, 0.0])
    obj = api.LQRObjective(_eye(n, 1e-2), _eye(m, 1e-2), _eye(n, 100.0), xf, N)
    bnd1 = api.BoundConstraint(n, m, u_min=-2.0, u_max=2.0)
    bnd2 = api.BoundConstraint(n, m, x_min=[-0.25, -0.001, -np.inf], x_max=[0.25, 1.001, np.inf], u_min=-2.0, u_max=2.0)
    cons = api.Constraints(N)
    cons.add(0, bnd1)
    for k in range(1, N - 1):
        cons.add(k, bnd2)
    cons.add(N - 1, api.goal_constraint(xf))
    X0 = line_trajectory(x0, xf, N) if infeasible else None
    return api.Problem(model, obj, constraints=cons, x0=x0, xf=xf, N=N, dt=dt, U0=np.ones((N - 1, m)), X0=X0)


def car_3obs():
    """problems/car_3obs.jl: car past three circular obstacles (circle_constraint rows at 1<k<N), goal at N."""
    model = api.rk3(api.Dynamics.car)
    n, m, N, dt = 3, 2, 101, 0.05
    x0, xf = np.zeros(n), np.array([1.0, 1.0, 0.0])
    obj = api.LQRObjective(_eye(n, 1.0), _eye(m, 1e-1), _eye(n, 100.0), xf, N)
    obs = api.CircleConstraints([(0.25, 0.25, 0.1), (0.5, 0.5, 0.1), (0.75, 0.75, 0.1)], "obs")
    cons = api.Constraints(N)
    for k in range(1, N - 1):
        cons.add(k, obs)
    cons.add(N - 1, api.goal_constraint(xf))
    return api.Problem(model, obj, constraints=cons, x0=x0, xf=xf, N=N, dt=dt, U0=np.full((N - 1, m), 0.01))


def quad_obs():
    """problems/quad_obs.jl: quadrotor among 4 cylinders and 3 spheres with state + control bounds.  Two things the file does
    are kept as they are: `sphere_constraint(x, s[1], s[2], s[3], s[3] + r_quad)` takes the sphere's z as its radius (the tuples'
    fourth entry is unused), and `initial_controls!(quadrotor, U_hover)` initialises the OTHER problem, so quad_obs starts from
    U = 0 (problems/quad_obs.jl:69-70,86)."""
    model = api.rk3(api.Dynamics.quadrotor)
    n, m, N = 13, 4, 101
    dt = 5.0 / (N - 1)
    x0 = np.zeros(n)
    x0[0:3] = [0.0, 0.0, 10.0]
    x0[3] = 1.0
    xf = np.zeros(n)
    xf[0:3] = [0.0, 60.0, 10.0]
    xf[3] = 1.0
    obj = api.LQRObjective(_eye(n, 1e-3), _eye(m, 1e-2), _eye(n, 1.0), xf, N)
    x_max, x_min = np.full(n, np.inf), np.full(n, -np.inf)
    x_max[0:3] = [25.0, np.inf, 20.0]
    x_min[0:3] = [-25.0, -np.inf, 0.0]
    bnd_u = api.BoundConstraint(n, m, u_min=0.0, u_max=50.0)
    bnd = api.BoundConstraint(n, m, u_min=0.0, u_max=50.0, x_min=x_min, x_max=x_max)
    xU, xL = xf.copy(), xf.copy()
    xU[3:7], xL[3:7] = np.inf, -np.inf
    xU[7:10], xL[7:10] = 0.0, 0.0
    bnd_xf = api.BoundConstraint(n, m, x_min=xL, x_max=xU)
    r_quad = 2.0
    cyl = api.CircleConstraints([(0.0, 10.0, 3.0 + r_quad), (10.0, 30.0, 3.0 + r_quad), (-13.0, 25.0, 2.0 + r_quad),
                                 (5.0, 50.0, 4.0 + r_quad)], "cylinders")
    sph = api.SphereConstraints([(0.0, 40.0, 5.0, 5.0 + r_quad), (-5.0, 15.0, 3.0, 3.0 + r_quad), (10.0, 20.0, 7.0, 7.0 + r_quad)],
                                "spheres")
    cons = api.Constraints(N)
    cons.add(0, bnd_u)
    for k in range(1, N - 1):
        cons.add(k, bnd)
        cons.add(k, cyl)
        cons.add(k, sph)
    cons.add(N - 1, bnd_xf)
    return api.Problem(model, obj, constraints=cons, x0=x0, xf=xf, N=N, dt=dt)


def maze_cylinders():
    """problems/quadrotor_maze.jl:24-63: the 44 cylinders (x, y, radius) of the maze, in the file's order."""
    r = 2.0
    cyl = []
    for i in np.linspace(-25.0, -10.0, 5):
        cyl.append((float(i), 10.0, r))
    for i in np.linspace(10.0, 25.0, 5):
        cyl.append((float(i), 10.0, r))
    for i in np.linspace(-5.0, 5.0, 4):
        cyl.append((float(i), 30.0, r))
    for i in np.linspace(-25.0, -10.0, 5):
        cyl.append((float(i), 50.0, r))
    for i in np.linspace(10.0, 25.0, 5):
        cyl.append((float(i), 50.0, r))
    for i in np.linspace(10.0 + 2 * r, 50.0 - 2 * r, 10):
        cyl.append((-25.0, float(i), r))
    for i in np.linspace(10.0 + 2 * r, 50.0 - 2 * r, 10):
        cyl.append((25.0, float(i), r))
    return cyl


def quadrotor_maze(with_guess=False):
    """problems/quadrotor_maze.jl: the quadrotor of problems/quadrotor.jl through a maze of 44 cylinders (circle rows with the
    quadrotor radius added, :66-70), control bounds at k = 1, state + control bounds and the maze at 1 < k < N, terminal box.
    `with_guess=True` also sets the state guess X0 (natural cubic spline through the 7 way points, :110-118), which makes ALTRO
    add 13 slack controls (m = 17): wider than the 16-lane group of the backward pass, so the engine refuses that variant
    loudly ("no sm_100a kernel instantiated"); without the guess it is an AL / ALTRO problem like quad_obs."""
    model = api.rk3(api.Dynamics.quadrotor)
    n, m, N, tf = 13, 4, 101, 5.0
    dt = tf / (N - 1)
    x0 = np.zeros(n)
    x0[0:3] = [0.0, 0.0, 10.0]
    x0[3] = 1.0
    xf = np.zeros(n)
    xf[0:3] = [0.0, 60.0, 10.0]
    xf[3] = 1.0
    Q = _eye(n, 1e-3)
    Q[3:7, 3:7] = _eye(4, 1e-2)
    obj = api.LQRObjective(Q, _eye(m, 1e-4), _eye(n, 1000.0), xf, N)
    x_max, x_min = np.full(n, np.inf), np.full(n, -np.inf)
    x_max[0:3] = [25.0, np.inf, 20.0]
    x_min[0:3] = [-25.0, -np.inf, 0.0]
    bnd1 = api.BoundConstraint(n, m, u_min=0.0, u_max=50.0)
    bnd2 = api.BoundConstraint(n, m, u_min=0.0, u_max=50.0, x_min=x_min, x_max=x_max)
    xU, xL = xf.copy(), xf.copy()
    xU[3:7], xL[3:7] = np.inf, -np.inf
    bnd_xf = api.BoundConstraint(n, m, x_min=xL, x_max=xU)
    r_quad = 2.0
    maze = api.CircleConstraints([(cx, cy, cr + r_quad) for cx, cy, cr in maze_cylinders()], "maze")
    cons = api.Constraints(N)
    cons.add(0, bnd1)
    for k in range(1, N - 1):
        cons.add(k, bnd2)
        cons.add(k, maze)
    cons.add(N - 1, bnd_xf)
    X0 = None
    if with_guess:
        Xg = np.zeros((7, n))
        Xg[:, 3] = 1.0
        Xg[0], Xg[6] = x0, xf
        Xg[1:6, 0] = [0.0, -12.5, -20.0, -12.5, 0.0]
        Xg[1:6, 1] = [15.0, 20.0, 30.0, 40.0, 45.0]
        Xg[1:6, 2] = 10.0
        X0 = natural_spline_rows(N, tf, [list(r) for r in Xg])
    return api.Problem(model, obj, constraints=cons, x0=x0, xf=xf, N=N, dt=dt, U0=np.full((N - 1, m), 0.5 * 9.81 / 4.0), X0=X0)


def acrobot(N=151, dt=0.01, Qs=1e-2, Rs=1e-2, Qfs=100.0):
    """problems/acrobot.jl (goal constraint at N)."""
    model = api.rk3(api.Dynamics.acrobot_model)
    n, m = 4, 1
    xf = np.array([np.pi, 0.0, 0.0, 0.0])
    obj = api.LQRObjective(_eye(n, Qs), _eye(m, Rs), _eye(n, Qfs), xf, N)
    cons = api.Constraints(N)
    cons.add(N - 1, api.goal_constraint(xf))
    return api.Problem(model, obj, constraints=cons, x0=np.zeros(n), xf=xf, N=N, dt=dt, U0=np.ones((N - 1, m)))


def doublependulum():
    """problems/doublependulum.jl (unconstrained)."""
    model = api.rk3(api.Dynamics.doublependulum)
    n, m, N, dt = 4, 2, 101, 0.01
    xf = np.array([np.pi, 0.0, 0.0, 0.0])
    obj = api.LQRObjective(_eye(n, 1e-2), _eye(m, 1e-2), _eye(n, 100.0), xf, N)
    return api.Problem(model, obj, x0=np.zeros(n), xf=xf, N=N, dt=dt, U0=np.ones((N - 1, m)))


# ------------------------------------------------------------------------------------------
# synthetic batches: splitmix64 -> uniform [0,1)
# ------------------------------------------------------------------------------------------
def splitmix_uniform(seed, count):
    """count uniforms in [0,1) from splitmix64 seeded with `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _u(rnd, a, b):
    return a + (b - a) * rnd


def batch_x0(config, B, offset=0):
    """Per-problem initial states for the named BASELINE.json configs; problem b uses the
    stream seeded 1000*config_id + (offset+b)."""
    if config == "quadrotor":
        out = np.zeros((B, 13))
        for b in range(B):
            r = splitmix_uniform(3000 + offset + b, 13)
            out[b, 0:3] = np.array([0.0, 0.0, 10.0]) + _u(r[0:3], -2, 2)
            q = np.array([1.0, 0, 0, 0]) + 0.1 * _u(r[3:7], -1, 1)
            out[b, 3:7] = q / np.linalg.norm(q)
            out[b, 7:13] = _u(r[7:13], -0.5, 0.5)
        return out
    if config == "pendulum":
        out = np.zeros((B, 2))
        for b in range(B):
            r = splitmix_uniform(2000 + offset + b, 2)
            out[b] = [_u(r[0], -np.pi / 2, np.pi / 2), _u(r[1], -1, 1)]
        return out
    if config == "cartpole":
        out = np.zeros((B, 4))
        for b in range(B):
            out[b] = _u(splitmix_uniform(2500 + offset + b, 4), -0.5, 0.5)
        return out
    if config == "car_escape":
        out = np.zeros((B, 3))
        for b in range(B):
            out[b] = np.array([2.5, 2.5, 0.0]) + _u(splitmix_uniform(4000 + offset + b, 3), -0.3, 0.3)
        return out
    if config == "parallel_park":
        out = np.zeros((B, 3))
        for b in range(B):
            r = splitmix_uniform(4500 + offset + b, 3)
            out[b] = [_u(r[0], -0.1, 0.1), _u(r[1], -0.05, 0.05), _u(r[2], -0.2, 0.2)]
        return out
    if config == "car_3obs":
        out = np.zeros((B, 3))
        for b in range(B):
            r = splitmix_uniform(6000 + offset + b, 3)
            out[b] = [_u(r[0], -0.05, 0.05), _u(r[1], -0.05, 0.05), _u(r[2], -0.1, 0.1)]
        return out
    if config in ("acrobot", "doublependulum"):
        out = np.zeros((B, 4))
        for b in range(B):
            r = splitmix_uniform(5000 + offset + b, 2)
            out[b, 0:2] = _u(r, -0.2, 0.2)
        return out
    raise KeyError(config)


def quadrotor_bench_options():
    """AL phase of benchmark/quadrotor_benchmarks.jl:12-34 with projected Newton off
    (constraint_tolerance = projected_newton_tolerance 1e-3, altro_methods.jl:6-9)."""
    il = api.iLQRSolverOptions(iterations=300)
    al = api.AugmentedLagrangianSolverOptions(opts_uncon=il, iterations=40, cost_tolerance=1e-5,
                                              cost_tolerance_intermediate=1e-4, constraint_tolerance=1e-3,
                                              penalty_scaling=10.0, penalty_initial=1.0)
    return api.ALTROSolverOptions(opts_al=al, R_inf=1e-8, resolve_feasible_problem=False, projected_newton=False)
