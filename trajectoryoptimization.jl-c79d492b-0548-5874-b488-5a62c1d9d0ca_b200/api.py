"""Host-side mirror of the TrajectoryOptimization.jl v0.1.1 user API for the iLQR/AL/ALTRO path.

Julia is not available in this image, so the host side above the C ABI is Python with the
reference's names and argument meaning (README.md:29-67):

    Model / Dynamics.*            src/model.jl:103-115, dynamics/*.jl
    rk3 / rk4 / midpoint          src/model.jl:641-643
    LQRObjective, QuadraticCost   src/objective.jl:102-114, src/cost.jl:112-169
    BoundConstraint               src/constraints.jl:140-188
    goal_constraint               src/constraints.jl:299-304
    circle / sphere obstacle sets src/utils.jl:140-156, problems/car_escape.jl:36-41
    Constraints                   src/constraint_sets.jl:157-206
    Problem, initial_controls_b   src/problem.jl:37-157      (`!` spelled `_b`)
    iLQRSolverOptions, AugmentedLagrangianSolverOptions, ALTROSolverOptions
    solve_b(prob, opts)           src/solvers.jl:91-94, altro_methods.jl:2-53
    solve(prob, opts)             src/solvers.jl:104-108 (copying variant)
    solve_b([prob...], opts)      NEW: batched solve over B same-shape problems
    format_log(solver)            src/logger.jl: the verbose InnerLoop / OuterLoop tables, from the returned histories
    ProjectedNewtonSolverOptions  src/solvers/direct/direct_solvers.jl:14-30 (solve_type = :feasible)

`solve_b` marshals the Problem into the POD descriptor of include/trajopt_b200.h and calls the CUDA
engine.  There is no CPU path here: without the built extension or without a GPU it raises.
"""
import ctypes as C
import math

import numpy as np

from . import abi


# ------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------
class Model:
    """A model from the reference's `Dynamics` zoo.  Arbitrary Julia closures cannot run on the
    device, so a Model is identified by its id (SURVEY §7 'user closures')."""

    def __init__(self, model_id, integrator=None):
        self.id = model_id
        self.n, self.m = abi.MODEL_DIMS[model_id]
        self.integrator = integrator  # None = continuous

    def discretized(self, integ):
        return Model(self.id, integ)

    def __repr__(self):
        return "Model(%s, n=%d, m=%d, integrator=%s)" % (abi.MODEL_NAMES[self.id], self.n, self.m, self.integrator)


class Dynamics:
    doubleintegrator = Model(abi.MODEL_DOUBLE_INTEGRATOR)
    pendulum = Model(abi.MODEL_PENDULUM)
    car = Model(abi.MODEL_CAR)
    cartpole = Model(abi.MODEL_CARTPOLE)
    quadrotor = Model(abi.MODEL_QUADROTOR)
    acrobot_model = Model(abi.MODEL_ACROBOT)
    doublependulum = Model(abi.MODEL_DOUBLEPENDULUM)


def rk3(model):
    return model.discretized(abi.INTEG_RK3)


def rk4(model):
    return model.discretized(abi.INTEG_RK4)


def midpoint(model):
    return model.discretized(abi.INTEG_MIDPOINT)


# ------------------------------------------------------------------------------------------
# objective
# ------------------------------------------------------------------------------------------
class QuadraticCost:
    def __init__(self, Q, R, H=None, q=None, r=None, c=0.0):
        self.Q = np.array(Q, dtype=np.float64)
        self.R = np.array(R, dtype=np.float64)
        n, m = self.Q.shape[0], self.R.shape[0]
        self.H = np.zeros((m, n)) if H is None else np.array(H, dtype=np.float64)
        self.q = np.zeros(n) if q is None else np.array(q, dtype=np.float64)
        self.r = np.zeros(m) if r is None else np.array(r, dtype=np.float64)
        self.c = float(c)


class Objective:
    """Stage cost (same at every k<N) + terminal cost.  (The reference allows per-knot costs;
    every fixture on the hot path uses a uniform LQRObjective.)"""

    def __init__(self, stage, terminal, N):
        self.stage, self.terminal, self.N = stage, terminal, N

    def __len__(self):
        return self.N


def LQRObjective(Q, R, Qf, xf, N):
    """src/objective.jl:102-114: q=-Q*xf, c=0.5*xf'Q*xf, qf=-Qf*xf, cf=0.5*xf'Qf*xf."""
    Q = np.array(Q, dtype=np.float64)
    R = np.array(R, dtype=np.float64)
    Qf = np.array(Qf, dtype=np.float64)
    xf = np.array(xf, dtype=np.float64)
    q = -_matvec(Q, xf)
    c = float((0.5 * xf) @ Q @ xf)
    qf = -_matvec(Qf, xf)
    cf = float((0.5 * xf) @ Qf @ xf)
    stage = QuadraticCost(Q, R, None, q, None, c)
    term = QuadraticCost(Qf, np.zeros((0, 0)), np.zeros((0, Q.shape[0])), qf, np.zeros(0), cf)
    return Objective(stage, term, N)


def _matvec(A, x):
    return A @ x


def _dot(a, b):
    return float(a @ b)


# ------------------------------------------------------------------------------------------
# constraints
# ------------------------------------------------------------------------------------------
class BoundConstraint:
    """src/constraints.jl:155-188; stacking [x_max; u_max; x_min; u_min], infinite bounds trimmed."""
    label = "bound"

    def __init__(self, n, m, x_min=-np.inf, x_max=np.inf, u_min=-np.inf, u_max=np.inf):
        def vecn(v, k):
            v = np.array(v, dtype=np.float64)
            return np.full(k, float(v)) if v.ndim == 0 else v.copy()
        self.n, self.m = n, m
        self.x_max, self.x_min = vecn(x_max, n), vecn(x_min, n)
        self.u_max, self.u_min = vecn(u_max, m), vecn(u_min, m)
        for hi, lo, k in ((self.x_max, self.x_min, n), (self.u_max, self.u_min, m)):
            if len(hi) != k or len(lo) != k:
                raise ValueError("limit of wrong length")
            if not np.all(hi >= lo):
                raise ValueError("u_max must be greater than u_min")

    def rows(self, terminal):
        out = []
        n = self.n
        for i in range(n):
            if np.isfinite(self.x_max[i]):
                out.append(dict(kind=abi.ROW_LINEAR, eq=0, var=i, bound=1, sign=1.0, a=self.x_max[i]))
        if not terminal:
            for i in range(self.m):
                if np.isfinite(self.u_max[i]):
                    out.append(dict(kind=abi.ROW_LINEAR, eq=0, var=n + i, bound=1, sign=1.0, a=self.u_max[i]))
        for i in range(n):
            if np.isfinite(self.x_min[i]):
                out.append(dict(kind=abi.ROW_LINEAR, eq=0, var=i, bound=1, sign=-1.0, a=self.x_min[i]))
        if not terminal:
            for i in range(self.m):
                if np.isfinite(self.u_min[i]):
                    out.append(dict(kind=abi.ROW_LINEAR, eq=0, var=n + i, bound=1, sign=-1.0, a=self.u_min[i]))
        return out


class GoalConstraint:
    """goal_constraint(xf): terminal equality x_N - xf (src/constraints.jl:299-304)."""
    label = "goal"

    def __init__(self, xf):
        self.xf = np.array(xf, dtype=np.float64)

    def rows(self, terminal):
        if not terminal:
            return []
        return [dict(kind=abi.ROW_LINEAR, eq=1, var=i, bound=0, sign=1.0, a=self.xf[i]) for i in range(len(self.xf))]


def goal_constraint(xf):
    return GoalConstraint(xf)


class CircleConstraints:
    """A stage inequality set of planar circles, c_i = circle_constraint(x, cx, cy, r)
    (src/utils.jl:140-144; problems/car_escape.jl:36-41 `trap`)."""

    def __init__(self, circles, label="obstacles"):
        self.circles = [tuple(map(float, c)) for c in circles]
        self.label = label

    def rows(self, terminal):
        if terminal:
            return []
        return [dict(kind=abi.ROW_CIRCLE, eq=0, var=0, bound=0, sign=1.0, a=c[0], b=c[1], r=c[2]) for c in self.circles]


class SphereConstraints:
    """src/utils.jl:150-156."""

    def __init__(self, spheres, label="spheres"):
        self.spheres = [tuple(map(float, s)) for s in spheres]
        self.label = label

    def rows(self, terminal):
        if terminal:
            return []
        return [dict(kind=abi.ROW_SPHERE, eq=0, var=0, bound=0, sign=1.0, a=s[0], b=s[1], c=s[2], r=s[3])
                for s in self.spheres]


class Constraints:
    """src/constraint_sets.jl:157-206: one constraint set per knot; `C[k] += con`."""

    def __init__(self, N):
        self.C = [[] for _ in range(N)]

    def __getitem__(self, k):
        return self.C[k]

    def __setitem__(self, k, v):
        self.C[k] = v

    def __len__(self):
        return len(self.C)

    def add(self, k, con):
        self.C[k] = self.C[k] + [con]

    def num_constraints(self):
        N = len(self.C)
        return [sum(len(c.rows(k == N - 1)) for c in self.C[k]) for k in range(N)]


# ------------------------------------------------------------------------------------------
# problem
# ------------------------------------------------------------------------------------------
class Problem:
    """src/problem.jl:37-124.  `tf=0` selects a minimum-time problem; `X0` (finite) selects the
    infeasible-start path of ALTRO (altro_methods.jl:102)."""

    def __init__(self, model, obj, constraints=None, x0=None, xf=None, N=None, dt=float("nan"), tf=float("nan"), U0=None, X0=None):
        if model.integrator is None:
            raise ValueError("Problem needs a discretized model: use rk3(model)")
        self.model, self.obj = model, obj
        N = len(obj) if N is None else N
        N, tf, dt = _validate_time(N, tf, dt)
        self.N, self.dt, self.tf = N, dt, tf
        self.constraints = Constraints(N) if constraints is None else constraints
        n, m = model.n, model.m
        self.x0 = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
        self.xf = np.zeros(n) if xf is None else np.array(xf, dtype=np.float64)
        self.X = np.full((N, n), np.nan) if X0 is None else np.array(X0, dtype=np.float64).reshape(N, n).copy()
        self.U = np.zeros((N - 1, m)) if U0 is None else np.array(U0, dtype=np.float64).reshape(N - 1, m).copy()

    def copy(self):
        import copy as _c
        p = _c.copy(self)
        p.x0, p.xf, p.X, p.U = self.x0.copy(), self.xf.copy(), self.X.copy(), self.U.copy()
        return p


def _validate_time(N, tf, dt):
    """_validate_time (src/problem.jl:169-220).  N is always known here (the objective carries one cost per knot), so the
    reference's N == -1 branches do not arise; the error cases are the reference's ArgumentErrors."""
    has_dt, has_tf = not math.isnan(dt), not math.isnan(tf)
    if not has_dt and not has_tf:
        raise ValueError("Must specify at least 2: N, dt, or tf")
    if has_tf and tf == 0.0:  # minimum time (tf == 0 or :min): dt is the initial time step
        if not has_dt:
            raise ValueError("minimum-time problems need an initial dt")
    elif has_tf and tf > 0:
        if not has_dt:
            dt = tf / (N - 1)
        elif dt != tf / (N - 1):
            raise ValueError("Specified time step, number of knot points, and final time do not agree (%r != %r/%d)" % (dt, tf, N - 1))
        if dt == 0:
            raise ValueError("dt must be non-zero for non-minimum time problems")
    elif not has_tf:
        if not dt > 0:
            raise ValueError("dt must be positive for a non-minimum-time problem")
        tf = dt * (N - 1)
    else:
        raise ValueError("Invalid input for tf")
    if N < 0:
        raise ValueError("%d is not a valid entry for N. Number of knot points must be a positive integer." % N)
    if dt < 0:
        raise ValueError("%r is not a valid entry for dt. Time step must be positive." % dt)
    return N, tf, dt


def initial_controls_b(prob, U0):
    """initial_controls!(prob, U0) (src/problem.jl:149-150)."""
    U0 = np.array(U0, dtype=np.float64)
    prob.U[:] = U0.reshape(-1, prob.model.m)[: prob.N - 1]


def initial_states_b(prob, X0):
    prob.X[:] = np.array(X0, dtype=np.float64).reshape(prob.N, prob.model.n)


# ------------------------------------------------------------------------------------------
# options (defaults == the reference's @with_kw defaults)
# ------------------------------------------------------------------------------------------
class _Opts:
    _defaults = {}

    def __init__(self, **kw):
        for k, v in self._defaults.items():
            setattr(self, k, v() if callable(v) else v)
        for k, v in kw.items():
            if k not in self._defaults and k not in ("verbose", "live_plotting"):
                raise TypeError("unknown option %s" % k)
            setattr(self, k, v)


class iLQRSolverOptions(_Opts):
    _defaults = dict(cost_tolerance=1e-4, gradient_norm_tolerance=1e-5, iterations=300, dJ_counter_limit=10,
                     square_root=False, iterations_linesearch=20, line_search_lower_bound=1e-8,
                     line_search_upper_bound=10.0, bp_reg_increase_factor=1.6, bp_reg_max=1e8, bp_reg_min=1e-8,
                     bp_reg_fp=10.0, max_cost_value=1e8, max_state_value=1e8, max_control_value=1e8)

    # symbol-valued options (ilqr_solver.jl:47-52,76-80), spelled as in Julia (":state") or bare ("state")
    _symbols = dict(bp_reg_type=dict(control=0, state=1),
                    gradient_type={"todorov": 0, "feedforward": 1, "ℓ2": 2, "l2": 2, "ℓinf": 3, "linf": 3})
    _ignored = ("bp_reg_initial", "bp_sqrt_inv_type")  # declared at ilqr_solver.jl:45,64 and read nowhere in the reference

    def __init__(self, **kw):
        for k in self._ignored:
            kw.pop(k, None)
        sym = {k: kw.pop(k) for k in list(kw) if k in self._symbols}
        super().__init__(**kw)
        self.bp_reg_type = sym.get("bp_reg_type", "control")
        self.gradient_type = sym.get("gradient_type", "todorov")

    def to_c(self):
        o = abi.TOiLQROptions()
        ints = ("iterations", "dJ_counter_limit", "square_root", "iterations_linesearch")
        for k in self._defaults:
            v = getattr(self, k)
            setattr(o, k, int(v) if k in ints else float(v))
        for k, table in self._symbols.items():
            got = getattr(self, k)
            got = got.lstrip(":") if isinstance(got, str) else got
            if got not in table:
                raise ValueError("iLQRSolverOptions.%s=%r: expected one of %s" % (k, got, sorted(table)))
            setattr(o, k, table[got])
        if o.bp_reg_type == 1 and o.square_root:
            raise NotImplementedError("bp_reg_type=:state with square_root=true is not on the device path (the square-root pass "
                                      "regularises with :control only)")
        return o


class AugmentedLagrangianSolverOptions(_Opts):
    _defaults = dict(opts_uncon=iLQRSolverOptions, cost_tolerance=1e-4, cost_tolerance_intermediate=1e-3,
                     gradient_norm_tolerance=1e-5, gradient_norm_tolerance_intermediate=1e-5,
                     constraint_tolerance=1e-3, iterations=30, kickout_max_penalty=False, dual_min=-1e8,
                     dual_max=1e8, penalty_max=1e8, penalty_initial=1.0, penalty_scaling=10.0)

    def to_c(self):
        o = abi.TOALOptions()
        o.opts_uncon = self.opts_uncon.to_c()
        for k in self._defaults:
            if k == "opts_uncon":
                continue
            v = getattr(self, k)
            setattr(o, k, int(v) if k in ("iterations", "kickout_max_penalty") else float(v))
        return o


class ProjectedNewtonSolverOptions(_Opts):
    """src/solvers/direct/direct_solvers.jl:14-30.  Only solve_type = :feasible is on the device path (the one ALTRO uses
    in every reference benchmark and example)."""
    _defaults = dict(n_steps=1, solve_type="feasible", active_set_tolerance=1e-3, feasibility_tolerance=1e-6)


class ALTROSolverOptions(_Opts):
    _defaults = dict(opts_al=AugmentedLagrangianSolverOptions, R_inf=1.0, dynamically_feasible_projection=True,
                     resolve_feasible_problem=True, R_minimum_time=1.0, dt_max=1.0, dt_min=1e-3,
                     projected_newton=False, opts_pn=ProjectedNewtonSolverOptions, projected_newton_tolerance=1e-3)
    # declared at altro_solver.jl:15,27-52 and read nowhere in the reference: accepted and ignored
    _ignored = ("constraint_tolerance_infeasible", "penalty_initial_infeasible", "penalty_scaling_infeasible",
                "penalty_initial_minimum_time_inequality", "penalty_initial_minimum_time_equality",
                "penalty_scaling_minimum_time_inequality", "penalty_scaling_minimum_time_equality")

    def __init__(self, **kw):
        for k in self._ignored:
            kw.pop(k, None)
        super().__init__(**kw)

    def to_c(self):
        o = abi.TOALTROOptions()
        o.projected_newton = int(bool(self.projected_newton))
        pn = self.opts_pn
        if self.projected_newton and str(pn.solve_type).lstrip(":") != "feasible":
            raise NotImplementedError("ProjectedNewtonSolverOptions.solve_type=%r is not on the device path (only :feasible)" % (pn.solve_type,))
        o.pn_n_steps = int(pn.n_steps)
        o.projected_newton_tolerance = float(self.projected_newton_tolerance)
        o.pn_feasibility_tolerance = float(pn.feasibility_tolerance)
        o.pn_active_set_tolerance = float(pn.active_set_tolerance)
        o.opts_al = self.opts_al.to_c()
        o.R_inf = self.R_inf
        o.dynamically_feasible_projection = int(self.dynamically_feasible_projection)
        o.resolve_feasible_problem = int(self.resolve_feasible_problem)
        o.R_minimum_time, o.dt_max, o.dt_min = self.R_minimum_time, self.dt_max, self.dt_min
        return o


def as_altro_options(opts):
    """Wrap iLQR / AL options into the nested ALTRO struct the C ABI takes; returns (mode, struct)."""
    if isinstance(opts, iLQRSolverOptions):
        a = ALTROSolverOptions()
        a.opts_al.opts_uncon = opts
        return 0, a.to_c()
    if isinstance(opts, AugmentedLagrangianSolverOptions):
        a = ALTROSolverOptions()
        a.opts_al = opts
        return 1, a.to_c()
    if isinstance(opts, ALTROSolverOptions):
        return 2, opts.to_c()
    raise TypeError("unknown solver options %r" % (opts,))


# ------------------------------------------------------------------------------------------
# marshalling Problem -> TOProblemDesc
# ------------------------------------------------------------------------------------------
class Marshalled:
    """Keeps the numpy buffers alive for as long as the ctypes descriptor is used."""

    def __init__(self, prob):
        n, m, N = prob.model.n, prob.model.m, prob.N
        st, tm = prob.obj.stage, prob.obj.terminal
        f = np.asfortranarray
        self.Q, self.R, self.H = f(st.Q), f(st.R), f(st.H)
        self.q, self.r = np.ascontiguousarray(st.q), np.ascontiguousarray(st.r)
        self.Qf, self.qf = f(tm.Q), np.ascontiguousarray(tm.q)
        # constraint classes: identical row lists share a class
        classes, keys, class_of_knot = [], {}, []
        for k in range(N):
            rows = []
            for con in prob.constraints[k]:
                rows += con.rows(k == N - 1)
            if not rows:
                class_of_knot.append(-1)
                continue
            key = (k == N - 1, tuple(tuple(sorted(r.items())) for r in rows))
            if key not in keys:
                keys[key] = len(classes)
                classes.append(rows)
            class_of_knot.append(keys[key])
        self.class_of_knot = np.array(class_of_knot, dtype=np.int32)
        starts, flat = [0], []
        for rows in classes:
            flat += rows
            starts.append(len(flat))
        self.class_row_start = np.array(starts, dtype=np.int32)
        self.rows = (abi.TOConstraintRow * max(1, len(flat)))()
        for i, r in enumerate(flat):
            cr = self.rows[i]
            cr.kind, cr.equality, cr.var, cr.is_bound = r["kind"], r["eq"], r["var"], r["bound"]
            cr.sign, cr.a, cr.b, cr.c, cr.r = r["sign"], r.get("a", 0.0), r.get("b", 0.0), r.get("c", 0.0), r.get("r", 0.0)
        self.num_rows_per_knot = [0 if c < 0 else len(classes[c]) for c in class_of_knot]
        d = abi.TOProblemDesc()
        d.model, d.integrator, d.n, d.m, d.N = prob.model.id, prob.model.integrator, n, m, N
        d.dt, d.tf = prob.dt, prob.tf
        dp = lambda a: a.ctypes.data_as(abi.c_double_p)
        d.Q, d.R, d.H, d.q, d.r, d.c = dp(self.Q), dp(self.R), dp(self.H), dp(self.q), dp(self.r), st.c
        d.Qf, d.qf, d.cf = dp(self.Qf), dp(self.qf), tm.c
        d.n_classes = len(classes)
        d.class_of_knot = self.class_of_knot.ctypes.data_as(abi.c_int32_p)
        d.class_row_start = self.class_row_start.ctypes.data_as(abi.c_int32_p)
        d.rows = C.cast(self.rows, C.POINTER(abi.TOConstraintRow))
        self.desc = d


    def mismatch(self, other):
        """None if `other` (a Marshalled problem) describes the same device problem -- model, horizon, time step, cost blocks and
        per-knot constraint rows -- else the name of the first field that differs.  A batch shares ONE descriptor: only x0, U0 and
        X0 are per-problem data (include/trajopt_b200.h: to_create / to_set_batch)."""
        a, b = self.desc, other.desc
        for f in ("model", "integrator", "n", "m", "N", "dt", "tf", "c", "cf", "n_classes"):
            if getattr(a, f) != getattr(b, f):
                return f
        for f in ("Q", "R", "H", "q", "r", "Qf", "qf", "class_of_knot", "class_row_start"):
            if not np.array_equal(getattr(self, f), getattr(other, f)):
                return f
        if bytes(self.rows) != bytes(other.rows):
            return "constraint rows"
        return None


class SolverStats:
    """What `solver.stats` holds in the reference (ilqr_solver.jl:146-154,
    augmented_lagrangian_solver.jl:173-181), rebuilt from the returned histories."""

    def __init__(self, result, inner, outer):
        self.status = result["status"]
        self.stats = dict(iterations=int(result["iterations_outer"]) or int(result["iterations_total"]),
                          iterations_total=int(result["iterations_total"]), cost=[], c_max=[], iterations_inner=[])
        if outer is not None and len(outer):
            self.stats["cost"] = [float(v) for v in outer["cost"]]
            self.stats["c_max"] = [float(v) for v in outer["c_max"]]
            self.stats["penalty_max"] = [float(v) for v in outer["penalty_max"]]
            self.stats["iterations_inner"] = [int(v) for v in outer["iterations_inner"]]
        elif inner is not None and len(inner):
            self.stats["cost"] = [float(v) for v in inner["cost"]]
            self.stats["dJ"] = [float(v) for v in inner["dJ"]]
            self.stats["gradient"] = [float(v) for v in inner["gradient"]]
        self.inner = inner
        self.J = float(result["J"])
        self.c_max = float(result["c_max"])


RESULT_DTYPE = np.dtype([("J", "f8"), ("c_max", "f8"), ("iterations_total", "i4"), ("iterations_outer", "i4"),
                         ("status", "i4"), ("steps", "i4")])
INNER_DTYPE = np.dtype([("cost", "f8"), ("dJ", "f8"), ("gradient", "f8"), ("expected", "f8"), ("z", "f8"),
                        ("alpha", "f8"), ("rho", "f8"), ("outer", "i4"), ("iter", "i4")])
OUTER_DTYPE = np.dtype([("cost", "f8"), ("c_max", "f8"), ("penalty_max", "f8"), ("iterations_inner", "i4"), ("pad", "i4")])


class BatchSolver:
    """A device workspace for B problems of one shape (to_create .. to_destroy)."""

    def __init__(self, prob, B, device=0, inner_trace=0, outer_trace=0):
        self.lib = abi.load_library()
        self.marsh = Marshalled(prob)
        self.prob, self.B = prob, int(B)
        self.n, self.m, self.N = prob.model.n, prob.model.m, prob.N
        self.h = C.c_void_p()
        rc = self.lib.to_create(C.byref(self.marsh.desc), self.B, device, C.byref(self.h))
        if rc != 0:
            raise RuntimeError("to_create failed (%d): %s" % (rc, (self.lib.to_last_error(None) or b"").decode()))
        self.inner_cap, self.outer_cap = int(inner_trace), int(outer_trace)
        self._check(self.lib.to_set_trace(self.h, self.inner_cap, self.outer_cap))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("trajopt_b200 error %d: %s" % (rc, (self.lib.to_last_error(self.h) or b"").decode()))

    def close(self):
        if self.h:
            self.lib.to_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_batch(self, x0, U0, X0=None):
        B, n, m, N = self.B, self.n, self.m, self.N
        self._x0 = np.ascontiguousarray(np.broadcast_to(np.asarray(x0, dtype=np.float64).reshape(-1, n), (B, n)))
        self._U0 = np.ascontiguousarray(np.broadcast_to(np.asarray(U0, dtype=np.float64).reshape(-1, N - 1, m), (B, N - 1, m)))
        self._X0 = None if X0 is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(X0, dtype=np.float64).reshape(-1, N, n), (B, N, n)))
        self._check(self.lib.to_set_batch(self.h, self._x0.ctypes.data, self._U0.ctypes.data,
                                          None if self._X0 is None else self._X0.ctypes.data))

    def warm_start_shift(self, x0=None, shift=1):
        """MPC re-solve on the resident workspaces: next initial controls = the last solution shifted by `shift` knots (last
        control repeated); next initial states = `x0` (B x n, e.g. the measured states) or the plan's own X[shift]."""
        if x0 is not None:
            self._x0 = np.ascontiguousarray(np.broadcast_to(np.asarray(x0, dtype=np.float64).reshape(-1, self.n), (self.B, self.n)))
        self._check(self.lib.to_warm_start_shift(self.h, None if x0 is None else self._x0.ctypes.data, int(shift)))

    def solve(self, opts):
        mode, o = as_altro_options(opts)
        if mode == 0:
            self._check(self.lib.to_solve_ilqr(self.h, C.byref(o.opts_al.opts_uncon)))
        elif mode == 1:
            self._check(self.lib.to_solve_al(self.h, C.byref(o.opts_al)))
        else:
            self._check(self.lib.to_solve_altro(self.h, C.byref(o)))

    def kernel_ms(self):
        ms = C.c_float()
        self._check(self.lib.to_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def launches(self):
        c = C.c_int32()
        self._check(self.lib.to_last_launch_count(self.h, C.byref(c)))
        return c.value

    def solution(self):
        B, n, m, N = self.B, self.n, self.m, self.N
        X, U, dts = np.empty((B, N, n)), np.empty((B, N - 1, m)), np.empty((B, N - 1))
        self._check(self.lib.to_get_solution(self.h, X.ctypes.data, U.ctypes.data, dts.ctypes.data))
        return X, U, dts

    def results(self):
        res = np.empty(self.B, dtype=RESULT_DTYPE)
        self._check(self.lib.to_get_results(self.h, res.ctypes.data))
        return res

    def trace(self):
        B = self.B
        inner = np.zeros((B, max(1, self.inner_cap)), dtype=INNER_DTYPE)
        outer = np.zeros((B, max(1, self.outer_cap)), dtype=OUTER_DTYPE)
        ni, no = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        self._check(self.lib.to_get_trace(self.h, inner.ctypes.data, ni.ctypes.data, outer.ctypes.data, no.ctypes.data))
        return [inner[b, : ni[b]] for b in range(B)], [outer[b, : no[b]] for b in range(B)]

    def duals(self):
        P = C.c_int32()
        self._check(self.lib.to_num_constraint_rows(self.h, C.byref(P)))
        P = P.value
        lam, mu, act = np.zeros((self.B, P)), np.zeros((self.B, P)), np.zeros((self.B, P), dtype=np.uint8)
        if P == 0:  # unconstrained solve, or no trace enabled: nothing was recorded (to_get_duals would say so)
            return lam, mu, act
        self._check(self.lib.to_get_duals(self.h, lam.ctypes.data, mu.ctypes.data, act.ctypes.data))
        return lam, mu, act


def _has_x0_traj(prob):
    return bool(np.all(np.isfinite(prob.X))) if not np.all(np.isnan(prob.X[0])) else False


def solve_b(prob, opts, device=0, trace=True):
    """solve!(prob, opts): mutates prob.X / prob.U in place and returns a solver-like object with
    `.stats`.  `prob` may be a list of same-shape Problems (batched solve!, one result each)."""
    probs = prob if isinstance(prob, (list, tuple)) else [prob]
    p0 = probs[0]
    B = len(probs)
    oc = ic = 0
    if trace:
        al = opts.opts_al if isinstance(opts, ALTROSolverOptions) else opts
        il = al.opts_uncon if isinstance(al, AugmentedLagrangianSolverOptions) else al
        oc = (al.iterations + 2) * 2 if isinstance(al, AugmentedLagrangianSolverOptions) else 0
        ic = min(4096, il.iterations * max(1, oc // 2 if oc else 1))
    bs = BatchSolver(p0, B, device, ic, oc)
    try:
        # the batch shares ONE problem description (cost, constraints, model, horizon); only x0 / U0 / X0 are per-problem data.
        # A list whose problems differ in anything else must not be solved as copies of problem 0.
        use_X0 = not np.all(np.isnan(p0.X[0]))
        for b, p in enumerate(probs[1:], 1):
            diff = bs.marsh.mismatch(Marshalled(p))
            if diff is None and (not np.all(np.isnan(p.X[0]))) != use_X0:
                diff = "X (infeasible start given for some problems only)"
            if diff is not None:
                raise ValueError("batched solve!: problem %d differs from problem 0 in `%s`; a batch shares one problem description "
                                 "(only x0, U and X may differ) -- solve such problems in separate batches" % (b, diff))
        x0 = np.stack([p.x0 for p in probs])
        U0 = np.stack([p.U for p in probs])
        X0 = np.stack([p.X for p in probs]) if use_X0 else None
        bs.set_batch(x0, U0, X0)
        bs.solve(opts)
        X, U, dts = bs.solution()
        res = bs.results()
        inner, outer = bs.trace() if trace else ([None] * B, [None] * B)
        out = []
        for b, p in enumerate(probs):
            p.X[:], p.U[:] = X[b], U[b]
            p.dts = dts[b]
            out.append(SolverStats(res[b], inner[b], outer[b]))
        return out if isinstance(prob, (list, tuple)) else out[0]
    finally:
        bs.close()


def solve(prob, opts, device=0, trace=True):
    """solve(prob, opts) (src/solvers.jl:104-108): the copying variant -- the caller's problem(s) stay untouched, the solved
    copies are returned with the solver-like object(s): `(prob_solved, solver)`, or two lists for a batch."""
    if isinstance(prob, (list, tuple)):
        copies = [p.copy() for p in prob]
        return copies, solve_b(copies, opts, device, trace)
    cp = prob.copy()
    return cp, solve_b(cp, opts, device, trace)


# ------------------------------------------------------------------------------------------
# SolverLogger tables (src/logger.jl) rebuilt from the returned histories
# ------------------------------------------------------------------------------------------
def _trim_entry(v, width):
    """trim_entry (src/logger.jl:171-214): a value within `width` characters, fixed or exponent notation as the reference picks"""
    if isinstance(v, (int, np.integer)):
        return str(int(v)).ljust(width)
    v = float(v)
    if not math.isfinite(v):
        return ("NaN" if v != v else ("Inf" if v > 0 else "-Inf")).ljust(width)
    base = math.log10(abs(v)) if v != 0.0 else -math.inf
    if -math.ceil(width / 2) + 1 < base < math.floor(width / 2):
        prec = width - math.ceil(base) - 3 if base > 0 else width - 4
        if prec <= 0:
            width = width - prec + 1
            prec = 1
        txt = ("%.*f" % (int(prec), v)).rstrip("0").rstrip(".")
    else:
        w = 10 if width <= 8 else width
        mant, ex = ("%.*e" % (max(0, w - 8), v)).split("e")
        if "." in mant:
            mant = mant.rstrip("0").rstrip(".")
        txt = "%se%s%02d" % (mant, ex[0], abs(int(ex)))
    if v >= 0:
        txt = " " + txt  # positivespace=true
    return txt.ljust(width)


def _table(cols, widths, rows, indent):
    head = " " * indent + "".join(c.ljust(w) for c, w in zip(cols, widths)) + "\n" + "_" * indent + "-" * sum(widths)
    return "\n".join([head] + [" " * indent + "".join(_trim_entry(v, w) for v, w in zip(r, widths)) for r in rows])


def format_log(solver):
    """The verbose tables of the reference's SolverLogger (src/logger.jl:10-38,155-168): one InnerLoop table per inner solve
    (iter cost expected z α ρ dJ grad, the columns record_iteration! and forwardpass! log, ilqr_methods.jl:84-88,
    forward_pass.jl:75-78) followed by its OuterLoop row (iter total c_max, augmented_lagrangian_methods.jl:93-96).
    Needs the histories (`trace=True`)."""
    inner, stats = solver.inner, solver.stats
    out = []
    icols, iw = ["iter", "cost", "expected", "z", "α", "ρ", "dJ", "grad"], [5, 14, 12, 10, 10, 10, 10, 10]
    ocols, ow = ["iter", "total", "c_max"], [6, 7, 12]
    n_outer = len(stats.get("c_max", [])) if stats.get("iterations_inner") else 0
    total = 0
    for o in range(max(1, n_outer)):
        rows = []
        if inner is not None and len(inner):
            sel = inner[inner["outer"] == max(0, o - 1)] if n_outer else inner
            if n_outer and o == 0:
                sel = sel[:0]  # the first outer record is taken before any inner solve
            rows = [(int(r["iter"]), r["cost"], r["expected"], r["z"], r["alpha"], r["rho"], r["dJ"], r["gradient"]) for r in sel]
        if rows:
            out.append(_table(icols, iw, rows, 4))
        if n_outer:
            total += int(stats["iterations_inner"][o])
            out.append(_table(ocols, ow, [(o + 1, total, stats["c_max"][o])], 0))
    return "\n".join(out)


def max_violation(prob):
    """max_violation(prob) (src/problem.jl:242-267), evaluated on the host for reporting."""
    N = prob.N
    cmax = 0.0
    for k in range(N):
        for con in prob.constraints[k]:
            for r in con.rows(k == N - 1):
                x = prob.X[k]
                z = np.concatenate([x, prob.U[k]]) if k < N - 1 else x
                if r["kind"] == abi.ROW_LINEAR:
                    c = r["sign"] * (z[r["var"]] - r["a"])
                elif r["kind"] == abi.ROW_CIRCLE:
                    c = -((x[0] - r["a"]) ** 2 + (x[1] - r["b"]) ** 2 - r["r"] ** 2)
                else:
                    c = -((x[0] - r["a"]) ** 2 + (x[1] - r["b"]) ** 2 + (x[2] - r["c"]) ** 2 - r["r"] ** 2)
                cmax = max(cmax, abs(c) if r["eq"] else max(0.0, c))
    return cmax


def total_time(prob):
    """minimum_time.jl:74-82"""
    if prob.tf == 0.0 and getattr(prob, "dts", None) is not None:
        return float(np.sum(prob.dts))
    return prob.dt * (prob.N - 1)
