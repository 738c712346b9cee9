"""Projected-Newton polish of ALTRO (src/solvers/direct/projected_newton.jl, solve_type = :feasible; hook at
altro/altro_methods.jl:6-14,31-39).

No committed reference artefact pins the PN iterates (the reference factorises with CHOLMOD; the oracle and the device use a
block-tridiagonal dense Cholesky of the same matrix), so the CPU tests pin PROPERTIES the reference guarantees -- the constraint
violation AND the dynamics defect of the polished trajectory are below opts_pn.feasibility_tolerance (what
examples/quadrotor/altro_times.txt-style runs report as 1e-8 feasibility), the AL phase stops at projected_newton_tolerance,
the polish moves the trajectory only slightly -- and the GPU test pins the device against the oracle at the north-star 1e-8.
"""
import numpy as np
import pytest

from trajopt_b200 import api, problems
from cases import CASES


def _pn_options(base, feas=1e-8, tol=1e-3):
    o = api.ALTROSolverOptions(opts_al=base.opts_al, R_inf=base.R_inf, resolve_feasible_problem=base.resolve_feasible_problem,
                               dynamically_feasible_projection=base.dynamically_feasible_projection,
                               projected_newton=True, projected_newton_tolerance=tol,
                               opts_pn=api.ProjectedNewtonSolverOptions(feasibility_tolerance=feas))
    return o


def _dynamics_defect(oracle, prob, X, U, x0):
    """max |f(x_k,u_k) - x_{k+1}| and |x_1 - x0| through the oracle's own discrete dynamics"""
    spec = oracle.Spec(prob)
    d = np.abs(X[0] - x0).max()
    for k in range(prob.N - 1):
        d = max(d, np.abs(spec.dynamics(X[k], U[k], jac=False)[0] - X[k + 1]).max())
    return d


@pytest.mark.parametrize("name", ["quad_altro", "cart_altro", "di_altro"])
def test_oracle_pn_reaches_feasibility(oracle, name):
    B = 2
    prob, opts, x0, X0 = CASES[name](B)
    al = oracle.solve(prob, opts, x0=x0, X0=X0, B=B)
    pn = oracle.solve(prob, _pn_options(opts), x0=x0, X0=X0, B=B)
    assert np.all(pn["results"]["status"] & 64 == 0)
    # the AL phase is the same solve as without PN when its tolerance equals projected_newton_tolerance
    if opts.opts_al.constraint_tolerance == 1e-3:
        assert np.array_equal(al["results"]["iterations_total"], pn["results"]["iterations_total"])
    assert np.all(pn["results"]["c_max"] <= 1e-8)
    for b in range(B):
        assert _dynamics_defect(oracle, prob, pn["X"][b], pn["U"][b], x0[b]) <= 1e-8
        # a polish, not another solve: the trajectory moves by about the AL phase's violation
        assert np.abs(pn["X"][b] - al["X"][b]).max() < 0.05 and np.abs(pn["U"][b] - al["U"][b]).max() < 0.05
        assert np.abs(pn["X"][b] - al["X"][b]).max() > 0.0


def test_pn_options_change_the_al_tolerance(oracle):
    # altro_methods.jl:6-14: with projected Newton the AL phase stops at projected_newton_tolerance
    prob, opts, x0, _ = CASES["quad_altro"](1)
    loose = oracle.solve(prob, _pn_options(opts, tol=1e-2), x0=x0, B=1)
    tight = oracle.solve(prob, _pn_options(opts, tol=1e-3), x0=x0, B=1)
    assert loose["results"]["iterations_outer"][0] < tight["results"]["iterations_outer"][0]
    assert loose["results"]["c_max"][0] <= 1e-8 and tight["results"]["c_max"][0] <= 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["quad_altro", "cart_altro", "di_altro", "park_inf_altro"])
def test_gpu_pn_matches_oracle(to, oracle, name):
    B = 4 if name != "quad_altro" else 3
    prob, opts, x0, X0 = CASES[name](B)
    o = _pn_options(opts)
    # infeasible start: as in the reference's examples that polish (examples/IROS_2019/car_escape.jl:12-32) the slack-free problem
    # is not re-solved afterwards -- a re-solve started from two polishes that differ in the last bits (1e-14, block-Cholesky
    # rounding) may legitimately take a different number of iterations
    o.resolve_feasible_problem = False
    ref = oracle.solve(prob, o, x0=x0, X0=X0, B=B)
    bs = to.api.BatchSolver(prob, B, 0, 0, 0)
    try:
        bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), X0)
        bs.solve(o)
        X, U, _ = bs.solution()
        res = bs.results()
    finally:
        bs.close()
    for f in ("iterations_total", "iterations_outer", "status", "steps"):
        assert np.array_equal(ref["results"][f], res[f]), (f, ref["results"][f], res[f])
    assert np.allclose(ref["results"]["J"], res["J"], rtol=1e-8, atol=0)
    ok = (res["status"] == 0)
    assert np.all(res["c_max"][ok] <= 1e-8) and ok.sum() >= B - 1
    assert np.allclose(ref["results"]["c_max"], res["c_max"], rtol=1e-6, atol=1e-12)
    assert np.allclose(ref["X"], X, rtol=1e-8, atol=1e-10) and np.allclose(ref["U"], U, rtol=1e-8, atol=1e-10)


@pytest.mark.gpu
def test_gpu_pn_refuses_minimum_time(to):
    prob, opts, x0, X0 = CASES["pend_mintime"](2)
    bs = to.api.BatchSolver(prob, 2, 0, 0, 0)
    try:
        bs.set_batch(x0, np.broadcast_to(prob.U, (2,) + prob.U.shape), X0)
        with pytest.raises(RuntimeError, match="minimum time"):
            bs.solve(_pn_options(opts))
    finally:
        bs.close()
