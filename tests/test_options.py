"""Solver-option front-end (no GPU): the reference's defaults, and the options that exist in the reference but are not on the device
path are refused loudly instead of being ignored (ilqr_solver.jl:7-81)."""
import pytest

import trajopt_b200 as to


def test_defaults_follow_the_reference():
    o = to.iLQRSolverOptions()
    assert (o.cost_tolerance, o.gradient_norm_tolerance, o.iterations, o.dJ_counter_limit) == (1e-4, 1e-5, 300, 10)
    assert (o.iterations_linesearch, o.line_search_lower_bound, o.line_search_upper_bound) == (20, 1e-8, 10.0)
    assert (o.bp_reg_increase_factor, o.bp_reg_max, o.bp_reg_min, o.bp_reg_fp) == (1.6, 1e8, 1e-8, 10.0)
    a = to.AugmentedLagrangianSolverOptions()
    assert (a.iterations, a.penalty_initial, a.penalty_scaling, a.constraint_tolerance) == (30, 1.0, 10.0, 1e-3)
    assert (a.cost_tolerance_intermediate, a.dual_max, a.penalty_max) == (1e-3, 1e8, 1e8)
    t = to.ALTROSolverOptions()
    assert (t.R_inf, t.R_minimum_time, t.dt_max, t.dt_min, t.resolve_feasible_problem) == (1.0, 1.0, 1.0, 1e-3, True)


def test_unknown_option_is_an_error():
    with pytest.raises(TypeError, match="unknown option"):
        to.iLQRSolverOptions(not_an_option=1)


def test_reference_options_outside_the_device_path_are_refused():
    c = to.iLQRSolverOptions(bp_reg_type=":control", gradient_type=":todorov").to_c()   # the defaults, in Julia symbol spelling
    assert (c.bp_reg_type, c.gradient_type) == (0, 0)
    to.iLQRSolverOptions(bp_reg_initial=1.0, bp_sqrt_inv_type=":pseudo").to_c()     # declared but read nowhere in the reference
    # ilqr_solver.jl:47-52,76-80: the other regularisation / gradient types map onto the ABI's enums
    for kw, want in ((dict(bp_reg_type=":state"), (1, 0)), (dict(gradient_type=":feedforward"), (0, 1)),
                     (dict(gradient_type="l2"), (0, 2)), (dict(gradient_type=":ℓinf"), (0, 3))):
        c = to.iLQRSolverOptions(**kw).to_c()
        assert (c.bp_reg_type, c.gradient_type) == want
    with pytest.raises(ValueError, match="expected one of"):
        to.iLQRSolverOptions(gradient_type=":newton").to_c()
    with pytest.raises(NotImplementedError, match="not on the device path"):
        to.iLQRSolverOptions(bp_reg_type=":state", square_root=True).to_c()
    with pytest.raises(NotImplementedError, match="solve_type"):
        to.ALTROSolverOptions(projected_newton=True, opts_pn=to.ProjectedNewtonSolverOptions(solve_type=":optimal")).to_c()


def test_projected_newton_options_reach_the_abi():
    # altro_solver.jl:54-65 + direct_solvers.jl:14-30; the dead ALTRO fields are accepted and ignored
    o = to.ALTROSolverOptions(projected_newton=True, projected_newton_tolerance=1e-4, constraint_tolerance_infeasible=1e-5,
                              penalty_scaling_minimum_time_equality=2.0,
                              opts_pn=to.ProjectedNewtonSolverOptions(feasibility_tolerance=1e-8, n_steps=2)).to_c()
    assert (o.projected_newton, o.pn_n_steps, o.projected_newton_tolerance) == (1, 2, 1e-4)
    assert (o.pn_feasibility_tolerance, o.pn_active_set_tolerance) == (1e-8, 1e-3)
    d = to.ALTROSolverOptions().to_c()
    assert (d.projected_newton, d.pn_n_steps, d.pn_feasibility_tolerance) == (0, 1, 1e-6)


def test_options_nest_like_the_reference():
    il = to.iLQRSolverOptions(iterations=50, square_root=True)
    al = to.AugmentedLagrangianSolverOptions(opts_uncon=il, iterations=7)
    c = to.ALTROSolverOptions(opts_al=al, R_inf=0.5).to_c()
    assert c.opts_al.opts_uncon.iterations == 50 and c.opts_al.opts_uncon.square_root == 1
    assert c.opts_al.iterations == 7 and c.R_inf == 0.5


def test_trim_entry_follows_the_reference_logger():
    # src/logger.jl:171-194: fixed notation inside the column's range, exponent notation outside, a leading space for positives
    from trajopt_b200.api import _trim_entry
    assert _trim_entry(4458.9547984, 14) == " 4458.9547984 ".ljust(14)
    assert _trim_entry(0.125, 10) == " 0.125".ljust(10)
    assert _trim_entry(-3.5, 10) == "-3.5".ljust(10)
    assert _trim_entry(1.0e-9, 10).strip() == "1e-09" and _trim_entry(12, 6) == "12".ljust(6)
    assert _trim_entry(float("inf"), 10).strip() == "Inf"


def test_time_validation_follows_the_reference():
    # src/problem.jl:169-220 (and test/problem_tests.jl:53-55,77-80: dt == 0.3 from tf, N; disagreement throws)
    from trajopt_b200.api import _validate_time
    nan = float("nan")
    assert _validate_time(11, 3.0, nan) == (11, 3.0, 0.3)
    assert _validate_time(21, nan, 0.05) == (21, 0.05 * 20, 0.05)
    assert _validate_time(21, 0.0, 0.1) == (21, 0.0, 0.1)                  # minimum time keeps the initial dt
    for args in ((11, 3.0, 0.25), (11, nan, nan), (11, nan, -0.1), (11, 0.0, nan), (11, -1.0, 0.1)):
        with pytest.raises(ValueError):
            _validate_time(*args)
