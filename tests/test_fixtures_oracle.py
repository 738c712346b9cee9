"""CPU checks of the fixtures added from the reference's problem zoo (problems/car_3obs.jl, problems/quad_obs.jl) on the oracle:
the solves end feasible and the returned trajectories keep clear of the obstacles the rows describe (what the reference's own
solve-level tests assert for its other fixtures, e.g. test/quadrotor_tests.jl:38-60)."""
import numpy as np

import trajopt_b200 as to
from cases import CASES


def test_car_3obs_altro_clears_the_obstacles(oracle):
    prob, opts, x0, _ = CASES["car_3obs_altro"](4)
    ref = oracle.solve(prob, opts, x0=x0, B=4, inner_cap=0, outer_cap=0)
    r = ref["results"]
    assert np.all(r["status"] == 0) and np.all(r["c_max"] < opts.opts_al.constraint_tolerance)
    X = ref["X"]
    assert np.abs(X[:, -1] - prob.xf).max() < 1e-3
    for cx, cy, rad in ((0.25, 0.25, 0.1), (0.5, 0.5, 0.1), (0.75, 0.75, 0.1)):
        d2 = (X[:, 1:-1, 0] - cx) ** 2 + (X[:, 1:-1, 1] - cy) ** 2
        assert d2.min() > rad ** 2 - 1e-3


def test_quad_obs_rows_and_feasible_solve(oracle):
    prob = to.problems.quad_obs()
    # knot 0: 8 control bounds; interior: 8 control bounds + 4 finite state bounds + 4 cylinders + 3 spheres; terminal: 2 x 9 state rows
    assert prob.constraints.num_constraints()[:2] == [8, 19] and prob.constraints.num_constraints()[-1] == 18
    assert np.all(prob.U == 0.0)  # problems/quad_obs.jl:86 initialises the OTHER problem's controls
    prob, opts, x0, _ = CASES["quad_obs_al"](2)
    ref = oracle.solve(prob, opts, x0=x0, B=2, inner_cap=0, outer_cap=0)
    r = ref["results"]
    assert np.all(r["status"] == 0) and np.all(r["c_max"] < opts.constraint_tolerance)
    X, U = ref["X"], ref["U"]
    assert U.min() > -1e-4 and U.max() < 50.0 + 1e-4
    for cx, cy, rad in ((0.0, 10.0, 5.0), (10.0, 30.0, 5.0), (-13.0, 25.0, 4.0), (5.0, 50.0, 6.0)):
        d2 = (X[:, 1:-1, 0] - cx) ** 2 + (X[:, 1:-1, 1] - cy) ** 2
        assert d2.min() > rad ** 2 - 1e-2
