"""Consumes the traces `baseline/run_reference.jl` writes when somebody runs the REAL reference (Julia 1.1 +
TrajectoryOptimization.jl v0.1.1): tests/golden/julia_quadrotor_traces.json, julia_sqrt_traces.json.

They pin the parts of the CPU oracle no committed reference artefact pins (quadrotor model, regularisation restarts, square-root
pass inside a full solve).  The build image has no Julia, so the files are normally absent and these tests skip; the tolerance
is the north-star's: integer observables exact, floats 1e-8 relative.
"""
import json
import os

import numpy as np
import pytest

from trajopt_b200 import problems

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("%s not present (produced by baseline/run_reference.jl on a machine with Julia)" % name)
    return json.load(open(path))


def _num(v):
    return float(v) if not isinstance(v, str) else float(v.replace("\"", ""))


def _check_trace(tr, ref):
    res, outer, inner = ref["results"][0], ref["outer"][0], ref["inner"][0]
    assert int(res["iterations_outer"]) == tr["iterations_outer"]
    assert int(res["iterations_total"]) == tr["iterations_total"]
    assert [int(o["iterations_inner"]) for o in outer] == [int(v) for v in tr["iterations_inner"]]
    np.testing.assert_allclose([o["cost"] for o in outer], [_num(v) for v in tr["cost"]], rtol=1e-8)
    np.testing.assert_allclose([o["c_max"] for o in outer], [_num(v) for v in tr["c_max"]], rtol=1e-8, atol=1e-12)
    for oi, st in enumerate(tr["inner"]):
        ours = [r for r in inner if r["outer"] == oi]
        assert len(ours) == st["iterations"]
        np.testing.assert_allclose([r["cost"] for r in ours], [_num(v) for v in st["cost"]], rtol=1e-8)
    np.testing.assert_allclose(ref["X"][0], np.array([[_num(v) for v in row] for row in tr["X"]]), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(ref["U"][0], np.array([[_num(v) for v in row] for row in tr["U"]]), rtol=1e-8, atol=1e-10)


def test_quadrotor_traces_from_julia(oracle):
    data = _load("julia_quadrotor_traces.json")
    prob, opts = problems.quadrotor(), problems.quadrotor_bench_options()
    for tr in data["traces"]:
        b = int(tr["problem"])
        x0 = problems.batch_x0("quadrotor", 1, offset=b)
        _check_trace(tr, oracle.solve(prob, opts, x0=x0, B=1))


def test_sqrt_full_solve_from_julia(oracle):
    data = _load("julia_sqrt_traces.json")
    prob, opts = problems.pendulum(), problems.quadrotor_bench_options()
    opts.opts_al.opts_uncon.square_root = True
    for tr in data["traces"]:
        _check_trace(tr, oracle.solve(prob, opts, B=1))
