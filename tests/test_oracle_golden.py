"""Pin the CPU oracle against REAL reference output: the verbose iteration tables the reference's
authors left in examples/acrobot/Acrobot.ipynb (cell 19) and examples/car/Car Escape.ipynb
(cell 23), extracted verbatim by tests/golden/extract_notebook_traces.py.

Those notebooks numbered inner iterations from 1 at the first step!; the snapshot counts the
initial record too (ilqr_methods.jl:20,78), so notebook row i is our record with iter == i+1.
"""
import json
import os

import numpy as np

from helpers import acrobot_notebook, car_escape_notebook, printed_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tables(name):
    return json.load(open(os.path.join(GOLD, name)))["tables"]


def _check_inner(rows, recs, exact_rows, drift_rel):
    """rows: notebook rows [iter cost expected z α ρ dJ grad zero]; recs: our records after the initial one."""
    assert len(recs) >= len(rows)
    for i, row in enumerate(rows):
        r = recs[i]
        assert int(row[0]) == i + 1
        slack = 1e-9 if i < exact_rows else drift_rel  # 1e-9: last printed digit of 11-digit costs
        assert printed_close(r["alpha"], row[4]), ("alpha", i, r["alpha"], row[4])  # α sequence is bit-exact
        assert printed_close(r["cost"], row[1], slack), ("cost", i, r["cost"], row[1])
        assert printed_close(r["expected"], row[2], max(slack * 50, 0)), ("expected", i, r["expected"], row[2])
        assert printed_close(r["z"], row[3], max(slack * 50, 0)), ("z", i, r["z"], row[3])
        assert printed_close(r["dJ"], row[6], max(slack * 50, 0)), ("dJ", i, r["dJ"], row[6])
        assert printed_close(r["gradient"], row[7], max(slack * 200, 0)), ("grad", i, r["gradient"], row[7])
        assert float(row[5]) == r["rho"] == 0.0


def test_acrobot_notebook_trace(oracle):
    prob, opts = acrobot_notebook()
    res = oracle.solve(prob, opts)
    inner = res["inner"][0]
    tabs = [t for t in _tables("acrobot_notebook_trace.json") if t["columns"][1] == "cost" and t["columns"][0] == "iter"
            and "expected" in t["columns"]]
    # first outer iteration: 24 printed rows spread over three tables (the logger re-prints the header)
    first = [r for t in tabs[:3] for r in t["rows"]]
    ours = [r for r in inner if r["outer"] == 0][1:]
    # J0 (= cost + dJ of row 1 in the notebook) pins rk3 + the RBD double-pendulum restatement
    j0 = [r for r in inner if r["outer"] == 0][0]["cost"]
    assert abs(j0 - (4458.9547984 + 1277.453)) < 2e-3
    assert abs(j0 - 5736.408019391526) < 1e-6  # SURVEY Appendix F value
    # rows 1-11 agree to every printed digit; afterwards the chaotic system amplifies rounding
    # (SURVEY Appendix F item 2): allow 2e-5 relative there but still demand the identical α sequence.
    _check_inner(first, ours, exact_rows=11, drift_rel=2e-5)
    assert len(ours) >= 24


def test_car_escape_notebook_trace(oracle):
    prob, opts = car_escape_notebook()
    res = oracle.solve(prob, opts)
    inner, outer = res["inner"][0], res["outer"][0]
    tabs = _tables("car_escape_notebook_trace.json")
    outer_rows, inner_groups, cur = [], [], None
    for t in tabs:
        if "c_max" in t["columns"]:
            outer_rows += t["rows"]
            cur = None
        else:
            if cur is None:
                cur = []
                inner_groups.append(cur)
            cur += t["rows"]
    # --- outer loop of the infeasible-start solve: 6 records, c_max and cost to the printed digits
    exp = outer_rows[:6]
    for i, row in enumerate(exp):
        assert printed_close(outer[i]["c_max"], row[2]), (i, outer[i]["c_max"], row[2])
        assert printed_close(outer[i]["cost"], row[3], 2e-6), (i, outer[i]["cost"], row[3])
    # inner iteration counts per outer iteration: the notebook's `total` column excludes the
    # initial records (6, 9, 10, 11, 19); ours include them
    totals = [int(r[1]) for r in exp]
    ours_inner = [int(o["iterations_inner"]) - 1 for o in outer[1:6]]
    assert np.cumsum(ours_inner).tolist() == totals[1:]
    # --- inner rows of outer iterations 1..5
    for g in range(5):
        ours = [r for r in inner if r["outer"] == g][: ours_inner[g] + 1][1:]
        # the last inner solve runs at μ = 6.25e7 where rounding-level differences show up mid-solve
        _check_inner(inner_groups[g], ours, exact_rows=(99 if g < 4 else 1), drift_rel=2e-6)
    # --- then the feasible re-solve starts from the projected trajectory (Q9): first record
    assert printed_close(outer[6]["c_max"], outer_rows[6][2]) and printed_close(outer[6]["cost"], outer_rows[6][3])
    assert res["results"][0]["c_max"] < 1e-3 and res["results"][0]["status"] == 0
