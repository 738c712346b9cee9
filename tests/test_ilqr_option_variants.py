"""bp_reg_type = :state (src/solvers/ilqr/backward_pass.jl:38-46) and gradient_type = :feedforward / :ℓ2 / :ℓinf
(src/solvers/ilqr/ilqr_methods.jl:91-137).  No test, benchmark or example of the reference sets them, so nothing pins them beyond
the source; the oracle restates the source and the device is compared with the oracle (integers exact, floats 1e-8)."""
import numpy as np
import pytest

from trajopt_b200 import api
from cases import CASES


def _with(opts, **kw):
    """the case's options with the iLQR option(s) replaced"""
    al = opts.opts_al if isinstance(opts, api.ALTROSolverOptions) else opts
    il = al.opts_uncon if isinstance(al, api.AugmentedLagrangianSolverOptions) else al
    for k, v in kw.items():
        setattr(il, k, v)
    return opts


def test_oracle_state_regularisation_changes_only_regularised_solves(oracle):
    # cartpole ALTRO never regularises (rho stays 0): :state and :control are the same solve; the quadrotor batch does regularise
    prob, opts, x0, X0 = CASES["cart_altro"](2)
    a = oracle.solve(prob, opts, x0=x0, B=2)
    b = oracle.solve(prob, _with(CASES["cart_altro"](2)[1], bp_reg_type=":state"), x0=x0, B=2)
    assert a["results"].tobytes() == b["results"].tobytes()
    prob, opts, x0, X0 = CASES["quad_regdiv"](4)
    a = oracle.solve(prob, opts, x0=x0, B=4)
    b = oracle.solve(prob, _with(CASES["quad_regdiv"](4)[1], bp_reg_type=":state"), x0=x0, B=4)
    assert any(np.any(r["rho"] > 0) for r in a["inner"])
    assert not np.array_equal(a["results"]["iterations_total"], b["results"]["iterations_total"])


def test_oracle_gradient_types(oracle):
    prob, opts, x0, X0 = CASES["cart_altro"](1)
    base = oracle.solve(prob, opts, x0=x0, B=1)["inner"][0]
    ff = oracle.solve(prob, _with(CASES["cart_altro"](1)[1], gradient_type=":feedforward"), x0=x0, B=1)["inner"][0]
    l2 = oracle.solve(prob, _with(CASES["cart_altro"](1)[1], gradient_type=":ℓ2"), x0=x0, B=1)["inner"][0]
    li = oracle.solve(prob, _with(CASES["cart_altro"](1)[1], gradient_type=":ℓinf"), x0=x0, B=1)["inner"][0]
    # the initial record: d = 0 for :todorov and :feedforward, the cost gradient at the initial trajectory for the norms
    assert base["gradient"][0] == 0.0 and ff["gradient"][0] == 0.0 and l2["gradient"][0] > 0.0
    # m = 1: :todorov = mean_k |d|/(|u|+1) <= max_k |d| = :feedforward ; ||g||_inf <= ||g||_2
    n = min(len(base), len(ff), len(l2), len(li))
    assert np.all(ff["gradient"][1:n] >= base["gradient"][1:n]) and np.all(li["gradient"][:n] <= l2["gradient"][:n])
    assert np.array_equal(base["cost"][:5], ff["cost"][:5])   # the gradient only feeds the convergence test


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(bp_reg_type=":state"), dict(gradient_type=":feedforward"), dict(gradient_type=":ℓ2"),
                                dict(gradient_type=":ℓinf"), dict(bp_reg_type=":state", gradient_type=":ℓinf")],
                         ids=["state", "feedforward", "l2", "linf", "state+linf"])
@pytest.mark.parametrize("name", ["quad_regdiv", "cart_altro", "park_inf_altro", "pend_ilqr"])
def test_gpu_option_variants_match_oracle(to, oracle, name, kw):
    B = 6
    prob, opts, x0, X0 = CASES[name](B)
    opts = _with(opts, **kw)
    ref = oracle.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96)
    bs = to.api.BatchSolver(prob, B, 0, 2048, 96)
    try:
        bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), X0)
        bs.solve(opts)
        X, U, _ = bs.solution()
        res = bs.results()
        inner, _ = bs.trace()
    finally:
        bs.close()
    for f in ("iterations_total", "iterations_outer", "status", "steps"):
        assert np.array_equal(ref["results"][f], res[f]), (f, ref["results"][f], res[f])
    fin = np.isfinite(ref["results"]["J"])
    assert np.allclose(ref["results"]["J"][fin], res["J"][fin], rtol=1e-8, atol=0)
    ok = np.isfinite(ref["X"]).all(axis=(1, 2))
    assert np.allclose(ref["X"][ok], X[ok], rtol=1e-8, atol=1e-10) and np.allclose(ref["U"][ok], U[ok], rtol=1e-8, atol=1e-10)
    for b in range(B):
        assert len(ref["inner"][b]) == len(inner[b])
        g0, g1 = ref["inner"][b]["gradient"], inner[b]["gradient"]
        both = np.isfinite(g0) & np.isfinite(g1)
        assert np.array_equal(np.isfinite(g0), np.isfinite(g1)) and np.allclose(g0[both], g1[both], rtol=1e-8, atol=0)


@pytest.mark.gpu
def test_state_regularisation_on_the_second_engine(to, oracle, monkeypatch):
    """the warp-resident engine implements :state too (same oracle)"""
    monkeypatch.setenv("TRAJOPT_B200_ENGINE", "persistent")
    prob, opts, x0, X0 = CASES["quad_regdiv"](4)
    opts = _with(opts, bp_reg_type=":state")
    ref = oracle.solve(prob, opts, x0=x0, B=4, inner_cap=0, outer_cap=0)
    bs = to.api.BatchSolver(prob, 4, 0, 0, 0)
    try:
        bs.set_batch(x0, np.broadcast_to(prob.U, (4,) + prob.U.shape), None)
        bs.solve(opts)
        res = bs.results()
    finally:
        bs.close()
    assert ref["results"].tobytes() == res.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(bp_reg_type=":state", gradient_type=":ℓinf"), dict(gradient_type=":ℓ2"),
                                dict(gradient_type=":feedforward")], ids=["state+linf", "l2", "feedforward"])
def test_gpu_option_variants_on_the_lockstep_tick(to, oracle, monkeypatch, kw):
    """the same options with the resident kernel off: CTA-per-problem backward pass launched per tick (it serves :state for any
    number of live problems), gradient in the accept kernel"""
    monkeypatch.setenv("TRAJOPT_B200_RESIDENT_THRESHOLD", "0")
    B = 6
    prob, opts, x0, X0 = CASES["quad_regdiv"](B)
    opts = _with(opts, **kw)
    ref = oracle.solve(prob, opts, x0=x0, B=B, inner_cap=0, outer_cap=0)
    bs = to.api.BatchSolver(prob, B, 0, 0, 0)
    try:
        bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), None)
        bs.solve(opts)
        res = bs.results()
    finally:
        bs.close()
    for f in ("iterations_total", "iterations_outer", "status", "steps"):
        assert np.array_equal(ref["results"][f], res[f]), (f, ref["results"][f], res[f])
    fin = np.isfinite(ref["results"]["J"])
    assert np.allclose(ref["results"]["J"][fin], res["J"][fin], rtol=1e-8, atol=0)
