"""BASELINE.json's configs at their FULL batch sizes (marked `slow`): the batch is solved on the device and a spread sample of it
is compared with the CPU oracle bit for bit (result records and trajectories); size-independent properties cover the rest of the
batch (every problem ended, iteration counts within the option limits, shard results do not depend on what else is in the batch).
"""
import numpy as np
import pytest

from cases import CASES

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

# (case, BASELINE batch size, oracle sample size)
CONFIGS = [
    ("pend_ilqr", 4096, 16),
    ("cart_ilqr", 4096, 16),
    ("quad_altro", 65536, 8),
    ("escape_altro", 16384, 3),
    ("park_inf_altro", 16384, 8),
    ("acrobot_sqrt_mintime", 32768, 6),
    ("dp_sqrt_mintime", 32768, 8),
]


@pytest.mark.parametrize("name,B,ns", CONFIGS)
def test_baseline_config_at_full_size(to, oracle, name, B, ns):
    prob, opts, x0, X0 = CASES[name](B)
    bs = to.api.BatchSolver(prob, B, 0, 0, 0)
    try:
        bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), X0)
        bs.solve(opts)
        res = bs.results()
        X, U, _ = bs.solution()
    finally:
        bs.close()
    idx = np.unique(np.linspace(0, B - 1, ns).astype(int))
    ref = oracle.solve(prob, opts, x0=x0[idx], X0=None if X0 is None else X0[idx], B=len(idx), inner_cap=0, outer_cap=0,
                       threads=min(len(idx), oracle.hw_threads()))
    assert ref["results"].tobytes() == res[idx].tobytes()          # integer AND float observables of the record: bit-exact
    assert np.array_equal(ref["X"], X[idx], equal_nan=True) and np.array_equal(ref["U"], U[idx], equal_nan=True)
    # the whole batch: every problem ended with a defined status, within the iteration limits of the options
    al = opts.opts_al if hasattr(opts, "opts_al") else opts
    il = al.opts_uncon if hasattr(al, "opts_uncon") else al
    outer_max = al.iterations if hasattr(al, "opts_uncon") else 1
    assert np.all(res["steps"] >= 0) and np.all(res["steps"] <= outer_max * il.iterations * 2)
    assert np.all((res["status"] & ~(1 | 2 | 4 | 8 | 16 | 32)) == 0)
    fin = (res["status"] & (2 | 4 | 32)) == 0
    assert np.all(np.isfinite(res["J"][fin]))
