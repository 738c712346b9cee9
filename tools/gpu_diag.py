"""Diagnostic (never asserts): run every parity case on the GPU and on the CPU oracle and print where
they first differ.  Usage on the GPU box:  python tools/gpu_diag.py [case ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajopt_b200 as to  # noqa: E402
from trajopt_b200 import api, problems  # noqa: E402
import oracle_py  # noqa: E402
from cases import CASES  # noqa: E402


def first_diff(a, b, fields):
    n = min(len(a), len(b))
    for i in range(n):
        for f in fields:
            x, y = a[i][f], b[i][f]
            if not (x == y or (x != x and y != y)):
                return i, f, x, y
    if len(a) != len(b):
        return n, "len", len(a), len(b)
    return None


def run_case(name, B):
    prob, opts, x0, X0 = CASES[name](B)
    t0 = time.time()
    ref = oracle_py.solve(prob, opts, x0=x0, X0=X0, B=B, inner_cap=2048, outer_cap=96, want_duals=False)
    t_cpu = time.time() - t0
    bs = api.BatchSolver(prob, B, 0, 2048, 96)
    U0 = np.broadcast_to(prob.U, (B,) + prob.U.shape)
    bs.set_batch(x0, U0, X0)
    t0 = time.time()
    bs.solve(opts)
    t_gpu = time.time() - t0
    res = bs.results()
    X, U, dts = bs.solution()
    inner, outer = bs.trace()
    ms = bs.kernel_ms()
    print("== %-22s B=%d  cpu %.3fs  gpu wall %.3fs kernel %.2f ms launches %d" % (name, B, t_cpu, t_gpu, ms, bs.launches()))
    nbad = 0
    for b in range(B):
        r, g = ref["results"][b], res[b]
        same_int = all(r[f] == g[f] for f in ("iterations_total", "iterations_outer", "status", "steps"))
        relJ = abs(r["J"] - g["J"]) / max(1e-300, abs(r["J"]))
        dX = np.max(np.abs(ref["X"][b] - X[b])) if np.all(np.isfinite(X[b])) else np.inf
        dU = np.max(np.abs(ref["U"][b] - U[b])) if np.all(np.isfinite(U[b])) else np.inf
        fd = first_diff(ref["inner"][b], inner[b], ("iter", "outer", "alpha", "cost", "dJ", "gradient", "expected", "z", "rho"))
        bit = fd is None and same_int and r["J"] == g["J"] and dX == 0 and dU == 0
        if not bit:
            nbad += 1
            if nbad <= 3:
                print("  b=%d int_equal=%s ref=%s gpu=%s relJ=%.2e dX=%.2e dU=%.2e first_inner_diff=%s" % (b, same_int, r, g, relJ, dX, dU, fd))
                fo = first_diff(ref["outer"][b], outer[b], ("cost", "c_max", "penalty_max", "iterations_inner"))
                print("     first_outer_diff=%s" % (fo,))
    print("   bit-exact problems: %d / %d" % (B - nbad, B))
    bs.close()
    return nbad


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    tot = 0
    for nm in names:
        B = 8
        if ":" in nm:
            nm, B = nm.split(":")
            B = int(B)
        try:
            tot += run_case(nm, B)
        except Exception as e:  # keep going: this is a diagnostic
            import traceback
            traceback.print_exc()
            tot += 1
    print("TOTAL mismatching problems:", tot)
