"""ALTRO WITH the projected-Newton polish (the reference's own benchmark setting, benchmark/quadrotor_benchmarks.jl:
projected_newton = true) on a quadrotor batch: solves/s, final constraint violation, statuses, the share of the polish, and a
sample checked against the CPU oracle.  The headline metric of bench.py is the AL phase alone (PN off), see BASELINE.md.
Usage on the GPU box:  python tools/pn_throughput.py [B]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajopt_b200 as to  # noqa: E402,F401
from trajopt_b200 import api  # noqa: E402
import oracle_py  # noqa: E402
from cases import CASES  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    prob, base, x0, _ = CASES["quad_altro"](B)
    pn = api.ALTROSolverOptions(opts_al=base.opts_al, R_inf=base.R_inf, resolve_feasible_problem=False, projected_newton=True,
                                projected_newton_tolerance=1e-3, opts_pn=api.ProjectedNewtonSolverOptions(feasibility_tolerance=1e-8))
    U0 = np.broadcast_to(prob.U, (B,) + prob.U.shape)
    out = {}
    for label, opts in (("al_only", base), ("al_plus_pn", pn)):
        bs = api.BatchSolver(prob, B, 0, 0, 0)
        try:
            bs.set_batch(x0, U0, None)
            bs.solve(opts)  # warm-up
            bs.set_batch(x0, U0, None)
            t0 = time.time()
            bs.solve(opts)
            wall = time.time() - t0
            ms = bs.kernel_ms()
            res = bs.results().copy()
            X, U, _ = bs.solution()
        finally:
            bs.close()
        st, cnt = np.unique(res["status"], return_counts=True)
        out[label] = {"device_ms": ms, "wall_s": wall, "solves_per_s": B / (ms * 1e-3),
                      "c_max_median": float(np.median(res["c_max"])), "c_max_max_of_status0": float(res["c_max"][res["status"] == 0].max()),
                      "status_histogram": {str(int(k)): int(v) for k, v in zip(st, cnt)}}
        if label == "al_plus_pn":
            idx = np.unique(np.linspace(0, B - 1, 6).astype(int))
            ref = oracle_py.solve(prob, pn, x0=x0[idx], B=len(idx), inner_cap=0, outer_cap=0)
            out[label]["oracle_sample"] = [int(i) for i in idx]
            out[label]["sample_status_iterations_equal"] = bool(all(np.array_equal(ref["results"][f], res[idx][f])
                                                                    for f in ("status", "iterations_total", "iterations_outer", "steps")))
            out[label]["sample_XU_max_abs_diff"] = float(max(np.nanmax(np.abs(ref["X"] - X[idx])), np.nanmax(np.abs(ref["U"] - U[idx]))))
    out["batch"] = B
    out["pn_share_of_device_time"] = 1.0 - out["al_only"]["device_ms"] / out["al_plus_pn"]["device_ms"]
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pn_throughput_b%d.json" % B), "w"), indent=1)


if __name__ == "__main__":
    main()
