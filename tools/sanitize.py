"""Small solves for compute-sanitizer (memcheck / racecheck): every kernel family of both engines on tiny batches.
    compute-sanitizer --tool memcheck python tools/sanitize.py [case ...]
Without arguments: the lockstep tick and the resident kernel on four small cases, then the kernels that only special options
reach -- both square-root passes, the projected-Newton polish, infeasible start + minimum time, bp_reg_type = :state and a
non-default gradient type."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajopt_b200 as to  # noqa: E402,F401
from trajopt_b200 import api  # noqa: E402
from cases import CASES  # noqa: E402


def run(name, B=3, env=None, mutate=None):
    saved = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        prob, opts, x0, X0 = CASES[name](B)
        if mutate:
            opts = mutate(opts)
        bs = api.BatchSolver(prob, B, 0, 64, 16)
        bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), X0)
        bs.solve(opts)
        print(name, env or "", bs.results(), flush=True)
        bs.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def with_pn(opts):
    o = api.ALTROSolverOptions(opts_al=opts.opts_al, R_inf=opts.R_inf, resolve_feasible_problem=False, projected_newton=True,
                               projected_newton_tolerance=1e-3, opts_pn=api.ProjectedNewtonSolverOptions(feasibility_tolerance=1e-8))
    return o


def with_ilqr(**kw):
    def f(opts):
        il = opts.opts_al.opts_uncon if hasattr(opts, "opts_al") else opts.opts_uncon
        for k, v in kw.items():
            setattr(il, k, v)
        return opts
    return f


if __name__ == "__main__":
    names = sys.argv[1:]
    if names:
        for nme in names:
            run(nme)
    else:
        for nme in ("di_altro", "pend_mintime", "park_inf_altro", "pend_sqrt_altro"):
            run(nme)
        run("di_altro", env={"TRAJOPT_B200_RESIDENT_THRESHOLD": "0"})                                  # lockstep tick, tail mode
        run("di_altro", env={"TRAJOPT_B200_RESIDENT_THRESHOLD": "0", "TRAJOPT_B200_TAIL_THRESHOLD": "0",
                             "TRAJOPT_B200_BP_CTA_THRESHOLD": "0"})                                       # bulk kernels
        run("dp_sqrt_ilqr", env={"TRAJOPT_B200_SQRT_WARP_THRESHOLD": "1000000"})                        # warp-per-problem sqrt pass
        run("pend_sqrt_altro", env={"TRAJOPT_B200_SQRT_WARP_THRESHOLD": "1000000"})
        run("di_altro", mutate=with_pn)                                                                  # projected Newton
        run("park_inf_mintime", B=2)                                                                     # both ALTRO transforms
        run("pend_altro", mutate=with_ilqr(bp_reg_type="state"))
        run("pend_altro", mutate=with_ilqr(gradient_type="feedforward"), env={"TRAJOPT_B200_RESIDENT_THRESHOLD": "0"})
