"""How close is the device projected-Newton polish to the oracle's?  (GPU box)  python tools/pn_diff.py [case] [B]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import trajopt_b200 as to  # noqa: E402
from trajopt_b200 import api  # noqa: E402
from cases import CASES  # noqa: E402
import oracle_py  # noqa: E402

oracle_py.build()
case = sys.argv[1] if len(sys.argv) > 1 else "quad_altro"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob, opts, x0, X0 = CASES[case](B)
o = api.ALTROSolverOptions(opts_al=opts.opts_al, R_inf=opts.R_inf, resolve_feasible_problem=False, projected_newton=True,
                           projected_newton_tolerance=1e-3, opts_pn=api.ProjectedNewtonSolverOptions(feasibility_tolerance=1e-8))
ref = oracle_py.solve(prob, o, x0=x0, X0=X0, B=B)
bs = api.BatchSolver(prob, B, 0, 0, 0)
bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), X0)
bs.solve(o)
X, U, _ = bs.solution()
res = bs.results()
print("kernel ms", bs.kernel_ms())
bs.close()
for b in range(B):
    dx, du = np.abs(X[b] - ref["X"][b]), np.abs(U[b] - ref["U"][b])
    print("problem %d: status %d/%d  J %.17g / %.17g  c_max %.3e / %.3e  X: max diff %.3e, %d of %d elements differ  U: %.3e, %d of %d" % (
        b, res["status"][b], ref["results"]["status"][b], res["J"][b], ref["results"]["J"][b], res["c_max"][b], ref["results"]["c_max"][b],
        dx.max(), int((dx > 0).sum()), dx.size, du.max(), int((du > 0).sum()), du.size))
