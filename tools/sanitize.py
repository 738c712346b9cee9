"""Small solves for compute-sanitizer (memcheck / racecheck): every kernel of both engines on tiny batches."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajopt_b200 as to  # noqa: E402
from trajopt_b200 import api  # noqa: E402
from cases import CASES  # noqa: E402

names = sys.argv[1:] or ["di_altro", "pend_mintime", "park_inf_altro", "pend_sqrt_altro"]
for name in names:
    B = 3
    prob, opts, x0, X0 = CASES[name](B)
    bs = api.BatchSolver(prob, B, 0, 64, 16)
    bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), X0)
    bs.solve(opts)
    print(name, bs.results())
    bs.close()
