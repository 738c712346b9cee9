"""Cycle profile of the CTA-per-problem resident kernel (resident.cuh): cycles per phase of the iLQR iteration, accumulated by
CTA 0 over every iteration of the problems it serves.  Usage on the GPU box: python tools/resident_profile.py [case] [B]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajopt_b200 as to  # noqa: E402
from trajopt_b200 import api  # noqa: E402
from cases import CASES  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "quad_altro"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prob, opts, x0, X0 = CASES[case](B)
bs = api.BatchSolver(prob, B, 0, 0, 0)
lib = bs.lib
bs.set_batch(x0, np.broadcast_to(prob.U, (B,) + prob.U.shape), X0)
assert lib.to_debug_enable(bs.h, 96) == 0
bs.solve(opts)
buf = np.zeros(96, dtype=np.int64)
assert lib.to_debug_read(bs.h, buf.ctypes.data, 96) == 0
names = ["jacobians", "expansion", "riccati", "T1 state chains", "T2 costs", "T3+accept 1", "copy", "accept 2 (+outer)"]
prof = buf[16:32]
iters = max(1, int(prof[15]))
tot = prof[:8].sum()
print("case %s B=%d: kernel %.2f ms; CTA 0 ran %d iLQR iterations, %.1f us per iteration at 1.965 GHz" %
      (case, B, bs.kernel_ms(), iters, tot / iters / 1965.0))
for nme, v in zip(names, prof[:8]):
    print("  %-20s %14d cycles  %5.1f%%  %9.0f cycles/iteration  %7.1f us" % (nme, v, 100.0 * v / max(1, tot), v / iters, v / iters / 1965.0))
rn = ["S update of k+1 (steps 4+5)", "wait: inputs + block barrier", "step 1: T, Tu, A'Sx, B'Sx", "wait: block barrier", "step 2: own product task",
      "Qx, Qu, LU factorisation", "wait: Qux of the other warps", "solves, K, d, KQ", "wait: block barrier (PD test)"]
sub = buf[48:57]
if sub.sum() > 0:
    knots = iters * (prob.N - 1)
    print("  Riccati recursion, thread 0 (warp 0 = the factorisation warp), cycles per knot:")
    for nme, v in zip(rn, sub):
        print("    %-34s %8.0f" % (nme, v / knots))
    print("    %-34s %8.0f" % ("sum", sub.sum() / knots))
tn = ["knot top (box flags, prefetch)", "control law + stage 1", "wait: role barrier", "stage 2", "wait: role barrier", "stage 3 + box test",
      "wait: inputs + role barrier"]
sub = buf[64:71]
if sub.sum() > 0:
    knots = iters * (prob.N - 1)
    print("  line-search state chain, thread 0 (role warp 0), cycles per knot:")
    for nme, v in zip(tn, sub):
        print("    %-34s %8.0f" % (nme, v / knots))
    print("    %-34s %8.0f" % ("sum", sub.sum() / knots))
bs.close()
