"""Cycle profile of the lockstep backward pass (one lane group): sections of the knot loop.
Usage on the GPU box: python tools/bp_profile.py [B]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trajopt_b200 as to  # noqa: E402
from trajopt_b200 import api, problems  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prob, opts = problems.quadrotor(), problems.quadrotor_bench_options()
bs = api.BatchSolver(prob, B, 0, 0, 0)
lib = bs.lib
bs.set_batch(problems.batch_x0("quadrotor", B), np.broadcast_to(prob.U, (B,) + prob.U.shape))
assert lib.to_debug_enable(bs.h, 64) == 0
bs.solve(opts)
buf = np.zeros(64, dtype=np.int64)
assert lib.to_debug_read(bs.h, buf.ctypes.data, 64) == 0
names = ["wait+sync", "phaseA T=A'S", "phaseB T*A", "expansion + Q add", "publish+sync+prefetch", "chol+LU", "solves+KQ+sync", "S update+sync", "sym+dV"]
tot = buf[:9].sum()
ticks = C.c_int32()
lib.to_debug_ticks(bs.h, C.byref(ticks))
res = bs.results()
print("B=%d ticks=%d kernel %.1f ms; group 0 handled problem list slot 0 in every tick" % (B, ticks.value, bs.kernel_ms()))
knots = ticks.value * (prob.N - 1)
for nme, v in zip(names, buf[:9]):
    print("  %-18s %12d cycles  %5.1f%%  ~%7.0f cycles/knot" % (nme, v, 100.0 * v / max(1, tot), v / max(1, knots)))
print("  total %d cycles ~ %.0f cycles/knot (upper bound on knots: slot 0 is not live in every tick)" % (tot, tot / max(1, knots)))
bs.close()
