"""Square-root backward pass: thread-per-problem kernel against the warp-per-problem kernel (sqrt_bp.cuh) on few live problems.
Prints the device time of whole solves (the pass dominates them) and checks that the two give identical records.
Usage on the GPU box:  python tools/sqrt_pass_timing.py [B]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajopt_b200 as to  # noqa: E402,F401
from trajopt_b200 import api  # noqa: E402
from cases import CASES  # noqa: E402


def run(name, B, threshold):
    os.environ["TRAJOPT_B200_SQRT_WARP_THRESHOLD"] = str(threshold)
    prob, opts, x0, X0 = CASES[name](B)
    bs = api.BatchSolver(prob, B, 0, 0, 0)
    try:
        U0 = np.broadcast_to(prob.U, (B,) + prob.U.shape)
        best = 1e30
        for _ in range(3):
            bs.set_batch(x0, U0, X0)
            bs.solve(opts)
            best = min(best, bs.kernel_ms())
        return best, bs.results().copy(), bs.launches()
    finally:
        bs.close()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    for name in ("quad_sqrt_ilqr", "acrobot_sqrt_al", "pend_sqrt_altro", "dp_sqrt_ilqr"):
        t_thread, r_thread, _ = run(name, B, 0)
        t_warp, r_warp, _ = run(name, B, 1 << 30)
        steps = int(r_thread["steps"].max())
        print("%-18s B=%d  slowest problem %4d iterations   thread/problem %9.3f ms   warp/problem %9.3f ms   ratio %.2f   identical records: %s"
              % (name, B, steps, t_thread, t_warp, t_thread / t_warp, r_thread.tobytes() == r_warp.tobytes()), flush=True)


if __name__ == "__main__":
    main()
