"""trajopt_b200 — B200-native batched iLQR / AL-iLQR / ALTRO engine behind TrajectoryOptimization.jl's API.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of include/trajopt_b200.h),
`abi.py` (ctypes mirror of that header), `api.py` (host mirror of the reference's user API) and
`problems.py` (the reference's problem zoo as data).
"""
from . import abi, api, problems, sharding  # noqa: F401
from .api import *  # noqa: F401,F403
