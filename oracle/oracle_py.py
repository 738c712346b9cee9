"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/liboracle.so).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import trajopt_b200 as to  # noqa: E402  (POD struct definitions + Problem marshalling only)
from trajopt_b200 import abi, api  # noqa: E402

LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    if force or not os.path.exists(LIB) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB)
            for f in ("oracle.cpp", "oracle_math.hpp", "oracle_models.hpp")):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.oracle_solve.argtypes = [C.POINTER(abi.TOProblemDesc), C.c_int, C.POINTER(abi.TOALTROOptions), C.c_int, vp, vp, vp,
                                   vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_int]
        L.oracle_solve.restype = C.c_int
        L.oracle_sincos.argtypes = [C.c_double, abi.c_double_p, abi.c_double_p]
        L.oracle_dynamics.argtypes = [C.c_int, vp, vp, vp]
        L.oracle_discrete.argtypes = [C.c_int, C.c_int, vp, vp, C.c_double, vp]
        L.oracle_discrete_jacobian.argtypes = [C.c_int, C.c_int, vp, vp, C.c_double, vp]
        L.oracle_spec_create.argtypes = [C.POINTER(abi.TOProblemDesc), C.c_int, C.c_int, C.POINTER(abi.TOALTROOptions)]
        L.oracle_spec_create.restype = vp
        L.oracle_spec_destroy.argtypes = [vp]
        L.oracle_spec_dims.argtypes = [vp, abi.c_int32_p, abi.c_int32_p]
        L.oracle_spec_num_rows.argtypes = [vp, C.c_int]
        L.oracle_spec_constraints.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        L.oracle_spec_stage_cost.argtypes = [vp, vp, vp]
        L.oracle_spec_stage_cost.restype = C.c_double
        L.oracle_spec_term_cost.argtypes = [vp, vp]
        L.oracle_spec_term_cost.restype = C.c_double
        L.oracle_spec_stage_expansion.argtypes = [vp] * 8
        L.oracle_spec_term_expansion.argtypes = [vp] * 4
        L.oracle_spec_dynamics.argtypes = [vp] * 6
        L.oracle_backwardpass.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double] + [vp] * 14
        L.oracle_backwardpass.restype = C.c_int
        L.oracle_hw_threads.restype = C.c_int
        _lib = L
    return _lib


def sincos(x):
    s, c = C.c_double(), C.c_double()
    lib().oracle_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def dynamics(model, x, u):
    n, _ = abi.MODEL_DIMS[model]
    x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros(n)
    lib().oracle_dynamics(model, x.ctypes.data, u.ctypes.data, out.ctypes.data)
    return out


def discrete(model, integ, x, u, dt):
    n, _ = abi.MODEL_DIMS[model]
    x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros(n)
    lib().oracle_discrete(model, integ, x.ctypes.data, u.ctypes.data, float(dt), out.ctypes.data)
    return out


def discrete_jacobian(model, integ, x, u, dt):
    n, m = abi.MODEL_DIMS[model]
    x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
    Z = np.zeros((n + m + 1, n))
    lib().oracle_discrete_jacobian(model, integ, x.ctypes.data, u.ctypes.data, float(dt), Z.ctypes.data)
    return Z.T.copy()  # n × (n+m+1)


class Spec:
    """The (possibly ALTRO-transformed) problem as the oracle sees it; primitive evaluations."""

    def __init__(self, prob, infeasible=False, min_time=False, opts=None):
        self.m = api.Marshalled(prob)
        o = (opts or api.ALTROSolverOptions()).to_c()
        self.h = lib().oracle_spec_create(C.byref(self.m.desc), int(infeasible), int(min_time), C.byref(o))
        n, mm = C.c_int32(), C.c_int32()
        lib().oracle_spec_dims(self.h, C.byref(n), C.byref(mm))
        self.n, self.mc = n.value, mm.value

    def __del__(self):
        try:
            lib().oracle_spec_destroy(self.h)
        except Exception:
            pass

    def num_rows(self, k):
        return lib().oracle_spec_num_rows(self.h, k)

    def constraints(self, k, x, u=None):
        p = self.num_rows(k)
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.zeros(self.mc) if u is None else np.ascontiguousarray(u, dtype=np.float64)
        c, jac, eq = np.zeros(p), np.zeros((p, self.n + self.mc)), np.zeros(p, dtype=np.int32)
        lib().oracle_spec_constraints(self.h, k, x.ctypes.data, u.ctypes.data, c.ctypes.data, jac.ctypes.data, eq.ctypes.data)
        return c, jac, eq

    def stage_cost(self, x, u):
        x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        return lib().oracle_spec_stage_cost(self.h, x.ctypes.data, u.ctypes.data)

    def term_cost(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().oracle_spec_term_cost(self.h, x.ctypes.data)

    def stage_expansion(self, x, u):
        n, m = self.n, self.mc
        x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        Qx, Qu, Qxx, Quu, Qux = np.zeros(n), np.zeros(m), np.zeros((n, n)), np.zeros((m, m)), np.zeros((n, m))
        lib().oracle_spec_stage_expansion(self.h, x.ctypes.data, u.ctypes.data, Qx.ctypes.data, Qu.ctypes.data,
                                          Qxx.ctypes.data, Quu.ctypes.data, Qux.ctypes.data)
        return Qx, Qu, Qxx.T.copy(), Quu.T.copy(), Qux.T.copy()  # Qux: m×n

    def term_expansion(self, x):
        n = self.n
        x = np.ascontiguousarray(x, dtype=np.float64)
        Qx, Qxx = np.zeros(n), np.zeros((n, n))
        lib().oracle_spec_term_expansion(self.h, x.ctypes.data, Qx.ctypes.data, Qxx.ctypes.data)
        return Qx, Qxx.T.copy()

    def dynamics(self, x, u, jac=True):
        n, m = self.n, self.mc
        x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        xn, A, B = np.zeros(n), np.zeros((n, n)), np.zeros((m, n))
        lib().oracle_spec_dynamics(self.h, x.ctypes.data, u.ctypes.data, xn.ctypes.data,
                                   A.ctypes.data if jac else None, B.ctypes.data if jac else None)
        return xn, A.T.copy(), B.T.copy()


def solve(prob, opts, x0=None, U0=None, X0=None, B=None, inner_cap=4096, outer_cap=128, threads=1, want_duals=False):
    """Solve B problems of prob's shape with the CPU oracle.  Returns a dict of numpy arrays."""
    m = api.Marshalled(prob)
    mode, o = api.as_altro_options(opts)
    n, mc, N = prob.model.n, prob.model.m, prob.N
    if x0 is None:
        x0 = prob.x0[None]
    x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(-1, n))
    B = x0.shape[0] if B is None else B
    x0 = np.ascontiguousarray(np.broadcast_to(x0, (B, n)))
    U0 = prob.U if U0 is None else U0
    U0 = np.ascontiguousarray(np.broadcast_to(np.asarray(U0, dtype=np.float64).reshape(-1, N - 1, mc), (B, N - 1, mc)))
    if X0 is None and not np.all(np.isnan(prob.X[0])):
        X0 = prob.X
    if X0 is not None:
        X0 = np.ascontiguousarray(np.broadcast_to(np.asarray(X0, dtype=np.float64).reshape(-1, N, n), (B, N, n)))
    X, U, dts = np.zeros((B, N, n)), np.zeros((B, N - 1, mc)), np.zeros((B, N - 1))
    res = np.zeros(B, dtype=api.RESULT_DTYPE)
    inner = np.zeros((B, max(1, inner_cap)), dtype=api.INNER_DTYPE)
    outer = np.zeros((B, max(1, outer_cap)), dtype=api.OUTER_DTYPE)
    ni, no = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    lam = mu = act = None
    if want_duals:
        sp = Spec(prob, infeasible=(mode == 2 and X0 is not None and not (opts.resolve_feasible_problem)),
                  min_time=(mode == 2 and prob.tf == 0.0), opts=opts if mode == 2 else None)
        P = sum(sp.num_rows(k) for k in range(N))
        lam, mu, act = np.zeros((B, P)), np.zeros((B, P)), np.zeros((B, P), dtype=np.uint8)
    rc = lib().oracle_solve(C.byref(m.desc), mode, C.byref(o), B, x0.ctypes.data, U0.ctypes.data,
                            None if X0 is None else X0.ctypes.data, X.ctypes.data, U.ctypes.data, dts.ctypes.data,
                            res.ctypes.data, inner.ctypes.data if inner_cap else None, ni.ctypes.data, inner_cap,
                            outer.ctypes.data if outer_cap else None, no.ctypes.data, outer_cap,
                            None if lam is None else lam.ctypes.data, None if mu is None else mu.ctypes.data,
                            None if act is None else act.ctypes.data, threads)
    assert rc == 0
    return dict(X=X, U=U, dts=dts, results=res, inner=[inner[b, : ni[b]] for b in range(B)],
                outer=[outer[b, : no[b]] for b in range(B)], lam=lam, mu=mu, act=act)


def hw_threads():
    return lib().oracle_hw_threads()
