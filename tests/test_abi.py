"""CPU-only: the C-ABI library loads and exports every symbol include/trajopt_b200.h declares; host
marshalling and argument validation work without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(to):
    lib = to.abi.load_library()
    hdr = open(os.path.join(ROOT, "include", "trajopt_b200.h")).read()
    declared = set(re.findall(r"\b(to_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(to.abi.EXPORTS), declared ^ set(to.abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_header(to, tmp_path):
    """ctypes mirrors == what a C compiler sees in include/trajopt_b200.h"""
    import subprocess
    a = to.abi
    names = ["TOConstraintRow", "TOProblemDesc", "TOiLQROptions", "TOALOptions", "TOALTROOptions", "TOResult", "TOIterRecord", "TOOuterRecord"]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "trajopt_b200.h"\nint main(){' +
                   "".join('printf("%%zu\\n", sizeof(%s));' % n for n in names) + "return 0;}")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(getattr(a, n)) for n in names]
    assert C.sizeof(a.TOResult) == 32 and C.sizeof(a.TOConstraintRow) == 56


def test_defaults_match_reference(to):
    lib = to.abi.load_library()
    o = to.abi.TOALTROOptions()
    lib.to_default_altro_options(C.byref(o))
    py = to.ALTROSolverOptions().to_c()
    assert bytes(o) == bytes(py)
    il = o.opts_al.opts_uncon
    assert (il.cost_tolerance, il.gradient_norm_tolerance, il.iterations, il.dJ_counter_limit) == (1e-4, 1e-5, 300, 10)
    assert (il.iterations_linesearch, il.line_search_lower_bound, il.line_search_upper_bound) == (20, 1e-8, 10.0)
    assert (il.bp_reg_increase_factor, il.bp_reg_min, il.bp_reg_fp, il.max_cost_value) == (1.6, 1e-8, 10.0, 1e8)
    al = o.opts_al
    assert (al.cost_tolerance_intermediate, al.constraint_tolerance, al.iterations, al.penalty_scaling) == (1e-3, 1e-3, 30, 10.0)
    assert (o.R_inf, o.R_minimum_time, o.dt_max, o.dt_min, o.resolve_feasible_problem) == (1.0, 1.0, 1.0, 1e-3, 1)


def test_marshalling_quadrotor(to):
    p = to.problems.quadrotor()
    m = to.api.Marshalled(p)
    d = m.desc
    assert (d.model, d.n, d.m, d.N) == (4, 13, 4, 101) and abs(d.dt - 0.05) < 1e-15
    assert d.n_classes == 2 and m.class_row_start.tolist() == [0, 4, 22]
    assert p.constraints.num_constraints() == [4] * 100 + [18]
    rows = [m.rows[i] for i in range(22)]
    assert all(r.var == 13 + i and r.sign == -1.0 and r.a == 0.0 and r.is_bound for i, r in enumerate(rows[:4]))  # u_min - u
    assert [r.var for r in rows[4:13]] == [0, 1, 2, 7, 8, 9, 10, 11, 12] and all(r.sign == 1.0 for r in rows[4:13])
    assert [r.a for r in rows[4:7]] == [0.0, 60.0, 10.0]


def test_time_validation(to):
    """test/problem_tests.jl:53-55,77-80"""
    from trajopt_b200.api import _validate_time
    assert _validate_time(11, 3.0, float("nan"))[2] == pytest.approx(0.3)
    N, tf, dt = _validate_time(51, float("nan"), 0.05)
    assert dt == 0.05 and tf == pytest.approx(2.5)
    with pytest.raises(ValueError):
        to.BoundConstraint(2, 1, u_max=-1.0, u_min=1.0)


def test_create_fails_loudly_without_gpu_or_on_bad_input(to):
    lib = to.abi.load_library()
    p = to.problems.doubleintegrator()
    m = to.api.Marshalled(p)
    h = C.c_void_p()
    bad = to.abi.TOProblemDesc.from_buffer_copy(m.desc)
    bad.n = 3
    assert lib.to_create(C.byref(bad), 1, 0, C.byref(h)) == -1  # TO_ERR_INVALID: n/m do not match the model
    assert b"match" in lib.to_last_error(None)
    if lib.to_device_count() == 0:
        rc = lib.to_create(C.byref(m.desc), 1, 0, C.byref(h))
        assert rc == -3 and b"no CPU fallback" in lib.to_last_error(None)  # TO_ERR_CUDA: there is no CPU path
        with pytest.raises(RuntimeError):
            to.solve_b(p, to.ALTROSolverOptions())


def test_missing_extension_raises(to, tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        to.abi.load_library(str(tmp_path / "libtrajopt_b200.so"))


def test_synthetic_inputs_are_deterministic(to):
    a = to.problems.batch_x0("quadrotor", 4)
    b = to.problems.batch_x0("quadrotor", 2, offset=2)
    assert np.array_equal(a[2:], b) and np.allclose(np.linalg.norm(a[:, 3:7], axis=1), 1.0)
    u = to.problems.splitmix_uniform(1, 1000)
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.05
