"""ctypes mirror of include/trajopt_b200.h (the C-ABI drop-in boundary) and the library loader.

The product library is `csrc/libtrajopt_b200.so` (hand-written CUDA for sm_100a).  There is NO CPU
fallback: `load_library()` raises if the extension is missing, and every solve call needs a GPU.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libtrajopt_b200.so")

# models (dynamics/*.jl)
MODEL_DOUBLE_INTEGRATOR, MODEL_PENDULUM, MODEL_CAR, MODEL_CARTPOLE = 0, 1, 2, 3
MODEL_QUADROTOR, MODEL_ACROBOT, MODEL_DOUBLEPENDULUM = 4, 5, 6
MODEL_DIMS = {0: (2, 1), 1: (2, 1), 2: (3, 2), 3: (4, 1), 4: (13, 4), 5: (4, 1), 6: (4, 2)}
MODEL_NAMES = {0: "doubleintegrator", 1: "pendulum", 2: "car", 3: "cartpole", 4: "quadrotor",
               5: "acrobot", 6: "doublependulum"}
INTEG_RK3, INTEG_RK4, INTEG_MIDPOINT = 0, 1, 2
ROW_LINEAR, ROW_CIRCLE, ROW_SPHERE = 0, 1, 2

STATUS_OK, STATUS_COST_BLOWUP, STATUS_COST_INCREASED = 0, 1, 2
STATUS_NOT_PD_SQRT, STATUS_MAX_OUTER, STATUS_TRACE_TRUNC = 4, 8, 16

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class TOConstraintRow(C.Structure):
    _fields_ = [("kind", C.c_int32), ("equality", C.c_int32), ("var", C.c_int32), ("is_bound", C.c_int32),
                ("sign", C.c_double), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double), ("r", C.c_double)]


class TOProblemDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("integrator", C.c_int32), ("n", C.c_int32), ("m", C.c_int32),
                ("N", C.c_int32), ("reserved0", C.c_int32), ("dt", C.c_double), ("tf", C.c_double),
                ("Q", c_double_p), ("R", c_double_p), ("H", c_double_p), ("q", c_double_p), ("r", c_double_p),
                ("c", C.c_double), ("Qf", c_double_p), ("qf", c_double_p), ("cf", C.c_double),
                ("n_classes", C.c_int32), ("reserved1", C.c_int32), ("class_of_knot", c_int32_p),
                ("class_row_start", c_int32_p), ("rows", C.POINTER(TOConstraintRow))]


class TOiLQROptions(C.Structure):
    _fields_ = [("cost_tolerance", C.c_double), ("gradient_norm_tolerance", C.c_double),
                ("iterations", C.c_int32), ("dJ_counter_limit", C.c_int32), ("square_root", C.c_int32),
                ("iterations_linesearch", C.c_int32), ("line_search_lower_bound", C.c_double),
                ("line_search_upper_bound", C.c_double), ("bp_reg_increase_factor", C.c_double),
                ("bp_reg_max", C.c_double), ("bp_reg_min", C.c_double), ("bp_reg_fp", C.c_double),
                ("max_cost_value", C.c_double), ("max_state_value", C.c_double), ("max_control_value", C.c_double),
                ("bp_reg_type", C.c_int32), ("gradient_type", C.c_int32)]


class TOALOptions(C.Structure):
    _fields_ = [("opts_uncon", TOiLQROptions), ("cost_tolerance", C.c_double),
                ("cost_tolerance_intermediate", C.c_double), ("gradient_norm_tolerance", C.c_double),
                ("gradient_norm_tolerance_intermediate", C.c_double), ("constraint_tolerance", C.c_double),
                ("iterations", C.c_int32), ("kickout_max_penalty", C.c_int32), ("dual_min", C.c_double),
                ("dual_max", C.c_double), ("penalty_max", C.c_double), ("penalty_initial", C.c_double),
                ("penalty_scaling", C.c_double)]


class TOALTROOptions(C.Structure):
    _fields_ = [("opts_al", TOALOptions), ("R_inf", C.c_double), ("dynamically_feasible_projection", C.c_int32),
                ("resolve_feasible_problem", C.c_int32), ("R_minimum_time", C.c_double), ("dt_max", C.c_double),
                ("dt_min", C.c_double), ("projected_newton", C.c_int32), ("pn_n_steps", C.c_int32),
                ("projected_newton_tolerance", C.c_double), ("pn_feasibility_tolerance", C.c_double),
                ("pn_active_set_tolerance", C.c_double)]


class TOResult(C.Structure):
    _fields_ = [("J", C.c_double), ("c_max", C.c_double), ("iterations_total", C.c_int32),
                ("iterations_outer", C.c_int32), ("status", C.c_int32), ("steps", C.c_int32)]


class TOIterRecord(C.Structure):
    _fields_ = [("cost", C.c_double), ("dJ", C.c_double), ("gradient", C.c_double), ("expected", C.c_double),
                ("z", C.c_double), ("alpha", C.c_double), ("rho", C.c_double), ("outer", C.c_int32),
                ("iter", C.c_int32)]


class TOOuterRecord(C.Structure):
    _fields_ = [("cost", C.c_double), ("c_max", C.c_double), ("penalty_max", C.c_double),
                ("iterations_inner", C.c_int32), ("pad", C.c_int32)]


# every symbol include/trajopt_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "to_default_ilqr_options", "to_default_al_options", "to_default_altro_options", "to_create", "to_destroy",
    "to_last_error", "to_set_batch", "to_set_batch_device", "to_warm_start_shift", "to_set_trace", "to_solve_ilqr", "to_solve_al",
    "to_solve_altro", "to_solve_altro_async", "to_sync", "to_last_kernel_ms", "to_last_launch_count",
    "to_get_solution", "to_get_results", "to_results_device_ptr", "to_get_trace", "to_num_constraint_rows",
    "to_get_duals", "to_stream", "to_copy_results_device", "to_last_linesearch_trials", "to_measure_fp64_peak",
    "to_device_count", "to_version",
]

_lib = None


def load_library(path=None):
    """Load the CUDA engine.  Raises (loudly) if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "trajopt_b200: CUDA extension %s is missing — run `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU fallback." % p)
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.to_default_ilqr_options.argtypes = [C.POINTER(TOiLQROptions)]
    lib.to_default_al_options.argtypes = [C.POINTER(TOALOptions)]
    lib.to_default_altro_options.argtypes = [C.POINTER(TOALTROOptions)]
    lib.to_create.argtypes = [C.POINTER(TOProblemDesc), C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.to_create.restype = C.c_int
    lib.to_destroy.argtypes = [vp]
    lib.to_destroy.restype = None
    lib.to_last_error.argtypes = [vp]
    lib.to_last_error.restype = C.c_char_p
    lib.to_set_batch.argtypes = [vp, vp, vp, vp]
    lib.to_set_batch_device.argtypes = [vp, vp, vp, vp]
    lib.to_warm_start_shift.argtypes = [vp, vp, C.c_int32]
    lib.to_set_trace.argtypes = [vp, C.c_int32, C.c_int32]
    lib.to_solve_ilqr.argtypes = [vp, C.POINTER(TOiLQROptions)]
    lib.to_solve_al.argtypes = [vp, C.POINTER(TOALOptions)]
    lib.to_solve_altro.argtypes = [vp, C.POINTER(TOALTROOptions)]
    lib.to_solve_altro_async.argtypes = [vp, C.POINTER(TOALTROOptions)]
    lib.to_sync.argtypes = [vp]
    lib.to_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.to_last_launch_count.argtypes = [vp, c_int32_p]
    lib.to_get_solution.argtypes = [vp, vp, vp, vp]
    lib.to_get_results.argtypes = [vp, vp]
    lib.to_results_device_ptr.argtypes = [vp, C.POINTER(vp)]
    lib.to_get_trace.argtypes = [vp, vp, vp, vp, vp]
    lib.to_num_constraint_rows.argtypes = [vp, c_int32_p]
    lib.to_get_duals.argtypes = [vp, vp, vp, vp]
    lib.to_stream.argtypes = [vp, C.POINTER(vp)]
    lib.to_copy_results_device.argtypes = [vp, vp]
    lib.to_last_linesearch_trials.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.to_measure_fp64_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    lib.to_device_count.restype = C.c_int
    lib.to_version.restype = C.c_char_p
    for name in ["to_set_batch", "to_set_batch_device", "to_warm_start_shift", "to_set_trace", "to_solve_ilqr", "to_solve_al", "to_solve_altro",
                 "to_solve_altro_async", "to_sync", "to_last_kernel_ms", "to_last_launch_count", "to_get_solution",
                 "to_get_results", "to_results_device_ptr", "to_get_trace", "to_num_constraint_rows", "to_get_duals",
                 "to_stream", "to_copy_results_device", "to_last_linesearch_trials", "to_measure_fp64_peak"]:
        getattr(lib, name).restype = C.c_int
    if path is None:
        _lib = lib
    return lib
