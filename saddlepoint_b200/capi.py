"""ctypes declarations for libsaddlepoint_b200.so (the C ABI in include/saddlepoint_b200.h).

The library is built in-tree by saddlepoint_b200.build; loading fails loudly if it is
missing -- there is no Python or CPU fallback for any entry point.
"""
import ctypes as C
import os

from . import build as _build

SPB_OK, SPB_NOT_SPD, SPB_ERR_ARG, SPB_ERR_CUDA, SPB_ERR_STATE, SPB_NOT_CONVERGED, SPB_ERR_NCCL = range(7)
SPB_HOST, SPB_DEVICE = 0, 1
SOLVER_CHOLESKY, SOLVER_PCG_JACOBI, SOLVER_PCG_IC0 = 0, 1, 2
STATUS_NAMES = ["SPB_OK", "SPB_NOT_SPD", "SPB_ERR_ARG", "SPB_ERR_CUDA", "SPB_ERR_STATE", "SPB_NOT_CONVERGED", "SPB_ERR_NCCL"]

vp = C.c_void_p

INIT_CB = C.CFUNCTYPE(None, vp, C.POINTER(C.c_double))
OBJJAC_CB = C.CFUNCTYPE(None, vp, vp, vp, vp, vp)
PRE_CB = C.CFUNCTYPE(None, vp, vp)
POST_CB = C.CFUNCTYPE(C.c_int, vp, vp)


class Traits(C.Structure):
    _fields_ = [("xSize", C.c_int32), ("ESize", C.c_int32), ("colptr", vp), ("rowidx", vp), ("pattern_version", C.c_uint64), ("row_begin", C.c_int32), ("row_end", C.c_int32), ("mem", C.c_int),
                ("user", vp), ("initial_solution", INIT_CB), ("objective_jacobian", OBJJAC_CB),
                ("pre_iteration", PRE_CB), ("post_iteration", POST_CB)]


class CheckResult(C.Structure):
    _fields_ = [("max_diff", C.c_double), ("max_row", C.c_int32), ("max_col", C.c_int32), ("discrepancies", C.c_int64),
                ("outside_max", C.c_double), ("outside_count", C.c_int64), ("ncolors", C.c_int32), ("evaluations", C.c_int32)]


class LMOptions(C.Structure):
    _fields_ = [("maxIterations", C.c_int32), ("funcTolerance", C.c_double), ("fooTolerance", C.c_double),
                ("lambda0", C.c_double), ("verbose", C.c_int32), ("dedupe_evaluations", C.c_int32),
                ("fixed_iterations", C.c_int32)]


class LMResult(C.Structure):
    _fields_ = [("ret", C.c_int32), ("exit_path", C.c_int32), ("currIter", C.c_int32), ("factorizations", C.c_int32),
                ("energy", C.c_double), ("fooOptimality", C.c_double), ("lambda_", C.c_double),
                ("linear_iterations", C.c_int64), ("seconds_total", C.c_double), ("seconds_analyze", C.c_double),
                ("unconverged_solves", C.c_int64)]


# every symbol include/saddlepoint_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("spb_create", C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    ("spb_destroy", C.c_int, [vp]),
    ("spb_last_error", C.c_char_p, [vp]),
    ("spb_version", C.c_char_p, []),
    ("spb_set_solver", C.c_int, [vp, C.c_int, C.c_double, C.c_int]),
    ("spb_analyze", C.c_int, [vp, C.c_int32, C.c_int32, vp, vp]),
    ("spb_analyze_update", C.c_int, [vp, C.c_int32, C.c_int32, vp, vp, C.POINTER(C.c_int)]),
    ("spb_pattern_matches", C.c_int, [vp, C.c_int32, C.c_int32, vp, vp, C.POINTER(C.c_int)]),
    ("spb_h_nnz", C.c_int, [vp, C.POINTER(C.c_int64)]),
    ("spb_h_pattern", C.c_int, [vp, vp, vp]),
    ("spb_h_values", C.c_int, [vp, vp, C.c_int]),
    ("spb_symbolic_h_pattern", C.c_int, [C.c_int32, C.c_int32, vp, vp, C.POINTER(C.c_int64), vp, vp]),
    ("spb_symbolic_plan_check", C.c_int, [C.c_int32, C.c_int32, vp, vp, vp]),
    ("spb_symbolic_pairs_v2", C.c_int, [C.c_int32, C.c_int32, vp, vp, vp, vp, vp]),
    ("spb_symbolic_chol_stats", C.c_int, [C.c_int32, C.c_int32, vp, vp, vp, vp]),
    ("spb_assemble", C.c_int, [vp, vp, vp, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    ("spb_get_rhs", C.c_int, [vp, vp, C.c_int]),
    ("spb_get_damping", C.c_int, [vp, vp, C.c_int]),
    ("spb_set_matrix", C.c_int, [vp, C.c_int32, vp, vp, vp, C.c_int]),
    ("spb_analyze_triplets", C.c_int, [vp, C.c_int32, C.c_int64, vp, vp, C.c_int]),
    ("spb_factorize_triplets", C.c_int, [vp, vp, C.c_int]),
    ("spb_factorize", C.c_int, [vp]),
    ("spb_solve", C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    ("spb_factor_info", C.c_int, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("spb_squared_norm", C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(C.c_double)]),
    ("spb_inf_norm", C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(C.c_double)]),
    ("spb_spmv", C.c_int, [vp, vp, vp, C.c_int]),
    ("spb_residual_check", C.c_int, [vp, vp, C.c_int, C.POINTER(C.c_double)]),
    ("spb_lm_default_options", None, [C.POINTER(LMOptions)]),
    ("spb_check_traits", C.c_int, [vp, vp, vp, C.c_int, C.c_double, C.c_double, vp, vp]),
    ("spb_check_traits", C.c_int, [vp, C.POINTER(Traits), vp, C.c_int, C.c_double, C.c_double, C.POINTER(CheckResult), vp]),
    ("spb_lm_solve", C.c_int, [vp, C.POINTER(Traits), C.POINTER(LMOptions), C.POINTER(LMResult), vp, C.c_int, vp,
                               C.c_int32, C.POINTER(C.c_int32)]),
    ("spb_problem_lf", C.c_int, [vp, C.c_int32, C.c_int32, C.c_uint64, C.POINTER(vp)]),
    ("spb_problem_lf_batch", C.c_int, [vp, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(vp)]),
    ("spb_lm_solve_batch", C.c_int, [vp, C.POINTER(Traits), C.c_int32, C.POINTER(LMOptions), C.POINTER(LMResult), vp, C.c_int]),
    ("spb_problem_rosenbrock", C.c_int, [vp, C.c_int32, vp, C.POINTER(vp)]),
    ("spb_problem_lf_cols", C.c_int, [vp, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(vp)]),
    ("spb_problem_blocks", C.c_int, [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_int32, vp, vp, C.c_double, C.c_double, vp, C.POINTER(vp)]),
    ("spb_problem_traits", C.c_int, [vp, C.POINTER(Traits)]),
    ("spb_problem_destroy", C.c_int, [vp]),
    ("spb_comm_unique_id", C.c_int, [vp]),
    ("spb_comm_init", C.c_int, [vp, vp, C.c_int, C.c_int]),
    ("spb_comm_destroy", C.c_int, [vp]),
    ("spb_dist_ring_halo", C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("spb_dist_peer_ready", C.c_int, [vp, C.POINTER(C.c_int)]),
    ("spb_analyze_partitioned", C.c_int, [vp, C.c_int32, C.c_int32, vp, vp, C.c_int32, C.c_int32]),
    ("spb_local_nnz", C.c_int, [vp, C.POINTER(C.c_int64)]),
    ("spb_local_pattern", C.c_int, [vp, vp, vp]),
    ("spb_collective_count", C.c_int64, [vp]),
    ("spb_problem_lf_rows", C.c_int, [vp, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(vp)]),
    ("spb_launch_count", C.c_int64, [vp]),
    ("spb_assembly_form", C.c_int, [vp]),
    ("spb_stage_ms", C.c_int, [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    ("spb_stage_reset", C.c_int, [vp, C.c_int]),
]

_LIB = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load the shared library and bind every declared symbol (raises if one is missing)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise FileNotFoundError(path)
        _build.build()
    L = C.CDLL(path)
    for name, res, args in SYMBOLS:
        f = getattr(L, name)  # AttributeError if the symbol is not exported
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


class SpbError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__("%s%s" % (STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status,
                                   (": " + msg) if msg else ""))
