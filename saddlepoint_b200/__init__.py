"""saddlepoint_b200 -- B200-native Levenberg-Marquardt inner step behind SaddlePoint's concept API.

The numeric work lives in libsaddlepoint_b200.so (CUDA, sm_100a); see include/saddlepoint_b200.h.
"""
from . import capi
from .capi import SOLVER_CHOLESKY, SOLVER_PCG_IC0, SOLVER_PCG_JACOBI, SpbError
from .solver import B200SolverWrapper, BuiltinProblem, DiagonalDamping, Handle, LMSolver, check_traits

__all__ = ["capi", "Handle", "B200SolverWrapper", "DiagonalDamping", "LMSolver", "BuiltinProblem", "SpbError", "check_traits",
           "SOLVER_CHOLESKY", "SOLVER_PCG_JACOBI", "SOLVER_PCG_IC0"]
