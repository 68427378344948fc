"""ctypes view of oracle/libsp_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (saddlepoint_b200/) never does.  PARITY UNPINNED: see the
header of sp_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "libsp_oracle.so")
    src = os.path.join(_HERE, "sp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsp_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsp_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp = C.c_void_p
        L.spo_from_triplets.restype = vp
        L.spo_from_triplets.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp]
        L.spo_from_csc.restype = vp
        L.spo_from_csc.argtypes = [C.c_int, C.c_int, vp, vp, vp]
        L.spo_matrix_free.argtypes = [vp]
        for f in ("spo_matrix_rows", "spo_matrix_cols", "spo_matrix_nnz"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [vp]
        L.spo_matrix_get.argtypes = [vp, vp, vp, vp]
        L.spo_damp_build.restype = vp
        L.spo_damp_build.argtypes = [vp, C.c_double, vp]
        L.spo_ata.restype = vp
        L.spo_ata.argtypes = [vp]
        L.spo_jt_vec_neg.argtypes = [vp, vp, vp]
        L.spo_lf_matrix.restype = vp
        L.spo_lf_matrix.argtypes = [C.c_int, C.c_int, C.c_uint64]
        L.spo_splitmix64_at.restype = C.c_uint64
        L.spo_splitmix64_at.argtypes = [C.c_uint64, C.c_uint64]
        L.spo_amd.argtypes = [C.c_int, vp, vp, vp]
        L.spo_llt_compute.restype = vp
        L.spo_llt_compute.argtypes = [vp, C.c_int, vp]
        L.spo_llt_nnz.restype = C.c_longlong
        L.spo_llt_nnz.argtypes = [vp]
        L.spo_llt_perm.argtypes = [vp, vp]
        L.spo_llt_solve.argtypes = [vp, vp, vp]
        L.spo_llt_get.argtypes = [vp, vp, vp, vp]
        L.spo_llt_free.argtypes = [vp]
        L.spo_pcg_jacobi.restype = C.c_int
        L.spo_pcg_jacobi.argtypes = [vp, vp, vp, C.c_double, C.c_int, vp]
        L.spo_lm_solve.restype = vp
        L.spo_lm_solve.argtypes = [vp, vp]
        L.spo_lm_free.argtypes = [vp]
        for f in ("spo_lm_ret", "spo_lm_exit_path", "spo_lm_curr_iter", "spo_lm_factorizations",
                  "spo_lm_xsize", "spo_lm_trace_len"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [vp]
        for f in ("spo_lm_energy", "spo_lm_foo", "spo_lm_lambda"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [vp]
        L.spo_lm_get_x.argtypes = [vp, vp]
        L.spo_lm_get_trace.argtypes = [vp, C.c_int, vp]
        L.spo_lm_get_vec.argtypes = [vp, C.c_int, vp]
        L.spo_lm_first_h.restype = vp
        L.spo_lm_first_h.argtypes = [vp]
        L.spo_lm_time.restype = C.c_double
        L.spo_lm_time.argtypes = [vp, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Matrix:
    """Owning handle on an oracle CSC matrix (Eigen::SparseMatrix<double> stand-in)."""

    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().spo_matrix_free(self.h)
            self.h = None

    @property
    def shape(self):
        return lib().spo_matrix_rows(self.h), lib().spo_matrix_cols(self.h)

    @property
    def nnz(self):
        return lib().spo_matrix_nnz(self.h)

    def csc(self):
        rows, cols = self.shape
        nz = self.nnz
        p = np.empty(cols + 1, np.int32)
        i = np.empty(nz, np.int32)
        x = np.empty(nz, np.float64)
        lib().spo_matrix_get(self.h, _p(p), _p(i), _p(x))
        return p, i, x

    def scipy(self):
        import scipy.sparse as sp
        p, i, x = self.csc()
        return sp.csc_matrix((x, i, p), shape=self.shape)


def from_triplets(rows, cols, ti, tj, tv):
    ti = np.ascontiguousarray(ti, np.int32)
    tj = np.ascontiguousarray(tj, np.int32)
    tv = np.ascontiguousarray(tv, np.float64)
    return Matrix(lib().spo_from_triplets(rows, cols, len(ti), _p(ti), _p(tj), _p(tv)))


def from_csc(rows, cols, colptr, rowidx, vals):
    colptr = np.ascontiguousarray(colptr, np.int32)
    rowidx = np.ascontiguousarray(rowidx, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    return Matrix(lib().spo_from_csc(rows, cols, _p(colptr), _p(rowidx), _p(vals)))


def damp_build(J, lam):
    n = J.shape[1]
    d = np.empty(n, np.float64)
    return Matrix(lib().spo_damp_build(J.h, float(lam), _p(d))), d


def ata(A):
    return Matrix(lib().spo_ata(A.h))


def jt_vec_neg(J, e):
    e = np.ascontiguousarray(e, np.float64)
    out = np.empty(J.shape[1], np.float64)
    lib().spo_jt_vec_neg(J.h, _p(e), _p(out))
    return out


def lf_matrix(g, rows_mult=2, seed=20211):
    return Matrix(lib().spo_lf_matrix(g, rows_mult, seed))


def splitmix64_at(seed, ctr):
    return int(lib().spo_splitmix64_at(seed, ctr))


def amd(n, colptr, rowidx):
    colptr = np.ascontiguousarray(colptr, np.int32)
    rowidx = np.ascontiguousarray(rowidx, np.int32)
    perm = np.empty(n, np.int32)
    lib().spo_amd(n, _p(colptr), _p(rowidx), _p(perm))
    return perm


class LLT:
    """SimplicialLLT restatement: compute (ordering + symbolic + numeric) then solve."""

    def __init__(self, A, ordering=1):
        ok = C.c_int(0)
        self.h = lib().spo_llt_compute(A.h, ordering, C.byref(ok))
        self.ok = bool(ok.value)
        self.n = A.shape[1]

    def __del__(self):
        if getattr(self, "h", None):
            lib().spo_llt_free(self.h)
            self.h = None

    @property
    def nnzL(self):
        return int(lib().spo_llt_nnz(self.h))

    def perm(self):
        p = np.empty(self.n, np.int32)
        lib().spo_llt_perm(self.h, _p(p))
        return p

    def solve(self, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.empty(self.n, np.float64)
        lib().spo_llt_solve(self.h, _p(b), _p(x))
        return x

    def factor(self):
        import scipy.sparse as sp
        Lp = np.empty(self.n + 1, np.int32)
        lib().spo_llt_get(self.h, _p(Lp), None, None)
        Li = np.empty(Lp[-1], np.int32)
        Lx = np.empty(Lp[-1], np.float64)
        lib().spo_llt_get(self.h, None, _p(Li), _p(Lx))
        return sp.csc_matrix((Lx, Li, Lp), shape=(self.n, self.n))


def pcg_jacobi(H, b, rtol=1e-13, maxit=10000):
    b = np.ascontiguousarray(b, np.float64)
    x = np.empty(len(b), np.float64)
    rr = C.c_double(0)
    it = lib().spo_pcg_jacobi(H.h, _p(b), _p(x), rtol, maxit, C.byref(rr))
    return x, it, rr.value


_CB = C.CFUNCTYPE(None, C.c_void_p, c_dbl_p, c_dbl_p, C.c_int, c_dbl_p)


class _Problem(C.Structure):
    _fields_ = [("kind", C.c_int), ("iparam0", C.c_int), ("iparam1", C.c_int), ("seed", C.c_uint64),
                ("x0", C.c_void_p), ("m", C.c_int), ("n", C.c_int), ("colptr", C.c_void_p),
                ("rowidx", C.c_void_p), ("cb", _CB), ("user", C.c_void_p)]


class _Options(C.Structure):
    _fields_ = [("maxIterations", C.c_int), ("funcTolerance", C.c_double), ("fooTolerance", C.c_double),
                ("lambda0", C.c_double), ("verbose", C.c_int), ("solver_kind", C.c_int),
                ("pcg_rtol", C.c_double), ("pcg_maxit", C.c_int), ("keep_first", C.c_int)]


class LMResult:
    EXIT = {1: "foo", 2: "factorize_failed", 3: "direction_small", 4: "func_tol", 5: "traits_stop", 6: "max_iter"}

    def __init__(self, h):
        L = lib()
        self.ret = bool(L.spo_lm_ret(h))
        self.exit_path = self.EXIT[L.spo_lm_exit_path(h)]
        self.currIter = L.spo_lm_curr_iter(h)
        self.factorizations = L.spo_lm_factorizations(h)
        self.energy = L.spo_lm_energy(h)
        self.fooOptimality = L.spo_lm_foo(h)
        self.currLambda = L.spo_lm_lambda(h)
        n = L.spo_lm_xsize(h)
        self.x = np.empty(n, np.float64)
        L.spo_lm_get_x(h, _p(self.x))
        T = L.spo_lm_trace_len(h)
        names = ["prevEnergy2", "newEnergy2", "lambda", "foo", "dirnorm", "accepted", "lin_iters"]
        self.trace = {}
        for k, nm in enumerate(names):
            a = np.empty(T, np.float64)
            L.spo_lm_get_trace(h, k, _p(a))
            self.trace[nm] = a
        self.first_direction = self.first_rhs = self.firstH = None
        if self.factorizations > 0:
            self.last_direction = np.empty(n, np.float64)
            L.spo_lm_get_vec(h, 2, _p(self.last_direction))
        self.times = {nm: L.spo_lm_time(h, k) for k, nm in enumerate(["assembly", "factor", "solve", "traits"])}
        self._h = h

    def load_first(self):
        L = lib()
        n = len(self.x)
        self.first_direction = np.empty(n, np.float64)
        L.spo_lm_get_vec(self._h, 0, _p(self.first_direction))
        self.first_rhs = np.empty(n, np.float64)
        L.spo_lm_get_vec(self._h, 1, _p(self.first_rhs))
        self.firstH = Matrix(L.spo_lm_first_h(self._h))
        return self

    def __del__(self):
        if getattr(self, "_h", None):
            lib().spo_lm_free(self._h)
            self._h = None


def set_threads(t):
    """Host threads of the solver_kind=3 baseline variant (threaded Jacobi-PCG + product); parity tests never use it."""
    lib().spo_set_threads(int(t))


def lm_solve(kind, iparam0=0, iparam1=0, seed=20211, x0=None, pattern=None, callback=None,
             maxIterations=100, funcTolerance=10e-10, fooTolerance=10e-10, lambda0=0.01,
             verbose=True, solver_kind=0, pcg_rtol=1e-13, pcg_maxit=10000, keep_first=False):
    """Run the restated LMSolver::solve.  kind: 'rosenbrock' | 'linear_dense' | 'lf' | 'callback'.

    callback(x, compute_jacobian) -> (E, Jvals or None) with Jvals in the CSC order of
    pattern=(m, n, colptr, rowidx).
    """
    kinds = {"rosenbrock": 0, "linear_dense": 1, "lf": 2, "callback": 3}
    P = _Problem()
    P.kind = kinds[kind]
    P.iparam0, P.iparam1, P.seed = iparam0, iparam1, seed
    keep = []
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, np.float64)
        keep.append(x0)
        P.x0 = x0.ctypes.data
    if kind == "callback":
        m, n, colptr, rowidx = pattern
        colptr = np.ascontiguousarray(colptr, np.int32)
        rowidx = np.ascontiguousarray(rowidx, np.int32)
        keep += [colptr, rowidx]
        P.m, P.n = m, n
        P.colptr, P.rowidx = colptr.ctypes.data, rowidx.ctypes.data
        nnz = int(colptr[-1])

        def _cb(user, xp, Ep, cj, Jp):
            x = np.ctypeslib.as_array(xp, shape=(n,))
            E, Jv = callback(x.copy(), bool(cj))
            np.ctypeslib.as_array(Ep, shape=(m,))[:] = E
            if cj:
                np.ctypeslib.as_array(Jp, shape=(nnz,))[:] = Jv

        P.cb = _CB(_cb)
        keep.append(P.cb)
    O = _Options(maxIterations, funcTolerance, fooTolerance, lambda0, int(verbose), solver_kind,
                 pcg_rtol, pcg_maxit, int(keep_first))
    h = lib().spo_lm_solve(C.byref(P), C.byref(O))
    r = LMResult(h)
    if keep_first and r.factorizations > 0:
        r.load_first()
    return r
