"""Host-side mirror of SaddlePoint's concept API over the C ABI.

Classes keep the reference's names and argument meaning so parity tests read like the
reference's own mains (benchmark/rosenbrock/main.cpp:71-87):

    LinearSolver  -> B200SolverWrapper      (EigenSolverWrapper.h:23-79: analyze/factorize/solve)
    DampingTraits -> DiagonalDamping        (DiagonalDamping.h:20-96: currLambda, init, update)
    LMSolver      -> LMSolver               (LMSolver.h:22-193: init(LS,ST,DT,...), solve(verbose))

Everything numeric happens in libsaddlepoint_b200.so on the GPU; this module only moves
pointers.  Arrays may be numpy (host) or torch CUDA tensors (device, zero-copy).
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import SPB_DEVICE, SPB_HOST, SPB_NOT_CONVERGED, SPB_NOT_SPD, SPB_OK, SpbError

EXIT_PATHS = {1: "foo", 2: "factorize_failed", 3: "direction_small", 4: "func_tol", 5: "traits_stop", 6: "max_iter"}


def _is_torch_cuda(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _ptr(a):
    """(pointer, mem, keepalive) for a float64 numpy array or torch CUDA tensor."""
    if a is None:
        return None, SPB_HOST, None
    if _is_torch_cuda(a):
        import torch
        if a.dtype != torch.float64 or not a.is_contiguous():
            raise TypeError("device arrays must be contiguous float64 tensors")
        return C.c_void_p(a.data_ptr()), SPB_DEVICE, a
    a = np.ascontiguousarray(a, np.float64)
    return C.c_void_p(a.ctypes.data), SPB_HOST, a


def _i32(a):
    a = np.ascontiguousarray(a, np.int32)
    return a, C.c_void_p(a.ctypes.data)


class Handle:
    """One spb_handle: a device, a stream, the analysed pattern and the device-resident H."""

    def __init__(self, device=0, stream=None):
        self.L = capi.load()
        h = C.c_void_p()
        st = self.L.spb_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if st != SPB_OK:
            raise SpbError(st, "spb_create failed: a CUDA device is required (no CPU fallback)")
        self.h = h
        self.device = device
        self.m = self.n = 0
        self.nnzJ = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.spb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, allow=()):
        if st != SPB_OK and st not in allow:
            raise SpbError(st, self.L.spb_last_error(self.h).decode())
        return st

    # ---- symbolic
    def analyze(self, m, n, colptr, rowidx):
        colptr, pc = _i32(colptr)
        rowidx, pr = _i32(rowidx)
        self._check(self.L.spb_analyze(self.h, m, n, pc, pr))
        self.m, self.n, self.nnzJ = m, n, int(colptr[-1])

    def analyze_update(self, m, n, colptr, rowidx):
        """spb_analyze_update: 0 = unchanged, 1 = patched (one appended single-entry row), 2 = full analysis."""
        colptr, pc = _i32(colptr)
        rowidx, pr = _i32(rowidx)
        how = C.c_int(2)
        self._check(self.L.spb_analyze_update(self.h, m, n, pc, pr, C.byref(how)))
        self.m, self.n, self.nnzJ = m, n, int(colptr[-1])
        return how.value

    # ---- multi-GPU (one big problem, rows partitioned across ranks)
    def comm_init(self, unique_id, rank, nranks):
        """Attach an NCCL communicator (unique_id: 128 bytes from comm_unique_id() on one rank)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.L.spb_comm_init(self.h, C.cast(buf, C.c_void_p), rank, nranks))
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def comm_unique_id():
        L = capi.load()
        buf = (C.c_char * 128)()
        st = L.spb_comm_unique_id(C.cast(buf, C.c_void_p))
        if st != SPB_OK:
            raise SpbError(st, "ncclGetUniqueId failed (is libnccl available?)")
        return bytes(buf)

    def analyze_partitioned(self, m, n, colptr, rowidx, row_begin, row_end):
        colptr, pc = _i32(colptr)
        rowidx, pr = _i32(rowidx)
        self._check(self.L.spb_analyze_partitioned(self.h, m, n, pc, pr, row_begin, row_end))
        self.m, self.n = row_end - row_begin, n
        nnz = C.c_int64(0)
        self._check(self.L.spb_local_nnz(self.h, C.byref(nnz)))
        self.nnzJ = nnz.value

    def ring_halo(self, own_begin, own_end, halo, left_rank, right_rank):
        """Subdomain (halo) mode (spb_dist_ring_halo): this handle holds the local problem of its rank."""
        self._check(self.L.spb_dist_ring_halo(self.h, own_begin, own_end, halo, left_rank, right_rank))

    def peer_exchange_ready(self):
        """True if the halo mode exchanges through NVLink peer-memory windows inside the PCG kernel (spb_dist_peer_ready)."""
        r = C.c_int(0)
        self._check(self.L.spb_dist_peer_ready(self.h, C.byref(r)))
        return bool(r.value)

    def local_pattern(self):
        colptr = np.empty(self.n + 1, np.int32)
        rowidx = np.empty(self.nnzJ, np.int32)
        self._check(self.L.spb_local_pattern(self.h, C.c_void_p(colptr.ctypes.data), C.c_void_p(rowidx.ctypes.data)))
        return colptr, rowidx

    def collective_count(self):
        return int(self.L.spb_collective_count(self.h))

    def pattern_matches(self, m, n, colptr, rowidx):
        colptr, pc = _i32(colptr)
        rowidx, pr = _i32(rowidx)
        same = C.c_int(0)
        self._check(self.L.spb_pattern_matches(self.h, m, n, pc, pr, C.byref(same)))
        return bool(same.value)

    def h_nnz(self):
        nnz = C.c_int64(0)
        self._check(self.L.spb_h_nnz(self.h, C.byref(nnz)))
        return nnz.value

    def h_pattern(self):
        nnz = self.h_nnz()
        colptr = np.empty(self.n + 1, np.int32)
        rowidx = np.empty(nnz, np.int32)
        self._check(self.L.spb_h_pattern(self.h, C.c_void_p(colptr.ctypes.data), C.c_void_p(rowidx.ctypes.data)))
        return colptr, rowidx

    def h_values(self):
        vals = np.empty(self.h_nnz(), np.float64)
        self._check(self.L.spb_h_values(self.h, C.c_void_p(vals.ctypes.data), SPB_HOST))
        return vals

    def h_scipy(self):
        import scipy.sparse as sp
        colptr, rowidx = self.h_pattern()
        return sp.csc_matrix((self.h_values(), rowidx, colptr), shape=(self.n, self.n))

    # ---- numeric
    def assemble(self, Jvals, EVec, lam):
        pj, mj, kj = _ptr(Jvals)
        pe, me, ke = _ptr(EVec)
        if EVec is not None and me != mj:
            raise TypeError("Jvals and EVec must live in the same memory space")
        foo = C.c_double(0)
        self._check(self.L.spb_assemble(self.h, pj, pe, float(lam), mj, C.byref(foo)))
        return foo.value

    def rhs(self):
        out = np.empty(self.n, np.float64)
        self._check(self.L.spb_get_rhs(self.h, C.c_void_p(out.ctypes.data), SPB_HOST))
        return out

    def damping(self):
        out = np.empty(self.n, np.float64)
        self._check(self.L.spb_get_damping(self.h, C.c_void_p(out.ctypes.data), SPB_HOST))
        return out

    def set_solver(self, kind, pcg_rtol=0.0, pcg_maxit=0):
        self._check(self.L.spb_set_solver(self.h, kind, pcg_rtol, pcg_maxit))

    def set_matrix(self, n, colptr, rowidx, vals):
        colptr, pc = _i32(colptr)
        rowidx, pr = _i32(rowidx)
        pv, mv, kv = _ptr(vals)
        self._check(self.L.spb_set_matrix(self.h, n, pc, pr, pv, mv))
        self.n = n

    def analyze_triplets(self, n, rows, cols, symmetric=False):
        """Legacy LS->analyze(HRows, HCols, symmetric) (GNSolver.h:131, SLSolver.h:123)."""
        rows, pr = _i32(rows)
        cols, pc = _i32(cols)
        if rows.shape != cols.shape:
            raise ValueError("rows and cols must have the same length")
        self._check(self.L.spb_analyze_triplets(self.h, n, rows.size, pr, pc, int(bool(symmetric))))
        self.n = n
        self._trip_nnz = rows.size

    def factorize_triplets(self, vals):
        """Legacy LS->factorize(HVals, symmetric) (GNSolver.h:169, SLSolver.h:159): one value per triplet."""
        pv, mv, kv = _ptr(vals)
        if len(vals) != getattr(self, "_trip_nnz", -1):
            raise ValueError("one value per analysed triplet is required")
        st = self._check(self.L.spb_factorize_triplets(self.h, pv, mv), allow=(SPB_NOT_SPD,))
        return st == SPB_OK

    def factorize(self):
        """True on success, False iff the matrix is not positive definite (the reference's bool)."""
        st = self._check(self.L.spb_factorize(self.h), allow=(SPB_NOT_SPD,))
        return st == SPB_OK

    def solve(self, rhs=None, out=None, nrhs=1):
        """Returns (x, iters, status).  rhs=None solves for the assembled gradient."""
        pr, mr, kr = _ptr(rhs)
        if out is None:
            if rhs is not None and mr == SPB_DEVICE:
                import torch
                out = torch.empty_like(rhs)
            else:
                out = np.empty(self.n * nrhs, np.float64)
        po, mo, ko = _ptr(out)
        if rhs is not None and mo != mr:
            raise TypeError("rhs and out must live in the same memory space")
        it = C.c_int(0)
        st = self._check(self.L.spb_solve(self.h, pr, po, nrhs, mo, C.byref(it)), allow=(SPB_NOT_SPD, SPB_NOT_CONVERGED))
        return out, it.value, st

    def residual_check(self, x):
        """||H x - rhs|| / ||rhs|| for the assembled system (global over ranks in halo mode)."""
        p, m, k = _ptr(x)
        out = C.c_double(0)
        self._check(self.L.spb_residual_check(self.h, p, m, C.byref(out)))
        return out.value

    def factor_info(self):
        nnzL, flops, ns, nl = C.c_int64(0), C.c_double(0), C.c_int32(0), C.c_int32(0)
        self._check(self.L.spb_factor_info(self.h, C.byref(nnzL), C.byref(flops), C.byref(ns), C.byref(nl)))
        return {"nnzL": nnzL.value, "flops": flops.value, "nsuper": ns.value, "nlevels": nl.value}

    def squared_norm(self, v):
        p, m, k = _ptr(v)
        out = C.c_double(0)
        self._check(self.L.spb_squared_norm(self.h, p, int(v.numel() if m == SPB_DEVICE else len(k)), m, C.byref(out)))
        return out.value

    def inf_norm(self, v):
        p, m, k = _ptr(v)
        out = C.c_double(0)
        self._check(self.L.spb_inf_norm(self.h, p, int(v.numel() if m == SPB_DEVICE else len(k)), m, C.byref(out)))
        return out.value

    def spmv(self, x, out=None):
        px, mx, kx = _ptr(x)
        if out is None:
            if mx == SPB_DEVICE:
                import torch
                out = torch.empty_like(x)
            else:
                out = np.empty(self.n, np.float64)
        po, mo, ko = _ptr(out)
        self._check(self.L.spb_spmv(self.h, px, po, mx))
        return out

    # ---- instrumentation
    def launch_count(self):
        return int(self.L.spb_launch_count(self.h))

    def assembly_form(self):
        """3 stream form, 2 register-pipelined general kernel, 1 one-CTA-per-chunk general kernel (last assemble)."""
        return int(self.L.spb_assembly_form(self.h))

    def stage_reset(self, enable_timing=False):
        self._check(self.L.spb_stage_reset(self.h, int(enable_timing)))

    def stage_ms(self):
        names = ["assembly", "spmv", "pcg_vector", "factor", "trisolve", "reductions", "traits", "pcg_fused"]
        out = {}
        for k, nm in enumerate(names):
            ms = C.c_double(0)
            cnt = C.c_int64(0)
            self._check(self.L.spb_stage_ms(self.h, k, C.byref(ms), C.byref(cnt)))
            out[nm] = (ms.value, cnt.value)
        return out


def _as_csc(J):
    """(rows, cols, colptr, rowidx, vals) from a scipy sparse matrix or an explicit tuple."""
    if isinstance(J, tuple):
        return J
    J = J.tocsc()
    if not J.has_sorted_indices:
        J = J.sorted_indices()
    return J.shape[0], J.shape[1], J.indptr.astype(np.int32), J.indices.astype(np.int32), J.data.astype(np.float64)


class B200SolverWrapper:
    """GPU LinearSolver concept (replaces EigenSolverWrapper<SimplicialLLT>, EigenSolverWrapper.h:23-79)."""

    def __init__(self, handle=None, solver=capi.SOLVER_CHOLESKY, pcg_rtol=1e-13, pcg_maxit=10000, device=0):
        self.handle = handle or Handle(device)
        self.handle.set_solver(solver, pcg_rtol, pcg_maxit)
        self.last_iters = 0

    def analyze(self, JPattern, HCols=None, symmetric=False, n=None):
        """analyze(const SparseMatrix<int>& JPattern) (EigenSolverWrapper.h:30): J's pattern.

        Legacy overload analyze(HRows, HCols, symmetric) (GNSolver.h:131, SLSolver.h:123): the matrix
        itself as triplets; n defaults to max index + 1."""
        if HCols is not None:
            rows = np.asarray(JPattern, np.int32)
            cols = np.asarray(HCols, np.int32)
            if n is None:
                n = int(max(rows.max(initial=-1), cols.max(initial=-1))) + 1
            self.handle.analyze_triplets(n, rows, cols, symmetric)
            return True
        m, n, colptr, rowidx, _ = _as_csc(JPattern)
        self.handle.analyze(m, n, colptr, rowidx)
        return True

    analyze_pattern = analyze

    def factorize(self, A, symmetric=None):
        """factorize(SparseMatrix<double> A) (EigenSolverWrapper.h:54): explicit symmetric matrix.

        Legacy overload factorize(HVals, symmetric) (GNSolver.h:169, SLSolver.h:159): a 1-D value
        array, one per triplet given to analyze(HRows, HCols, symmetric)."""
        if not hasattr(A, "tocsc") and np.ndim(A) == 1:
            return self.handle.factorize_triplets(A if _is_torch_cuda(A) else np.ascontiguousarray(A, np.float64))
        n, _, colptr, rowidx, vals = _as_csc(A)
        self.handle.set_matrix(n, colptr, rowidx, vals)
        return self.handle.factorize()

    def solve(self, rhs):
        """solve(rhs, x) (EigenSolverWrapper.h:62-78); returns x (matrix RHS: columns)."""
        rhs = np.asarray(rhs, np.float64)
        if rhs.ndim == 2:
            flat = np.ascontiguousarray(rhs.T).ravel()
            x, it, st = self.handle.solve(flat, nrhs=rhs.shape[1])
            self.last_iters = it
            return x.reshape(rhs.shape[1], -1).T.copy()
        x, it, st = self.handle.solve(rhs)
        self.last_iters = it
        return x


class DiagonalDamping:
    """DampingTraits concept (DiagonalDamping.h): carries currLambda; the augmented matrix
    dampJ of the reference is never materialised -- d and sqrt(d)^2 are produced inside the
    fused assembly kernel."""

    def __init__(self, currLambda=0.01):
        self.currLambda = currLambda


def _c_traits(ST):
    """spb_traits for a Python SolverTraits object (host callbacks) or a built-in device problem.
    Returns (traits, objects that must stay alive while the traits are in use)."""
    keep = []
    T = capi.Traits()
    if hasattr(ST, "c_traits"):  # built-in device-resident problem
        return ST.c_traits(), keep
    x0 = np.ascontiguousarray(ST.initial_solution(), np.float64)
    n, m = int(ST.xSize), int(ST.ESize)
    if hasattr(ST, "colptr") and hasattr(ST, "rowidx"):  # traits that publish their fixed CSC pattern
        colptr, rowidx = ST.colptr, ST.rowidx
    else:
        E0, J0 = ST.objective_jacobian(x0.copy(), True)
        rows, cols, colptr, rowidx, _ = _as_csc(J0)
    colptr, pc = _i32(colptr)
    rowidx, pr = _i32(rowidx)
    nnz = int(colptr[-1])
    keep += [colptr, rowidx]

    def _init(user, xp):
        np.ctypeslib.as_array(xp, shape=(n,))[:] = x0

    def _objjac(user, xp, Ep, Jp, stream):
        x = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), shape=(n,)).copy()
        E, J = ST.objective_jacobian(x, bool(Jp))
        np.ctypeslib.as_array(C.cast(Ep, C.POINTER(C.c_double)), shape=(m,))[:] = E
        if Jp:
            vals = _as_csc(J)[4] if not isinstance(J, np.ndarray) else J
            np.ctypeslib.as_array(C.cast(Jp, C.POINTER(C.c_double)), shape=(nnz,))[:] = vals

    def _pre(user, xp):
        if hasattr(ST, "pre_iteration"):
            ST.pre_iteration(np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), shape=(n,)).copy())

    def _post(user, xp):
        if hasattr(ST, "post_iteration"):
            return int(bool(ST.post_iteration(np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), shape=(n,)).copy())))
        return 0

    T.xSize, T.ESize = n, m
    T.colptr, T.rowidx = pc, pr
    T.pattern_version = 0
    T.mem = SPB_HOST
    T.initial_solution = capi.INIT_CB(_init)
    T.objective_jacobian = capi.OBJJAC_CB(_objjac)
    T.pre_iteration = capi.PRE_CB(_pre)
    T.post_iteration = capi.POST_CB(_post)
    keep += [T.initial_solution, T.objective_jacobian, T.pre_iteration, T.post_iteration]
    return T, keep


def check_traits(handle, ST, x, step=0.0, tol=0.0, return_fe=False, verbose=True):
    """check_traits(Traits, CurrSolution) (check_traits.h:19-83) with column-coloured central
    differences on the GPU library: 2*ncolors evaluations instead of 2*xSize.  Returns a dict
    (max_diff, max_row, max_col, discrepancies, outside_max, outside_count, ncolors, evaluations
    [, fe: finite-difference values in the Jacobian's CSC order])."""
    T, keep = _c_traits(ST)
    px, mx, kx = _ptr(x)
    nnz = int(np.ctypeslib.as_array(C.cast(T.colptr, C.POINTER(C.c_int32)), shape=(T.xSize + 1,))[-1])
    fe = np.empty(nnz, np.float64) if return_fe else None
    R = capi.CheckResult()
    if verbose:
        print("WARNING: FE gradient checking, reducing performance!!!")
        print("******************************************************")
        print("Solution Size:", T.xSize)
    st = handle.L.spb_check_traits(handle.h, C.byref(T), px, mx, step, tol, C.byref(R), C.c_void_p(fe.ctypes.data) if return_fe else None)
    handle._check(st)
    handle.m, handle.n = T.ESize, T.xSize
    out = {k: getattr(R, k) for k, _ in capi.CheckResult._fields_}
    if verbose:
        print("Maximum gradient difference:", R.max_diff)
        print("At Location: (%d,%d)" % (R.max_row, R.max_col))
    if return_fe:
        out["fe"] = fe
    return out


class LMSolver:
    """LMSolver<LinearSolver, SolverTraits, DampingTraits> (LMSolver.h:22-193).

    SolverTraits in Python: attributes xSize, ESize; methods initial_solution() -> x0,
    objective_jacobian(x, computeJacobian) -> (EVec, J or None) with J a scipy sparse matrix of
    fixed pattern (or just its CSC value array once the pattern is known via jacobian_pattern()),
    optional pre_iteration(x), post_iteration(x) -> bool.
    A device-resident built-in problem (saddlepoint_b200.problems) can be passed instead.
    """

    def __init__(self):
        self.x = None
        self.energy = 0.0
        self.fooOptimality = 0.0
        self.currIter = 0
        self.trace = None
        self.result = None

    def init(self, LS, ST, DT, maxIterations=100, funcTolerance=10e-10, fooTolerance=10e-10):
        self.LS, self.ST, self.DT = LS, ST, DT
        self.maxIterations = maxIterations
        self.funcTolerance = funcTolerance
        self.fooTolerance = fooTolerance

    def solve(self, verbose, dedupe_evaluations=False, fixed_iterations=0, x_out=None):
        H = self.LS.handle
        L = H.L
        T, keep = _c_traits(self.ST)
        O = capi.LMOptions()
        L.spb_lm_default_options(C.byref(O))
        O.maxIterations = self.maxIterations
        O.funcTolerance = self.funcTolerance
        O.fooTolerance = self.fooTolerance
        O.lambda0 = self.DT.currLambda
        O.verbose = int(bool(verbose))
        O.dedupe_evaluations = int(bool(dedupe_evaluations))
        O.fixed_iterations = int(fixed_iterations)
        R = capi.LMResult()
        cap = (fixed_iterations if fixed_iterations > 0 else self.maxIterations) + 2
        trace = np.zeros((cap, 8), np.float64)
        tlen = C.c_int32(0)
        if x_out is None:
            x_out = np.empty(T.xSize, np.float64)
        px, mx, kx = _ptr(x_out)
        st = L.spb_lm_solve(H.h, C.byref(T), C.byref(O), C.byref(R), px, mx, C.c_void_p(trace.ctypes.data), cap, C.byref(tlen))
        H._check(st)
        H.m, H.n = T.ESize, T.xSize
        self.x = x_out
        self.energy = R.energy
        self.fooOptimality = R.fooOptimality
        self.currIter = R.currIter
        self.DT.currLambda = R.lambda_
        self.factorizations = R.factorizations
        self.linear_iterations = R.linear_iterations
        self.exit_path = EXIT_PATHS.get(R.exit_path, R.exit_path)
        self.seconds_total = R.seconds_total
        self.seconds_analyze = R.seconds_analyze
        self.unconverged_solves = R.unconverged_solves
        names = ["prevEnergy2", "newEnergy2", "lambda", "foo", "dirnorm", "accepted", "lin_iters"]
        self.trace = {nm: trace[: tlen.value, k].copy() for k, nm in enumerate(names)}
        return bool(R.ret)


class BuiltinProblem:
    """Device-resident SolverTraits (spb_problem_*): the benchmark workloads evaluated on the GPU.
    kinds: lf | lf_rows | rosenbrock | blocks (A, b, bar_idx, bar_vol, s, w_bar, x0: see spb_problem_blocks)."""

    def __init__(self, handle, kind, **kw):
        self.handle = handle
        L = handle.L
        p = C.c_void_p()
        if kind == "lf":
            st = L.spb_problem_lf(handle.h, kw["g"], kw.get("rows_mult", 2), kw.get("seed", 20211), C.byref(p))
        elif kind == "lf_batch":  # nproblems LF problems (seeds seed, seed+1, ...) as one block-diagonal problem
            st = L.spb_problem_lf_batch(handle.h, kw["g"], kw.get("rows_mult", 2), kw.get("seed", 20211), kw.get("seed_stride", 1), kw["nproblems"], C.byref(p))
        elif kind == "lf_rows":  # row-partitioned LF: this rank owns residual rows [row_begin,row_end)
            st = L.spb_problem_lf_rows(handle.h, kw["g"], kw.get("rows_mult", 2), kw.get("seed", 20211), kw["row_begin"], kw["row_end"],
                                       C.byref(p))
        elif kind == "lf_cols":  # local part of LF for the subdomain (halo) mode: owned unknowns [col_begin,col_end)
            st = L.spb_problem_lf_cols(handle.h, kw["g"], kw.get("rows_mult", 2), kw.get("seed", 20211), kw["col_begin"], kw["col_end"],
                                       C.byref(p))
        elif kind == "rosenbrock":
            x0 = kw.get("x0")
            if x0 is not None:
                x0 = np.ascontiguousarray(x0, np.float64)
            st = L.spb_problem_rosenbrock(handle.h, kw["nblocks"], C.c_void_p(x0.ctypes.data) if x0 is not None else None, C.byref(p))
        elif kind == "blocks":  # constant sparse block + barrier block (seamless-integration structure)
            A = kw["A"].tocsc()
            A.sort_indices()
            n, m_lin = A.shape[1], A.shape[0]
            Ap, pAp = _i32(A.indptr)
            Ai, pAi = _i32(A.indices)
            Av = np.ascontiguousarray(A.data, np.float64)
            b = np.ascontiguousarray(kw["b"], np.float64)
            bidx, pbi = _i32(np.asarray(kw["bar_idx"]).reshape(-1))
            bvol = np.ascontiguousarray(kw["bar_vol"], np.float64).reshape(-1)
            x0 = np.ascontiguousarray(kw["x0"], np.float64)
            if b.size != m_lin or x0.size != n or bidx.size != 4 * bvol.size:
                raise ValueError("blocks problem: inconsistent sizes")
            st = L.spb_problem_blocks(handle.h, n, m_lin, pAp, pAi, C.c_void_p(Av.ctypes.data), C.c_void_p(b.ctypes.data), bvol.size, pbi,
                                      C.c_void_p(bvol.ctypes.data), float(kw["s"]), float(kw["w_bar"]), C.c_void_p(x0.ctypes.data), C.byref(p))
        else:
            raise ValueError(kind)
        handle._check(st)
        self.p = p
        T = capi.Traits()
        handle._check(L.spb_problem_traits(p, C.byref(T)))
        self._T = T
        self.xSize, self.ESize = T.xSize, T.ESize

    def c_traits(self):
        return self._T

    def pattern(self):
        n = self.xSize
        colptr = np.ctypeslib.as_array(C.cast(self._T.colptr, C.POINTER(C.c_int32)), shape=(n + 1,)).copy()
        rowidx = np.ctypeslib.as_array(C.cast(self._T.rowidx, C.POINTER(C.c_int32)), shape=(int(colptr[-1]),)).copy()
        return self.ESize, n, colptr, rowidx

    def close(self):
        if getattr(self, "p", None):
            self.handle.L.spb_problem_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
