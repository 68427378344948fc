"""Vectorised builder of the config-C3 stand-in in the form spb_problem_blocks consumes (test / bench
infrastructure): E = [A x - b ; w_bar * barrier(x)] with the Jacobian block structure and weights of
SIInitialSolutionTraits.h:115-248,255-259 on a regular grid triangulation (see tests/c3_problem.py for
the per-face reference construction; this one scales to the 500 000-face size of SURVEY 8d).

`BlocksHostTraits` evaluates the same problem in numpy with exactly the operation order of the device
kernels (problems.cu: blocks_linear_kernel / blocks_barrier_kernel), so host and device produce
identical bits and the oracle can be driven by it."""
import numpy as np
import scipy.sparse as sp

N = 4
W_INT, W_CONST, W_BARR, W_CLOSE, S_BAR = 10e3, 10e3, 0.0001, 1.0, 0.1


def build(nx, ny, seed=20212, flip_face=None):
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny))
    V = np.stack([ii.ravel(), jj.ravel()], 1).astype(float)          # vertex j*nx+i at (i, j)
    ci, cj = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1))
    a = (cj * nx + ci).ravel()
    b_, c_, d_ = a + 1, a + nx, a + nx + 1
    F = np.empty((2 * a.size, 3), np.int64)
    F[0::2] = np.stack([a, b_, d_], 1)
    F[1::2] = np.stack([a, d_, c_], 1)
    nV, nF = len(V), len(F)
    nfun = N // 2
    nx_vars, nfield = nfun * nV, 2 * N * nF
    n = nx_vars + nfield
    # per-face gradient operator: grad = M @ [ub-ua, uc-ua]
    E1, E2 = V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]
    M = np.linalg.inv(np.stack([E1, E2], 1))                          # (nF, 2, 2)
    rows, cols, vals = [], [], []
    f = np.arange(nF)
    for j in range(N):
        fun, sign = j % nfun, (1.0 if j < nfun else -1.0)
        for comp in range(2):
            r = 2 * N * f + 2 * j + comp
            for vtx, w in ((F[:, 0], -M[:, comp, 0] - M[:, comp, 1]), (F[:, 1], M[:, comp, 0]), (F[:, 2], M[:, comp, 1])):
                rows.append(r); cols.append(fun * nV + vtx); vals.append(sign * w)
    G2 = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nfield, nx_vars))
    L = 1.0
    xs, ys = V[:, 0], V[:, 1]
    u0 = xs + 0.15 * np.sin(0.7 * ys) + 0.02 * rng.standard_normal(nV)
    v0 = ys + 0.15 * np.sin(0.6 * xs) + 0.02 * rng.standard_normal(nV)
    x_init = np.concatenate([u0, v0])
    field0 = L * (G2 @ x_init) + 0.01 * rng.standard_normal(nfield)
    if flip_face is not None:
        k = 2 * N * flip_face
        field0[k:k + 2] = -field0[k:k + 2]
    fld = field0.reshape(nF, N, 2)
    nxt = np.roll(fld, -1, axis=1)
    vol = np.abs(fld[:, :, 0] * nxt[:, :, 1] - fld[:, :, 1] * nxt[:, :, 0]) + 1e-3   # (nF, N)
    fixed = np.array([0, nV])
    fixed_vals = x_init[fixed].copy()
    I = sp.identity(nfield, format="csr")
    gInt = sp.hstack([-L * G2, I]) * W_INT
    gClose = sp.hstack([sp.csr_matrix((nfield, nx_vars)), I]) * W_CLOSE
    gConst = sp.csr_matrix((np.ones(len(fixed)), (np.arange(len(fixed)), fixed)), shape=(len(fixed), n)) * W_CONST
    A = sp.vstack([gInt, gClose, gConst]).tocsr()
    A.sort_indices()
    b = np.concatenate([np.zeros(nfield), field0 * W_CLOSE, fixed_vals * W_CONST])
    jn = (np.arange(N) + 1) % N
    base = nx_vars + 2 * N * np.arange(nF)[:, None]
    bar_idx = np.stack([base + 2 * np.arange(N), base + 2 * np.arange(N) + 1, base + 2 * jn, base + 2 * jn + 1], 2).reshape(-1, 4)
    return {"A": A, "b": b, "bar_idx": bar_idx.astype(np.int32), "bar_vol": vol.reshape(-1), "s": S_BAR, "w_bar": W_BARR,
            "x0": np.concatenate([x_init, field0]), "nF": nF, "nV": nV}


class BlocksHostTraits:
    """numpy SolverTraits of the same problem, same bits as the device traits."""

    def __init__(self, P):
        self.P = P
        A = P["A"]
        self.A = A
        self.m_lin, self.xSize = A.shape
        self.nbar = len(P["bar_vol"])
        self.ESize = self.m_lin + self.nbar
        Ac = A.tocsc()
        Ac.sort_indices()
        n = self.xSize
        extra = np.bincount(P["bar_idx"].ravel(), minlength=n)
        self.colptr = np.concatenate([[0], np.cumsum(np.diff(Ac.indptr) + extra)]).astype(np.int32)
        # combined CSC: per column the constant rows, then the barrier rows (ascending)
        brow = np.repeat(np.arange(self.nbar), 4) + self.m_lin
        bcol = P["bar_idx"].ravel()
        acol = np.repeat(np.arange(n), np.diff(Ac.indptr))
        allr = np.concatenate([Ac.indices, brow])
        allc = np.concatenate([acol, bcol])
        order = np.lexsort((allr, allc))
        self.rowidx = allr[order].astype(np.int32)
        self._order = order
        self._avals = Ac.data
        self.calls = 0

    def initial_solution(self):
        return self.P["x0"].copy()

    def objective_jacobian(self, x, compute_jacobian):
        self.calls += 1
        P = self.P
        s, w, vol = P["s"], P["w_bar"], P["bar_vol"]
        idx = P["bar_idx"]
        c0, c1, n0, n1 = x[idx[:, 0]], x[idx[:, 1]], x[idx[:, 2]], x[idx[:, 3]]
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            imag = (c0 * n1 - c1 * n0) / vol
            t = imag / s
            bar = t * t * t - 3.0 * t * t + 3.0 * t
            res = 1.0 / bar - 1.0
            der = 3.0 * (imag * imag / (s * s * s)) - 6.0 * (imag / (s * s)) + 3.0 / s
            res = np.where(imag <= 0, np.inf, res)
            der = np.where(imag <= 0, np.inf, der)
            res = np.where(imag >= s, 0.0, res)
            der = np.where(imag >= s, 0.0, der)
            E = np.concatenate([self.A @ x - P["b"], res * w])
            if not compute_jacobian:
                return E, None
            dvec = -der / (bar * bar)
            dvec = np.where(np.abs(res) < 10e-9, 0.0, np.where(res == np.inf, np.inf, dvec))
            jb = np.stack([dvec * (n1 / vol) * w, dvec * (-n0 / vol) * w, dvec * (-c1 / vol) * w, dvec * (c0 / vol) * w], 1).ravel()
        return E, np.concatenate([self._avals, jb])[self._order]
