"""Synthetic stand-in for config C3 (seamless integration), test infrastructure only.

The reference's real input is an un-shipped .srl file and its traits need libigl/Directional/MATLAB
(SURVEY 2 #10), so this builds a problem with the SAME Jacobian block structure and weights on a
regular grid triangulation (SIInitialSolutionTraits.h:115-248, 255-259):

    unknowns  = [ param functions (N/2 per vertex) ; field (2N per face) ]        (no seams: UExt = I)
    residuals = [ wInt * (field - L * G2 x)          2N|F| rows, 4 nnz   (:127,179-189)
                  wClose * (field - field0)           2N|F| rows, 1 nnz   (:191-197)
                  wConst * (x_fixed - c)              |fixed| rows        (:199-205)
                  wBarr * barrier(imagProduct/vol) ]  N|F| rows, 4 nnz, values 0 / finite / inf (:148-172,216-233)
with wInt = wConst = 10e3, wClose = 1, wBarr = 1e-4, s = 0.1.  A Python SolverTraits object
(objective_jacobian form) that both the oracle's callback traits and the GPU LMSolver consume.
"""
import numpy as np
import scipy.sparse as sp


class SeamlessLikeTraits:
    N = 4
    wInt, wConst, wBarr, wClose, s = 10e3, 10e3, 0.0001, 1.0, 0.1

    def __init__(self, nx=8, ny=8, seed=20212, flip_face=None):
        rng = np.random.default_rng(seed)
        N = self.N
        V = np.array([(i, j) for j in range(ny) for i in range(nx)], float)
        F = []
        for j in range(ny - 1):
            for i in range(nx - 1):
                a, b, c, d = j * nx + i, j * nx + i + 1, (j + 1) * nx + i, (j + 1) * nx + i + 1
                F += [(a, b, d), (a, d, c)]
        F = np.array(F)
        nV, nF = len(V), len(F)
        self.nV, self.nF = nV, nF
        nfun = N // 2
        self.nx_vars = nfun * nV
        self.nfield = 2 * N * nF
        self.xSize = self.nx_vars + self.nfield
        # per-face gradient of the nfun param functions: G2 maps x -> 2N|F| (vectors j and j+N/2 are +/- the gradient)
        rows, cols, vals = [], [], []
        for f, (a, b, c) in enumerate(F):
            M = np.linalg.inv(np.array([V[b] - V[a], V[c] - V[a]]))  # grad = M @ [ub-ua, uc-ua]
            for j in range(N):
                fun, sign = j % nfun, (1.0 if j < nfun else -1.0)
                for comp in range(2):
                    r = 2 * N * f + 2 * j + comp
                    for vtx, w in ((a, -M[comp, 0] - M[comp, 1]), (b, M[comp, 0]), (c, M[comp, 1])):
                        rows.append(r); cols.append(fun * nV + vtx); vals.append(sign * w)
        self.G2 = sp.csr_matrix((vals, (rows, cols)), shape=(self.nfield, self.nx_vars))
        self.L = 1.0
        # smooth random field: gradient of two smooth functions + noise, so cross products stay positive
        xs, ys = V[:, 0], V[:, 1]
        u0 = xs + 0.15 * np.sin(0.7 * ys) + 0.02 * rng.standard_normal(nV)
        v0 = ys + 0.15 * np.sin(0.6 * xs) + 0.02 * rng.standard_normal(nV)
        self.x_init = np.concatenate([u0, v0])
        self.field0 = self.L * (self.G2 @ self.x_init) + 0.01 * rng.standard_normal(self.nfield)
        if flip_face is not None:  # make one barrier term start in the infeasible region (inf residual / Jacobian)
            k = 2 * N * flip_face
            self.field0[k:k + 2] = -self.field0[k:k + 2]
        self.vol = np.ones((nF, N))
        for f in range(nF):
            for j in range(N):
                c, nvec = self.field0[2 * N * f + 2 * j:2 * N * f + 2 * j + 2], self.field0[2 * N * f + 2 * ((j + 1) % N):2 * N * f + 2 * ((j + 1) % N) + 2]
                self.vol[f, j] = abs(c[0] * nvec[1] - c[1] * nvec[0]) + 1e-3
        self.fixed = np.array([0, nV])  # the two function values at vertex 0
        self.fixed_vals = self.x_init[self.fixed].copy()
        self.ESize = self.nfield + self.nfield + len(self.fixed) + N * nF
        # fixed Jacobian pattern (value-independent): stack the four blocks
        gInt = sp.hstack([-self.L * self.G2, sp.identity(self.nfield, format="csr")]).tocsr()
        gClose = sp.hstack([sp.csr_matrix((self.nfield, self.nx_vars)), sp.identity(self.nfield, format="csr")]).tocsr()
        gConst = sp.csr_matrix((np.ones(len(self.fixed)), (np.arange(len(self.fixed)), self.fixed)), shape=(len(self.fixed), self.xSize))
        self._gInt, self._gClose, self._gConst = gInt * self.wInt, gClose * self.wClose, gConst * self.wConst
        br, bc = [], []
        for f in range(nF):
            for j in range(N):
                r = N * f + j
                for q in (2 * N * f + 2 * j, 2 * N * f + 2 * j + 1, 2 * N * f + 2 * ((j + 1) % N), 2 * N * f + 2 * ((j + 1) % N) + 1):
                    br.append(r); bc.append(self.nx_vars + q)
        self._brow, self._bcol = np.array(br), np.array(bc)
        pat = sp.vstack([self._gInt, self._gClose, self._gConst,
                         sp.csr_matrix((np.ones(len(br)), (br, bc)), shape=(N * nF, self.xSize))]).tocsc()
        pat.sort_indices()
        self.colptr, self.rowidx = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
        self.calls = 0

    def initial_solution(self):
        return np.concatenate([self.x_init, self.field0])

    def _barrier(self, field):
        N, nF, s = self.N, self.nF, self.s
        fb = np.zeros(N * nF)
        jv = np.zeros((N * nF, 4))
        for f in range(nF):
            for j in range(N):
                c = field[2 * N * f + 2 * j:2 * N * f + 2 * j + 2]
                nvec = field[2 * N * f + 2 * ((j + 1) % N):2 * N * f + 2 * ((j + 1) % N) + 2]
                vol = self.vol[f, j]
                imag = (c[0] * nvec[1] - c[1] * nvec[0]) / vol
                t = imag / s
                bar = t * t * t - 3.0 * t * t + 3.0 * t
                with np.errstate(divide="ignore", invalid="ignore"):
                    res = 1.0 / bar - 1.0
                    der = 3.0 * (imag * imag / (s * s * s)) - 6.0 * (imag / (s * s)) + 3.0 / s
                    if imag <= 0:
                        res, der = np.inf, np.inf
                    if imag >= s:
                        res, der = 0.0, 0.0
                    dvec = -der / (bar * bar)
                    if abs(res) < 10e-9:
                        dvec = 0.0
                    elif res == np.inf:
                        dvec = np.inf
                    fb[N * f + j] = res
                    jv[N * f + j] = dvec * np.array([nvec[1] / vol, -nvec[0] / vol, -c[1] / vol, c[0] / vol])
        return fb, jv

    def objective_jacobian(self, x, compute_jacobian):
        self.calls += 1
        xs, field = x[:self.nx_vars], x[self.nx_vars:]
        fInt = field - self.L * (self.G2 @ xs)
        fClose = field - self.field0
        fConst = x[self.fixed] - self.fixed_vals
        fb, jv = self._barrier(field)
        E = np.concatenate([fInt * self.wInt, fClose * self.wClose, fConst * self.wConst, fb * self.wBarr])
        if not compute_jacobian:
            return E, None
        with np.errstate(invalid="ignore"):
            bvals = (jv * self.wBarr).ravel()
        # assemble values in the fixed CSC pattern (explicit zeros / inf / nan kept)
        r0 = 2 * self.nfield + len(self.fixed)
        coo_r = np.concatenate([self._gInt.tocoo().row, self._gClose.tocoo().row + self.nfield, self._gConst.tocoo().row + 2 * self.nfield,
                                self._brow + r0])
        coo_c = np.concatenate([self._gInt.tocoo().col, self._gClose.tocoo().col, self._gConst.tocoo().col, self._bcol])
        coo_v = np.concatenate([self._gInt.tocoo().data, self._gClose.tocoo().data, self._gConst.tocoo().data, bvals])
        order = np.lexsort((coo_r, coo_c))
        assert (coo_r[order] == self.rowidx).all()
        return E, coo_v[order]
