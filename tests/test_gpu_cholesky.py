"""GPU tier: supernodal Cholesky (the SimplicialLLT replacement) vs the oracle's LLT.
Bar: step vectors within 1e-10 relative; factorisation failure semantics identical
(false iff a pivot <= 0; NaN pivots do not fail)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import random_jacobian

pytestmark = pytest.mark.gpu

STEP_RTOL = 1e-10


def _system(orc, rng, m, n, k, lam=0.01, **kw):
    colptr, rowidx, vals = random_jacobian(rng, m, n, k, **kw)
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, _ = orc.damp_build(J, lam)
    return (m, n, colptr, rowidx, vals), orc.ata(D), J


@pytest.mark.parametrize("seed,m,n,k", [(0, 300, 100, 3), (1, 3000, 1200, 4), (2, 20000, 9000, 3)])
def test_solve_matches_oracle_llt(spb, orc, seed, m, n, k):
    rng = np.random.default_rng(seed)
    (m, n, colptr, rowidx, vals), H, J = _system(orc, rng, m, n, k, cover_cols=True)
    h = spb.Handle(0)
    h.set_solver(spb.SOLVER_CHOLESKY)
    E = rng.standard_normal(m)
    h.analyze(m, n, colptr, rowidx)
    h.assemble(vals, E, 0.01)
    assert h.factorize()
    x, it, st = h.solve()
    assert st == 0 and it == 0
    F = orc.LLT(H, 1)
    xref = F.solve(orc.jt_vec_neg(J, E))
    assert np.linalg.norm(x - xref) <= STEP_RTOL * np.linalg.norm(xref)
    info = h.factor_info()
    assert info["nnzL"] >= n and info["nsuper"] >= 1
    # second factorisation with new values reuses the symbolic phase
    h.assemble(2.0 * vals, E, 0.5)
    assert h.factorize()
    D2, _ = orc.damp_build(orc.from_csc(m, n, colptr, rowidx, 2.0 * vals), 0.5)
    x2, _, _ = h.solve()
    xr2 = orc.LLT(orc.ata(D2), 1).solve(orc.jt_vec_neg(orc.from_csc(m, n, colptr, rowidx, 2.0 * vals), E))
    assert np.linalg.norm(x2 - xr2) <= STEP_RTOL * np.linalg.norm(xr2)
    h.close()


@pytest.mark.parametrize("g", [12, 64, 150])
def test_lf_workload_direct(spb, orc, g):
    J = orc.lf_matrix(g, 2, 20211)
    colptr, rowidx, vals = J.csc()
    m, n = J.shape
    E = J.scipy() @ np.ones(n) - 1.0
    h = spb.Handle(0)
    h.set_solver(spb.SOLVER_CHOLESKY)
    h.analyze(m, n, colptr, rowidx)
    h.assemble(vals, E, 0.01)
    assert h.factorize()
    x, _, _ = h.solve()
    Hs = h.h_scipy()
    g_ = orc.jt_vec_neg(J, E)
    # residual check at every size; oracle LLT comparison where it is quick
    assert np.linalg.norm(Hs @ x - g_) <= 1e-12 * np.linalg.norm(g_)
    if g <= 64:
        D, _ = orc.damp_build(J, 0.01)
        xref = orc.LLT(orc.ata(D), 1).solve(g_)
        assert np.linalg.norm(x - xref) <= STEP_RTOL * np.linalg.norm(xref)
    # cross-check against the PCG path on the same handle
    h.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
    assert h.factorize()
    xp, _, _ = h.solve()
    assert np.linalg.norm(x - xp) <= STEP_RTOL * np.linalg.norm(x)
    h.close()


def test_dense_and_block_diagonal_patterns(spb, orc):
    # dense H (the reference linear_function: n=250) -> a single dense front; block-diagonal H
    # (Rosenbrock) -> a forest of tiny fronts
    rng = np.random.default_rng(5)
    m, n = 1000, 250
    A = np.vstack([np.eye(n) - 2.0 / m, -2.0 / m * np.ones((m - n, n))])
    colptr = (m * np.arange(n + 1)).astype(np.int32)
    rowidx = np.tile(np.arange(m), n).astype(np.int32)
    vals = np.ascontiguousarray(A.T).ravel()
    E = A @ np.ones(n) - 1.0
    h = spb.Handle(0)
    h.set_solver(spb.SOLVER_CHOLESKY)
    h.analyze(m, n, colptr, rowidx)
    h.assemble(vals, E, 0.01)
    assert h.factorize()
    x, _, _ = h.solve()
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, _ = orc.damp_build(J, 0.01)
    xref = orc.LLT(orc.ata(D), 1).solve(orc.jt_vec_neg(J, E))
    assert np.linalg.norm(x - xref) <= STEP_RTOL * np.linalg.norm(xref)
    nb = 300
    nn = 2 * nb
    xx = rng.standard_normal(nn) * 10
    colptr = (2 * np.arange(nn + 1)).astype(np.int32)
    rowidx = (np.repeat(2 * np.arange(nb), 4) + np.tile([0, 1, 0, 1], nb)).astype(np.int32)
    vals = np.empty(4 * nb)
    vals[0::4] = -50.0 * xx[0::2]
    vals[1::4] = -1.0
    vals[2::4] = 25.0
    vals[3::4] = 0.0
    E = rng.standard_normal(nn)
    h.analyze(nn, nn, colptr, rowidx)
    h.assemble(vals, E, 0.01)
    assert h.factorize()
    x, _, _ = h.solve()
    J = orc.from_csc(nn, nn, colptr, rowidx, vals)
    D, _ = orc.damp_build(J, 0.01)
    xref = orc.LLT(orc.ata(D), 1).solve(orc.jt_vec_neg(J, E))
    assert np.linalg.norm(x - xref) <= STEP_RTOL * np.linalg.norm(xref)
    h.close()


def test_failure_semantics(spb, orc):
    # indefinite => false; all-zero Jacobian column => H_jj = 0 => false (quirk F); NaN => passes (quirk H)
    ls = spb.B200SolverWrapper(solver=spb.SOLVER_CHOLESKY)
    A = sp.csc_matrix(np.array([[1.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, 1.0]]))
    assert ls.factorize(A) is False
    assert orc.LLT(orc.from_csc(3, 3, A.indptr, A.indices, A.data), 0).ok is False
    A = sp.csc_matrix(np.array([[4.0, 1.0, 0.0], [1.0, 3.0, 0.0], [0.0, 0.0, 2.0]]))
    assert ls.factorize(A) is True
    np.testing.assert_allclose(ls.solve(np.array([1.0, 2.0, 3.0])), np.linalg.solve(A.toarray(), [1.0, 2.0, 3.0]), rtol=1e-13)
    Z = sp.csc_matrix((np.array([1.0, 0.0]), np.array([0, 1]), np.array([0, 1, 2])), shape=(2, 2))
    assert ls.factorize(Z) is False
    N = sp.csc_matrix((np.array([np.nan, 1.0]), np.array([0, 1]), np.array([0, 1, 2])), shape=(2, 2))
    assert ls.factorize(N) is True  # d<=0 is false for NaN
    assert orc.LLT(orc.from_csc(2, 2, N.indptr, N.indices, N.data), 0).ok is True
    ls.handle.close()


def test_lm_with_direct_solver_matches_oracle(spb, orc):
    h = spb.Handle(0)
    for kind, kw, okw in [("rosenbrock", dict(nblocks=500), dict(iparam0=500)),
                          ("lf", dict(g=24, rows_mult=2, seed=20211), dict(iparam0=24, iparam1=2))]:
        prob = spb.BuiltinProblem(h, kind, **kw)
        ls = spb.B200SolverWrapper(h, spb.SOLVER_CHOLESKY)
        lm = spb.LMSolver()
        lm.init(ls, prob, spb.DiagonalDamping(0.01), 1000)
        ret = lm.solve(True)
        ref = orc.lm_solve(kind, maxIterations=1000, **okw)
        assert ret == ref.ret and lm.exit_path == ref.exit_path
        assert lm.currIter == ref.currIter and lm.factorizations == ref.factorizations
        np.testing.assert_allclose(lm.x, ref.x, rtol=1e-10, atol=1e-12)
    h.close()


def test_singular_column_returns_false_from_lm(spb, orc):
    class T:
        xSize, ESize = 2, 3

        def initial_solution(self):
            return np.zeros(2)

        def objective_jacobian(self, x, cj):
            E = np.array([x[0] - 1.0, 2 * x[0], 0.0])
            J = sp.csc_matrix((np.array([1.0, 2.0, 0.0]), np.array([0, 1, 2]), np.array([0, 2, 3])), shape=(3, 2))
            return E, (J if cj else None)

    ls = spb.B200SolverWrapper(solver=spb.SOLVER_CHOLESKY)
    lm = spb.LMSolver()
    lm.init(ls, T(), spb.DiagonalDamping(0.01), 10)
    assert lm.solve(True) is False and lm.exit_path == "factorize_failed"
    ls.handle.close()


@pytest.mark.parametrize("g", [40, 150])
def test_tma_staged_syrk_gives_the_same_factor(spb, orc, g, monkeypatch):
    """SPB_CHOL_TMA=1 stages the SYRK/GEMM operand tiles with TMA bulk copies (cp.async.bulk + mbarrier) instead of
    cp.async; the DMMA sequence and its operand order are the same, so the solve must return the same bits
    (odd panel leading dimensions exercise the alignment shift of the bulk copies)."""
    J = orc.lf_matrix(g, 2, 20211)
    colptr, rowidx, vals = J.csc()
    m, n = J.shape
    E = J.scipy() @ np.ones(n) - 1.0
    xs = []
    for tma in ("0", "1"):
        monkeypatch.setenv("SPB_CHOL_TMA", tma)
        h = spb.Handle(0)
        h.set_solver(spb.SOLVER_CHOLESKY)
        h.analyze(m, n, colptr, rowidx)
        h.assemble(vals, E, 0.01)
        assert h.factorize()
        x, _, _ = h.solve()
        xs.append(np.array(x, copy=True))
        h.close()
    assert np.array_equal(xs[0].view(np.uint64), xs[1].view(np.uint64))


@pytest.mark.parametrize("kind", ["chain", "star_of_pairs", "satellites"])
def test_tiny_fronts_with_children_and_condensed_groups(spb, orc, kind):
    """Low-degree structures: every front is tiny (<= 16 pivots, <= 32 rows) and most have children, or the ordering
    condenses satellite groups first -- the warp-per-front kernels (factor, extend-add of small children, both solves)
    against the oracle's LLT."""
    rng = np.random.default_rng(4)
    rows, cols = [], []
    if kind == "chain":            # H tridiagonal: residual r couples unknowns r and r+1
        n = 700
        for r in range(n - 1):
            rows += [r, r]
            cols += [r, r + 1]
        m = n - 1
    elif kind == "star_of_pairs":  # pairs of unknowns hanging off a few hubs
        hubs, n = 12, 12 + 2 * 300
        r = 0
        for p in range(300):
            a, b, hub = hubs + 2 * p, hubs + 2 * p + 1, p % hubs
            rows += [r, r, r + 1, r + 1, r + 1]
            cols += [a, b, a, b, hub]
            r += 2
        for hub in range(hubs):    # couple the hubs in a ring
            rows += [r, r]
            cols += [hub, (hub + 1) % hubs]
            r += 1
        m = r
    else:                          # mesh-like: 8 satellites per cell coupled to the cell's 4 grid nodes
        gsz = 14
        nodes = gsz * gsz
        cells = (gsz - 1) * (gsz - 1)
        n = nodes + 8 * cells
        r = 0
        for cy in range(gsz - 1):
            for cx in range(gsz - 1):
                c = cy * (gsz - 1) + cx
                corner = [cy * gsz + cx, cy * gsz + cx + 1, (cy + 1) * gsz + cx, (cy + 1) * gsz + cx + 1]
                for s in range(8):
                    sat = nodes + 8 * c + s
                    rows += [r, r, r]
                    cols += [sat, corner[s % 4], nodes + 8 * c + (s + 1) % 8]
                    r += 1
        for v in range(nodes):     # anchor every node
            rows.append(r)
            cols.append(v)
            r += 1
        m = r
    vals = rng.uniform(0.5, 1.5, len(rows)) * rng.choice([-1.0, 1.0], len(rows))
    Jm = sp.csc_matrix((vals, (rows, cols)), shape=(m, n))
    Jm.sum_duplicates()
    Jm.sort_indices()
    colptr, rowidx, v = Jm.indptr.astype(np.int32), Jm.indices.astype(np.int32), np.ascontiguousarray(Jm.data)
    J = orc.from_csc(m, n, colptr, rowidx, v)
    D, _ = orc.damp_build(J, 0.3)
    H = orc.ata(D)
    E = rng.standard_normal(m)
    h = spb.Handle(0)
    h.set_solver(spb.SOLVER_CHOLESKY)
    h.analyze(m, n, colptr, rowidx)
    h.assemble(v, E, 0.3)
    assert h.factorize()
    x, it, st = h.solve()
    assert st == 0
    xref = orc.LLT(H, 1).solve(orc.jt_vec_neg(J, E))
    assert np.linalg.norm(x - xref) <= STEP_RTOL * np.linalg.norm(xref)
    # second factorisation (graph replay) with other values
    h.assemble(1.5 * v, E, 0.1)
    assert h.factorize()
    x2, _, _ = h.solve()
    D2, _ = orc.damp_build(orc.from_csc(m, n, colptr, rowidx, 1.5 * v), 0.1)
    x2ref = orc.LLT(orc.ata(D2), 1).solve(orc.jt_vec_neg(orc.from_csc(m, n, colptr, rowidx, 1.5 * v), E))
    assert np.linalg.norm(x2 - x2ref) <= STEP_RTOL * np.linalg.norm(x2ref)
    h.close()
