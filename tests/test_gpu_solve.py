"""GPU tier: SpMV, reductions and the LinearSolver concept (factorize/solve) vs the oracle's
SimplicialLLT restatement.  Bar (north_star): step vectors within 1e-10 relative in FP64."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import random_jacobian

pytestmark = pytest.mark.gpu

STEP_RTOL = 1e-10  # north_star: "step vectors ... within 1e-10 relative in FP64"


def _assembled(handle, orc, rng, m, n, k, lam=0.01):
    colptr, rowidx, vals = random_jacobian(rng, m, n, k)
    E = rng.standard_normal(m)
    handle.analyze(m, n, colptr, rowidx)
    handle.assemble(vals, E, lam)
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, _ = orc.damp_build(J, lam)
    return orc.ata(D), orc.jt_vec_neg(J, E)


def test_spmv_matches_scipy(handle, orc):
    rng = np.random.default_rng(0)
    H, g = _assembled(handle, orc, rng, 3000, 1000, 5)
    x = rng.standard_normal(1000)
    y = handle.spmv(x)
    yref = H.scipy().tocsr() @ x
    np.testing.assert_allclose(y, yref, rtol=1e-13, atol=1e-12)


def test_reductions(handle):
    rng = np.random.default_rng(1)
    for n in (1, 31, 257, 100003):
        v = rng.standard_normal(n)
        assert abs(handle.squared_norm(v) - float(v @ v)) <= 1e-13 * float(v @ v)
        assert handle.inf_norm(v) == np.abs(v).max()
    # deterministic: same bits on repeated calls
    v = rng.standard_normal(777777)
    assert handle.squared_norm(v) == handle.squared_norm(v)


@pytest.mark.parametrize("seed,m,n,k", [(0, 900, 300, 3), (1, 6000, 2500, 5)])
def test_pcg_step_matches_llt(handle, spb, orc, seed, m, n, k):
    rng = np.random.default_rng(seed)
    H, g = _assembled(handle, orc, rng, m, n, k)
    handle.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
    assert handle.factorize()
    x, iters, st = handle.solve()  # solves for the assembled gradient
    assert st == 0 and iters > 0
    F = orc.LLT(H, 1)
    assert F.ok
    xref = F.solve(g)
    assert np.linalg.norm(x - xref) <= STEP_RTOL * np.linalg.norm(xref)
    # explicit rhs path, host and device
    b = rng.standard_normal(n)
    x2, _, _ = handle.solve(b)
    assert np.linalg.norm(x2 - F.solve(b)) <= STEP_RTOL * np.linalg.norm(F.solve(b))
    import torch
    x3, _, _ = handle.solve(torch.tensor(b, device="cuda"))
    assert (x3.cpu().numpy() == x2).all()  # deterministic and pointer-kind independent


def test_pcg_iteration_count_close_to_cpu_twin(handle, spb, orc):
    J = orc.lf_matrix(40, 2, 20211)
    colptr, rowidx, vals = J.csc()
    m, n = J.shape
    E = J.scipy() @ np.ones(n) - 1.0
    handle.analyze(m, n, colptr, rowidx)
    handle.assemble(vals, E, 0.01)
    handle.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
    handle.factorize()
    x, iters, st = handle.solve()
    D, _ = orc.damp_build(J, 0.01)
    H = orc.ata(D)
    xr, itr, rr = orc.pcg_jacobi(H, orc.jt_vec_neg(J, E), 1e-13, 10000)
    assert abs(iters - itr) <= 2
    assert np.linalg.norm(x - xr) <= 1e-11 * np.linalg.norm(xr)


def test_linear_solver_concept_with_explicit_matrix(spb, orc):
    # LinearSolver::factorize(SparseMatrix) / solve(rhs, x): the pure drop-in path used when the
    # reference's own LMSolver.h passes dampJ^T*dampJ (LMSolver.h:136,143)
    rng = np.random.default_rng(3)
    m, n = 1200, 500
    colptr, rowidx, vals = random_jacobian(rng, m, n, 4)
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, _ = orc.damp_build(J, 0.01)
    H = orc.ata(D)
    ls = spb.B200SolverWrapper()
    assert ls.factorize(H.scipy()) is True
    b = rng.standard_normal(n)
    x = ls.solve(b)
    xref = orc.LLT(H, 1).solve(b)
    assert np.linalg.norm(x - xref) <= STEP_RTOL * np.linalg.norm(xref)
    # matrix right-hand side (EigenSolverWrapper.h:62-70)
    B = rng.standard_normal((n, 3))
    X = ls.solve(B)
    for c in range(3):
        xr = orc.LLT(H, 1).solve(B[:, c])
        assert np.linalg.norm(X[:, c] - xr) <= STEP_RTOL * np.linalg.norm(xr)
    # same pattern, new values: no re-analysis needed, results follow the values
    assert ls.factorize(2.0 * H.scipy()) is True
    np.testing.assert_allclose(ls.solve(b), 0.5 * x, rtol=1e-9)
    ls.handle.close()


def test_not_spd_is_reported(spb):
    # indefinite matrix: the reference's factorize returns false (EigenSolverWrapper.h:59)
    A = sp.csc_matrix(np.array([[1.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, 1.0]]))
    ls = spb.B200SolverWrapper()
    ok = ls.factorize(A)
    x, it, st = ls.handle.solve(np.array([1.0, -1.0, 0.5])) if ok else (None, 0, spb.capi.SPB_NOT_SPD)
    assert (not ok) or st == spb.capi.SPB_NOT_SPD
    ls.handle.close()


@pytest.mark.parametrize("case", ["random", "lf", "rosenbrock", "dense"])
def test_ic0_pcg_step_matches_llt_and_needs_fewer_iterations(spb, orc, case):
    rng = np.random.default_rng(7)
    if case == "random":
        m, n = 6000, 2500
        colptr, rowidx, vals = random_jacobian(rng, m, n, 5, cover_cols=True)
    elif case == "lf":
        J = orc.lf_matrix(48, 2, 20211)
        colptr, rowidx, vals = J.csc()
        m, n = J.shape
    elif case == "rosenbrock":
        nb = 200
        m = n = 2 * nb
        xx = rng.standard_normal(n) * 5
        colptr = (2 * np.arange(n + 1)).astype(np.int32)
        rowidx = (np.repeat(2 * np.arange(nb), 4) + np.tile([0, 1, 0, 1], nb)).astype(np.int32)
        vals = np.empty(4 * nb)
        vals[0::4] = -50.0 * xx[0::2]
        vals[1::4] = -1.0
        vals[2::4] = 25.0
        vals[3::4] = 0.0
    else:
        m, n = 300, 40
        A = rng.standard_normal((m, n))
        import scipy.sparse as sps
        S = sps.csc_matrix(A)
        S.sort_indices()
        colptr, rowidx, vals = S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data
    E = rng.standard_normal(m)
    h = spb.Handle(0)
    h.analyze(m, n, colptr, rowidx)
    h.assemble(vals, E, 0.01)
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, _ = orc.damp_build(J, 0.01)
    xref = orc.LLT(orc.ata(D), 1).solve(orc.jt_vec_neg(J, E))
    h.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
    assert h.factorize()
    xj, itj, stj = h.solve()
    h.set_solver(spb.SOLVER_PCG_IC0, 1e-13, 10000)
    assert h.factorize()
    xi, iti, sti = h.solve()
    assert sti == 0 and stj == 0
    assert np.linalg.norm(xi - xref) <= STEP_RTOL * np.linalg.norm(xref)
    assert iti <= itj  # a block-diagonal / dense H is solved (almost) exactly by IC(0)
    if case in ("lf", "random"):
        assert iti < 0.6 * itj
    # deterministic
    xi2, iti2, _ = h.solve()
    assert iti2 == iti and (xi2 == xi).all()
    h.close()


def test_lm_with_ic0_matches_oracle(spb, orc):
    h = spb.Handle(0)
    prob = spb.BuiltinProblem(h, "lf", g=32, rows_mult=2, seed=20211)
    ls = spb.B200SolverWrapper(h, spb.SOLVER_PCG_IC0)
    lm = spb.LMSolver()
    lm.init(ls, prob, spb.DiagonalDamping(0.01), 100)
    ret = lm.solve(True)
    ref = orc.lm_solve("lf", iparam0=32, iparam1=2, maxIterations=100)
    assert ret == ref.ret and lm.currIter == ref.currIter and lm.exit_path == ref.exit_path
    assert np.linalg.norm(lm.x - ref.x) <= 1e-10 * np.linalg.norm(ref.x)
    h.close()


def _gn_triplets(rng, m, n, k):
    """The legacy GN flow (GNSolver.h:44-95 in spirit): J by row-sorted triplets, J^T J by one
    upper-triangle triplet per product of two entries of a residual row (duplicates galore)."""
    colptr, rowidx, vals = random_jacobian(rng, m, n, k, cover_cols=True)
    J = sp.csc_matrix((vals, rowidx, colptr), shape=(m, n)).tocsr()
    J.sort_indices()
    hr, hc, hv = [], [], []
    for r in range(m):
        cols = J.indices[J.indptr[r]:J.indptr[r + 1]]
        v = J.data[J.indptr[r]:J.indptr[r + 1]]
        for a in range(len(cols)):
            for b in range(len(cols)):
                if cols[b] >= cols[a]:
                    hr.append(cols[a])
                    hc.append(cols[b])
                    hv.append(v[a] * v[b])
    # a small diagonal shift keeps the matrix safely positive definite
    for j in range(n):
        hr.append(j)
        hc.append(j)
        hv.append(0.05)
    return np.array(hr, np.int32), np.array(hc, np.int32), np.array(hv, np.float64)


@pytest.mark.parametrize("solver", [0, 1, 2])
def test_legacy_triplet_interface_matches_oracle(handle, spb, orc, solver):
    """spb_analyze_triplets / spb_factorize_triplets: H bit-exact against the oracle's
    setFromTriplets restatement (duplicates summed in insertion order, upper triangle mirrored),
    steps within 1e-10 of its SimplicialLLT."""
    rng = np.random.default_rng(40 + solver)
    n = 700
    hr, hc, hv = _gn_triplets(rng, 2100, n, 4)
    handle.set_solver(solver, 1e-13, 10000)
    handle.analyze_triplets(n, hr, hc, symmetric=True)
    assert handle.factorize_triplets(hv)
    # oracle: the same triplets with the strict upper part mirrored, in the same insertion order per entry
    off = hr != hc
    H = orc.from_triplets(n, n, np.concatenate([hr, hc[off]]), np.concatenate([hc, hr[off]]), np.concatenate([hv, hv[off]]))
    p, i, x = H.csc()
    hp, hi = handle.h_pattern()
    assert np.array_equal(hp, p) and np.array_equal(hi, i)
    assert np.array_equal(handle.h_values(), x)  # bit-exact
    b = rng.standard_normal(n)
    xs, _, st = handle.solve(b)
    assert st == 0
    F = orc.LLT(H, 1)
    assert F.ok
    xref = F.solve(b)
    assert np.linalg.norm(xs - xref) <= STEP_RTOL * np.linalg.norm(xref)
    # new values, same pattern: no re-analysis needed; a non-SPD matrix reports false
    hv2 = hv * rng.uniform(0.5, 1.5)
    assert handle.factorize_triplets(hv2)
    xs2, _, _ = handle.solve(b)
    assert np.linalg.norm(xs2 * (hv2[0] / hv[0]) - xref) <= 1e-9 * np.linalg.norm(xref)
    if solver == 0:
        assert not handle.factorize_triplets(-hv)


def test_legacy_triplet_interface_rejects_bad_input(handle, spb):
    with pytest.raises(spb.SpbError):
        handle.analyze_triplets(3, [1, 0], [0, 1], symmetric=True)  # below the diagonal in symmetric mode
    with pytest.raises(spb.SpbError):
        handle.analyze_triplets(3, [0, 1, 2, 0], [0, 1, 2, 1], symmetric=False)  # not structurally symmetric
    with pytest.raises(spb.SpbError):
        handle.analyze_triplets(3, [0, 1], [0, 1], symmetric=True)  # missing diagonal entry
    # unsymmetric listing of a symmetric matrix, python wrapper overloads
    ls = spb.B200SolverWrapper(handle=handle, solver=spb.SOLVER_CHOLESKY)
    assert ls.analyze([0, 1, 2, 0, 1], [0, 1, 2, 1, 0], False)
    assert ls.factorize(np.array([4.0, 5.0, 6.0, 1.0, 1.0]))
    x = ls.solve(np.array([1.0, 2.0, 3.0]))
    A = np.array([[4.0, 1, 0], [1, 5.0, 0], [0, 0, 6.0]])
    np.testing.assert_allclose(x, np.linalg.solve(A, [1.0, 2.0, 3.0]), rtol=1e-13)


def test_pcg_breakdown_and_nan_systems(spb):
    """Jacobi-PCG exits of the persistent kernel that leave the loop without a grid barrier (round-1 advisor
    finding: the ||x||^2 epilogue must not reuse the partial array a slower block is still summing).
    p.Hp <= 0 => SPB_NOT_SPD, the iterative analogue of LLT's pivot <= 0 (EigenSolverWrapper.h:59);
    NaN in H or in the right-hand side => NaN step with SPB_OK, like Eigen's LLT, which fails only on d <= 0
     -- LMSolver.h:147,161 then reject the step.  Repeated, on a system wide enough that
    every block of the cooperative grid has rows: a race would hang or give varying answers."""
    n = 60000
    main = np.full(n, 4.0)
    off = np.full(n - 1, -1.0)
    A = sp.diags([off, main, off], [-1, 0, 1], format="csc")
    b = np.sin(np.arange(n) * 0.01)
    h = spb.Handle(0)
    h.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
    cp, ri = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    # healthy reference
    h.set_matrix(n, cp, ri, A.data)
    h.factorize()
    x0, it0, st0 = h.solve(b)
    assert st0 == spb.capi.SPB_OK and np.abs(A @ x0 - b).max() < 1e-10
    # indefinite: negative diagonal block in the middle => some p.Hp <= 0
    vals = A.data.copy()
    D = A.copy()
    D.setdiag(np.where((np.arange(n) > n // 3) & (np.arange(n) < n // 2), -4.0, 4.0))
    for _ in range(5):
        h.set_matrix(n, cp, ri, D.tocsc().data)
        h.factorize()
        x, it, st = h.solve(b)
        assert st == spb.capi.SPB_NOT_SPD
    # NaN in a few entries of H (only a few blocks' partial sums see it)
    for _ in range(5):
        v = vals.copy()
        v[3 * (n // 2)] = np.nan
        h.set_matrix(n, cp, ri, v)
        h.factorize()
        x, it, st = h.solve(b)
        assert st == spb.capi.SPB_OK and np.isnan(x).all()
    # NaN right-hand side
    h.set_matrix(n, cp, ri, vals)
    h.factorize()
    bn = b.copy()
    bn[17] = np.nan
    x, it, st = h.solve(bn)
    assert st == spb.capi.SPB_OK and np.isnan(x).all()
    # and the handle still solves the healthy system to the same bits
    x1, it1, st1 = h.solve(b)
    assert st1 == spb.capi.SPB_OK and it1 == it0 and np.array_equal(x1, x0)
    h.close()


@pytest.mark.parametrize("variant", ["classic", "sr"])
def test_relaxed_precision_products_keep_the_step(spb, monkeypatch, variant):
    """Late PCG iterations read the matrix values from an FP32 copy (inexact-Krylov relaxation: once ||r||/||b|| <= 1e-7
    a product error of 6e-8 ||H|| is below what an rtol of 1e-13 allows).  Against the all-FP64 solve on the same system:
    same iteration count (+-1), step within 1e-12 relative (north_star: 1e-10), true FP64 residual at the rtol level."""
    import torch
    monkeypatch.setenv("SPB_PCG_VARIANT", variant)
    g = 160
    out = {}
    for relax in ("0", "1e6"):
        monkeypatch.setenv("SPB_PCG_RELAX", relax)
        h = spb.Handle(0)
        prob = spb.BuiltinProblem(h, "lf", g=g, rows_mult=2, seed=20211)
        m, n, colptr, rowidx = prob.pattern()
        T = prob.c_traits()
        x = torch.ones(n, dtype=torch.float64, device="cuda")
        E = torch.empty(m, dtype=torch.float64, device="cuda")
        Jv = torch.empty(int(colptr[-1]), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        T.objective_jacobian(T.user, x.data_ptr(), E.data_ptr(), Jv.data_ptr(), None)
        torch.cuda.synchronize()
        h.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
        h.analyze(m, n, colptr, rowidx)
        res = []
        for lam in (0.01, 1e-4):
            h.assemble(Jv, E, lam)
            h.factorize()
            d, it, st = h.solve(None)
            assert st == spb.capi.SPB_OK
            res.append((d.copy(), it, h.residual_check(d)))
        out[relax] = res
        prob.close(); h.close()
    for (d0, it0, r0), (d1, it1, r1) in zip(out["0"], out["1e6"]):
        assert abs(it0 - it1) <= 1 and it0 > 30
        assert np.linalg.norm(d1 - d0) <= 1e-12 * np.linalg.norm(d0)
        assert r0 <= 2e-13 and r1 <= 2e-13
