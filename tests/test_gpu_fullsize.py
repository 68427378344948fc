"""GPU tier: BASELINE's full size (config C2, LF-1M: n = 1 048 576, m = 2n, nnzJ = 16.8 M) through
size-independent properties -- the CPU oracle takes minutes per factorisation there."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(spb):
    import torch
    h = spb.Handle(0, torch.cuda.current_stream().cuda_stream)
    prob = spb.BuiltinProblem(h, "lf", g=1024, rows_mult=2, seed=20211)
    m, n, colptr, rowidx = prob.pattern()
    T = prob.c_traits()
    x = torch.ones(n, dtype=torch.float64, device="cuda")
    E = torch.empty(m, dtype=torch.float64, device="cuda")
    Jv = torch.empty(int(colptr[-1]), dtype=torch.float64, device="cuda")
    T.objective_jacobian(T.user, x.data_ptr(), E.data_ptr(), Jv.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    h.analyze(m, n, colptr, rowidx)
    yield dict(h=h, prob=prob, m=m, n=n, colptr=colptr, rowidx=rowidx, E=E, Jv=Jv)
    prob.close()
    h.close()


def test_normal_matrix_is_linear_in_the_jacobian(c2):
    # H x == J^T (J x) + diag(sqrt(d)^2) x for a random x (host sparse products as the referee)
    h, n, m = c2["h"], c2["n"], c2["m"]
    h.assemble(c2["Jv"], c2["E"], 0.01)
    J = sp.csc_matrix((c2["Jv"].cpu().numpy(), c2["rowidx"], c2["colptr"]), shape=(m, n))
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n)
    y = h.spmv(x)
    d = h.damping()
    yref = J.T @ (J @ x) + np.sqrt(d) ** 2 * x
    assert np.linalg.norm(y - yref) <= 1e-13 * np.linalg.norm(yref)
    # gradient and damping: rhs = -J^T f ; d = lambda * column sums of squares ; foo = max |rhs|
    g = h.rhs()
    gref = -(J.T @ c2["E"].cpu().numpy())
    assert np.linalg.norm(g - gref) <= 1e-13 * np.linalg.norm(gref)
    np.testing.assert_allclose(d, 0.01 * np.asarray(J.multiply(J).sum(0)).ravel(), rtol=1e-13)
    assert h.h_nnz() == 25918392  # structural pattern size of this workload (~25.9 M)


def test_direct_and_iterative_solvers_agree(c2, spb):
    h = c2["h"]
    h.assemble(c2["Jv"], c2["E"], 0.01)
    h.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
    assert h.factorize()
    xp, iters, st = h.solve()
    assert st == 0 and 40 < iters < 200
    h.set_solver(spb.SOLVER_CHOLESKY)
    assert h.factorize()
    xc, _, _ = h.solve()
    assert np.linalg.norm(xp - xc) <= 1e-10 * np.linalg.norm(xc)
    r = h.spmv(xc) - h.rhs()
    assert np.linalg.norm(r) <= 1e-13 * np.linalg.norm(h.rhs())
    h.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)


def test_lm_converges_to_the_least_squares_solution(c2, spb):
    h, prob = c2["h"], c2["prob"]
    ls = spb.B200SolverWrapper(h, spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
    lm = spb.LMSolver()
    lm.init(ls, prob, spb.DiagonalDamping(0.01), 100)
    ret = lm.solve(True, dedupe_evaluations=True)
    assert lm.exit_path in ("func_tol", "direction_small", "foo")
    e = lm.trace["newEnergy2"]
    acc = lm.trace["accepted"] > 0
    assert (np.diff(np.concatenate([[lm.trace["prevEnergy2"][0]], e[acc]])) < 0).all()  # accepted steps decrease the energy
    assert lm.currIter <= 10
    # first-order optimality of the linear least-squares problem at the returned x
    assert lm.trace["foo"][-1] < 1e-3 * lm.trace["foo"][0]


def test_c2_assembly_is_bit_exact_against_the_oracle(c2, orc):
    """The whole C2 normal matrix (25.9 M entries), gradient and damping vector against the oracle's restatement of
    DiagonalDamping.h:31-43 + `dampJ.transpose()*dampJ` (LMSolver.h:136) + `-(J^T E)` (LMSolver.h:117): pattern and
    values bit for bit at BASELINE's size (the oracle's sparse product takes seconds there, not minutes)."""
    h, n, m = c2["h"], c2["n"], c2["m"]
    foo = h.assemble(c2["Jv"], c2["E"], 0.01)
    vals = c2["Jv"].cpu().numpy()
    E = c2["E"].cpu().numpy()
    J = orc.from_csc(m, n, c2["colptr"], c2["rowidx"], vals)
    D, d = orc.damp_build(J, 0.01)
    hp, hi, hx = orc.ata(D).csc()
    cp, ri = h.h_pattern()
    assert np.array_equal(cp, hp) and np.array_equal(ri, hi), "H pattern differs at C2 size"
    hv = h.h_values()
    bits = lambda a: np.ascontiguousarray(a, np.float64).view(np.uint64)
    assert np.array_equal(bits(hv), bits(hx)), "H values not bit-exact at C2 size"
    g = orc.jt_vec_neg(J, E)
    assert np.array_equal(bits(h.rhs()), bits(g)) and np.array_equal(bits(h.damping()), bits(d))
    assert foo == np.abs(g).max()
