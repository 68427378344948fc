"""GPU tier: incremental re-analysis (spb_analyze_update) when the Jacobian gains one single-entry residual row
(IterativeRoundingTraits.h:76-79 appends one constraint row per outer iteration of
benchmark/seamless_integration/main.cpp:69-80; rows below it move down by one).

Bar: the patched handle assembles H / rhs / d bit-identically to the oracle on the NEW Jacobian (so identically to a
from-scratch analysis), the H pattern is the from-scratch one, the kept Cholesky symbolic state still solves, and
anything that is not a single-entry appended row falls back to the full analysis with the same results."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import random_jacobian

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def _insert_row(J, rho, cols, vals):
    """J (scipy csr/csc, m x n) with one new row inserted BEFORE old row rho (rho == m appends)."""
    J = J.tocsr()
    m, n = J.shape
    new = sp.csr_matrix((np.asarray(vals, float), (np.zeros(len(cols), int), np.asarray(cols, int))), shape=(1, n))
    out = sp.vstack([J[:rho], new, J[rho:]]).tocsc()
    out.sort_indices()
    return out


def _csc(J):
    J = J.tocsc()
    J.sort_indices()
    return J.shape[0], J.shape[1], J.indptr.astype(np.int32), J.indices.astype(np.int32), J.data.astype(np.float64)


def _check_against_oracle(handle, orc, J, E, lam):
    m, n, colptr, rowidx, vals = _csc(J)
    foo = handle.assemble(vals, E, lam)
    Jo = orc.from_csc(m, n, colptr, rowidx, vals)
    D, d = orc.damp_build(Jo, lam)
    hp, hi, hx = orc.ata(D).csc()
    cp, ri = handle.h_pattern()
    assert (cp == hp).all() and (ri == hi).all(), "H pattern differs"
    assert (_bits(handle.h_values()) == _bits(hx)).all(), "H values not bit-exact after the update"
    g = orc.jt_vec_neg(Jo, E)
    assert (_bits(handle.rhs()) == _bits(g)).all(), "rhs not bit-exact after the update"
    assert (_bits(handle.damping()) == _bits(d)).all(), "damping not bit-exact after the update"
    assert foo == np.abs(g).max()


def _lf(orc, g):
    J = orc.lf_matrix(g, 2, 20211)
    return J.scipy().tocsc()


@pytest.mark.parametrize("where", ["end", "middle", "first", "chain"])
def test_single_entry_row_is_patched(handle, orc, where):
    rng = np.random.default_rng(5)
    J = _lf(orc, 40)
    m, n = J.shape
    handle.analyze(*_csc(J)[:4])
    handle.assemble(_csc(J)[4], rng.standard_normal(m), 0.01)
    picks = {"end": [(m, n - 1)], "middle": [(m // 2 + 3, n // 3)], "first": [(0, 0)],
             "chain": [(m, 5), (m // 2, 5), (17, n - 1), (m + 2, 777), (m + 3, 778)]}[where]
    for rho, c in picks:
        J = _insert_row(J, rho, [c], [rng.standard_normal()])
        how = handle.analyze_update(*_csc(J)[:4])
        assert how == 1, "a single-entry appended row must take the incremental path (got %d)" % how
        _check_against_oracle(handle, orc, J, rng.standard_normal(J.shape[0]), 0.01)
    # unchanged pattern: nothing to do
    assert handle.analyze_update(*_csc(J)[:4]) == 0


def test_random_pattern_with_explicit_zeros_and_empty_columns(handle, orc):
    rng = np.random.default_rng(11)
    m, n = 3000, 1200
    colptr, rowidx, vals = random_jacobian(rng, m, n, 4, explicit_zero_frac=0.05, empty_cols=[0, 31, 700])
    J = sp.csc_matrix((vals, rowidx, colptr), shape=(m, n))
    handle.analyze(m, n, colptr, rowidx)
    for rho, c in [(m, 31), (100, 0), (2000, 1199), (5, 600)]:  # includes columns that were empty
        J = _insert_row(J, rho, [c], [1.5])
        how = handle.analyze_update(*_csc(J)[:4])
        # an entry in a previously EMPTY column adds a diagonal entry that is already structural (full diagonal): patched
        assert how in (1, 2)
        _check_against_oracle(handle, orc, J, rng.standard_normal(J.shape[0]), 0.3)


def test_other_changes_fall_back_to_the_full_analysis(handle, orc):
    rng = np.random.default_rng(3)
    J = _lf(orc, 24)
    m, n = J.shape
    handle.analyze(*_csc(J)[:4])
    # a row with two entries creates a new H entry: full analysis
    J2 = _insert_row(J, m, [3, 400], [1.0, -2.0])
    assert handle.analyze_update(*_csc(J2)[:4]) == 2
    _check_against_oracle(handle, orc, J2, rng.standard_normal(m + 1), 0.01)
    # two rows at once: full analysis
    J3 = _insert_row(_insert_row(J2, 0, [1], [1.0]), 5, [2], [1.0])
    assert handle.analyze_update(*_csc(J3)[:4]) == 2
    _check_against_oracle(handle, orc, J3, rng.standard_normal(m + 3), 0.01)
    # a removed row: full analysis
    assert handle.analyze_update(*_csc(J2)[:4]) == 2
    _check_against_oracle(handle, orc, J2, rng.standard_normal(m + 1), 0.01)


@pytest.mark.parametrize("solver", ["cholesky", "pcg"])
def test_solves_after_the_update_use_the_kept_symbolic_state(handle, orc, solver):
    import saddlepoint_b200 as spb
    from saddlepoint_b200 import capi
    rng = np.random.default_rng(9)
    J = _lf(orc, 32)
    m, n = J.shape
    handle.set_solver(capi.SOLVER_CHOLESKY if solver == "cholesky" else capi.SOLVER_PCG_JACOBI, 1e-13, 10000)
    handle.analyze(*_csc(J)[:4])
    handle.assemble(_csc(J)[4], rng.standard_normal(m), 0.01)
    handle.factorize()
    handle.solve(handle.rhs())
    for rho, c in [(m, 10), (m // 2, 500)]:
        J = _insert_row(J, rho, [c], [3.0])
        assert handle.analyze_update(*_csc(J)[:4]) == 1
        E = rng.standard_normal(J.shape[0])
        handle.assemble(_csc(J)[4], E, 0.01)
        assert handle.factorize()
        x, _, _ = handle.solve(handle.rhs())
        # referee: scipy on the oracle-identical H
        cp, ri = handle.h_pattern()
        H = sp.csc_matrix((handle.h_values(), ri, cp), shape=(n, n))
        r = H @ x - handle.rhs()
        assert np.abs(r).max() <= 1e-10 * np.abs(handle.rhs()).max()


def test_lm_solve_re_entry_with_a_grown_jacobian(handle, orc):
    """spb_lm_solve re-validates the pattern at entry and takes the incremental path by itself."""
    import saddlepoint_b200 as spb
    rng = np.random.default_rng(21)
    g = 24
    J0 = _lf(orc, g)
    m, n = J0.shape
    b = np.ones(m)

    class Traits:
        def __init__(self, J, b):
            self.J, self.b = J.tocsc(), b
            self.J.sort_indices()
            self.xSize, self.ESize = J.shape[1], J.shape[0]

        def initial_solution(self):
            return np.ones(self.xSize)

        def objective_jacobian(self, x, compute_jacobian):
            return self.J @ x - self.b, (self.J if compute_jacobian else None)

        def pre_iteration(self, x):
            pass

        def post_iteration(self, x):
            return False

    ls = spb.B200SolverWrapper(handle)
    lm = spb.LMSolver()
    lm.init(ls, Traits(J0, b), spb.DiagonalDamping(0.01), 20)
    lm.solve(False)
    x_first = lm.x.copy()
    # pin variable 7 to its current value: one constraint row appended (IterativeRoundingTraits.h:76-79)
    J1 = _insert_row(J0, m, [7], [1.0])
    b1 = np.concatenate([b, [np.round(x_first[7])]])
    lm.init(ls, Traits(J1, b1), spb.DiagonalDamping(0.01), 20)
    lm.solve(False)
    # LF is linear: the solve ends at the least-squares solution of the GROWN system
    grad = J1.T @ (J1 @ lm.x - b1)
    assert np.abs(grad).max() < 1e-6 * max(1.0, np.abs(J1.T @ b1).max())
    assert handle.pattern_matches(*_csc(J1)[:4])
