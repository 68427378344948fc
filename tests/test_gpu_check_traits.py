"""GPU tier: check_traits (check_traits.h:19-83) with column-coloured finite differences vs the
reference's one-unknown-at-a-time procedure restated below (numpy, small sizes)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import random_jacobian

pytestmark = pytest.mark.gpu


class SineTraits:
    """E_r = sin(sum_j a_rj x_j): Jacobian a_rj cos(.), fixed random sparse pattern."""

    def __init__(self, rng, m, n, k, wrong_at=None, hidden=None):
        colptr, rowidx, vals = random_jacobian(rng, m, n, k, cover_cols=True)
        self.A = sp.csc_matrix((vals, rowidx, colptr), shape=(m, n))
        self.Ar = self.A.tocsr()
        self.xSize, self.ESize = n, m
        self.colptr, self.rowidx = colptr, rowidx
        self.wrong_at = wrong_at  # CSC position whose analytic value is off by 0.5
        self.hidden = hidden      # (row, col): a dependence the declared pattern does not list
        self.x0 = rng.standard_normal(n)
        self.evaluations = 0

    def initial_solution(self):
        return self.x0

    def objective_jacobian(self, x, computeJacobian):
        self.evaluations += 1
        arg = self.Ar @ x
        E = np.sin(arg)
        if self.hidden is not None:
            E[self.hidden[0]] += 0.3 * x[self.hidden[1]]
        if not computeJacobian:
            return E, None
        J = self.A.data * np.cos(arg)[self.A.indices]
        if self.wrong_at is not None:
            J = J.copy()
            J[self.wrong_at] += 0.5
        return E, J


def reference_check(ST, x):
    """check_traits.h:47-76 restated: one unknown at a time, dense FE columns, diff over the union."""
    n, m = ST.xSize, ST.ESize
    _, Jv = ST.objective_jacobian(x, True)
    J = sp.csc_matrix((Jv, ST.rowidx, ST.colptr), shape=(m, n)).toarray()
    patt = sp.csc_matrix((np.ones_like(Jv), ST.rowidx, ST.colptr), shape=(m, n)).toarray() > 0
    FE = np.zeros((m, n))
    for i in range(n):
        vh = np.zeros(n)
        vh[i] = 10e-5
        Ep, _ = ST.objective_jacobian(x + vh, False)
        Em, _ = ST.objective_jacobian(x - vh, False)
        g = (Ep - Em) / (2 * 10e-5)
        FE[:, i] = np.where(np.abs(g) > 10e-7, g, 0.0)
    D = np.abs(FE - J)
    union = patt | (FE != 0)
    best, loc = -32767.0, (-1, -1)
    for j in range(n):  # column-major, strict > keeps the first maximum
        for r in np.nonzero(union[:, j])[0]:
            if best < D[r, j]:
                best, loc = D[r, j], (r, j)
    return FE, best, loc, int((D[union] > 10e-7).sum())


def test_builtin_rosenbrock_jacobian_is_consistent(spb, handle):
    prob = spb.BuiltinProblem(handle, "rosenbrock", nblocks=300)
    x = np.tile([-100.0, 100.0], 300)
    out = spb.check_traits(handle, prob, x, verbose=False)
    assert out["ncolors"] == 2 and out["evaluations"] == 5
    assert out["discrepancies"] == 0 and out["outside_count"] == 0
    assert 0 <= out["max_diff"] < 10e-7


def test_lf_jacobian_is_consistent(spb, handle):
    prob = spb.BuiltinProblem(handle, "lf", g=24, rows_mult=2, seed=20211)
    x = np.random.default_rng(3).standard_normal(prob.xSize)
    out = spb.check_traits(handle, prob, x, verbose=False)
    assert out["evaluations"] == 1 + 2 * out["ncolors"] and out["ncolors"] < 64
    assert out["discrepancies"] == 0 and out["outside_count"] == 0


@pytest.mark.parametrize("wrong", [False, True])
def test_coloured_differences_equal_the_one_at_a_time_procedure(spb, handle, wrong, capsys):
    rng = np.random.default_rng(11)
    ST = SineTraits(rng, 90, 40, 3)
    if wrong:
        ST.wrong_at = 37
    x = rng.standard_normal(40)
    FE, best, loc, ndisc = reference_check(ST, x)
    ST.evaluations = 0
    out = spb.check_traits(handle, ST, x, return_fe=True)
    assert "Maximum gradient difference" in capsys.readouterr().out
    # every row answers to one perturbed column only: the compressed differences are the same bits
    fe = np.where(np.abs(out["fe"]) > 10e-7, out["fe"], 0.0)
    assert np.array_equal(fe, FE[ST.rowidx, np.repeat(np.arange(40), np.diff(ST.colptr))])
    assert out["max_diff"] == best and (out["max_row"], out["max_col"]) == loc
    assert out["discrepancies"] == ndisc and (ndisc == 1) == wrong
    assert out["outside_count"] == 0
    assert out["evaluations"] == 1 + 2 * out["ncolors"] < 2 * 40  # the point of the exercise
    assert ST.evaluations >= out["evaluations"]  # (+ the pattern probe of the python wrapper)


def test_dependence_outside_the_declared_pattern_is_reported(spb, handle):
    rng = np.random.default_rng(12)
    ST = SineTraits(rng, 60, 30, 3)
    dense = sp.csc_matrix((np.ones(len(ST.rowidx)), ST.rowidx, ST.colptr), shape=(60, 30)).toarray()
    r, c = map(int, np.argwhere(dense == 0)[5])
    ST.hidden = (r, c)
    out = spb.check_traits(handle, ST, rng.standard_normal(30), verbose=False)
    assert out["outside_count"] >= 1 and abs(out["outside_max"] - 0.3) < 1e-6
