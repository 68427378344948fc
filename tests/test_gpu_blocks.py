"""GPU tier: the device-resident block problem (spb_problem_blocks: constant sparse block + injectivity
barrier, the seamless-integration Jacobian structure) -- same bits as the numpy traits,
and the LM loop through it against the oracle driven by those numpy traits."""
import ctypes as C

import numpy as np
import pytest

import c3_blocks

pytestmark = pytest.mark.gpu


def _device_eval(spb, prob, x, with_j):
    import torch
    T = prob.c_traits()
    n, m = prob.xSize, prob.ESize
    m_, n_, colptr, rowidx = prob.pattern()
    xd = torch.tensor(x, dtype=torch.float64, device="cuda")
    Ed = torch.empty(m, dtype=torch.float64, device="cuda")
    Jd = torch.empty(int(colptr[-1]), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    T.objective_jacobian(T.user, xd.data_ptr(), Ed.data_ptr(), Jd.data_ptr() if with_j else None, None)
    torch.cuda.synchronize()
    return Ed.cpu().numpy(), Jd.cpu().numpy()


@pytest.mark.parametrize("flip", [None, 3])
def test_device_block_traits_equal_numpy_traits_bitwise(spb, handle, flip):
    P = c3_blocks.build(7, 6, flip_face=flip)
    host = c3_blocks.BlocksHostTraits(P)
    prob = spb.BuiltinProblem(handle, "blocks", **{k: P[k] for k in ("A", "b", "bar_idx", "bar_vol", "s", "w_bar", "x0")})
    m, n, colptr, rowidx = prob.pattern()
    assert (m, n) == (host.ESize, host.xSize)
    assert np.array_equal(colptr, host.colptr) and np.array_equal(rowidx, host.rowidx)
    rng = np.random.default_rng(5)
    for x in (P["x0"], P["x0"] * (1.0 + 0.05 * rng.standard_normal(n))):
        E, J = _device_eval(spb, prob, x, True)
        Eh, Jh = host.objective_jacobian(x, True)
        assert np.array_equal(E, Eh, equal_nan=True)
        assert np.array_equal(J, Jh, equal_nan=True)
    if flip is not None:
        assert np.isinf(Eh).any()
    prob.close()


def test_lm_through_device_block_traits_matches_oracle(spb, orc):
    P = c3_blocks.build(7, 6)
    host = c3_blocks.BlocksHostTraits(P)
    ls = spb.B200SolverWrapper(solver=spb.SOLVER_CHOLESKY)
    prob = spb.BuiltinProblem(ls.handle, "blocks", **{k: P[k] for k in ("A", "b", "bar_idx", "bar_vol", "s", "w_bar", "x0")})
    lm = spb.LMSolver()
    lm.init(ls, prob, spb.DiagonalDamping(0.01), 30)
    ret = lm.solve(True)
    ref = orc.lm_solve("callback", x0=host.initial_solution(), pattern=(host.ESize, host.xSize, host.colptr, host.rowidx),
                       callback=lambda x, cj: host.objective_jacobian(x, cj), maxIterations=30, keep_first=True)
    assert ret == ref.ret and lm.exit_path == ref.exit_path
    assert lm.currIter == ref.currIter and lm.factorizations == ref.factorizations
    assert (lm.trace["accepted"] == ref.trace["accepted"]).all()
    np.testing.assert_allclose(lm.trace["lambda"], ref.trace["lambda"], rtol=0)
    assert np.linalg.norm(lm.x - ref.x) <= 1e-9 * np.linalg.norm(ref.x)   # kappa ~ (1e4/1e-4)^2, as in test_gpu_c3
    assert abs(lm.energy - ref.energy) <= 1e-9 * max(abs(ref.energy), 1e-16 * ref.trace["prevEnergy2"][0])
    # gradient check of the device traits with the coloured finite differences
    chk = spb.check_traits(ls.handle, prob, lm.x, verbose=False)
    assert chk["outside_count"] == 0 and chk["evaluations"] == 1 + 2 * chk["ncolors"]
    prob.close()
    ls.handle.close()


def test_blocks_rejects_bad_input(spb, handle):
    P = c3_blocks.build(4, 4)
    bad = dict({k: P[k] for k in ("A", "b", "bar_idx", "bar_vol", "s", "w_bar", "x0")})
    bad["bar_idx"] = P["bar_idx"].copy()
    bad["bar_idx"][0, 1] = bad["bar_idx"][0, 0]
    with pytest.raises(spb.SpbError):
        spb.BuiltinProblem(handle, "blocks", **bad)
