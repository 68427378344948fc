"""GPU tier: fused assembly (H, rhs, damping, ||rhs||_inf) through the C ABI vs the oracle.
Bar: pattern bit-exact; values bit-exact (the kernels reproduce the reference's summation
order with separate multiply/add), including explicit zeros, empty columns/rows and inf/NaN."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import random_jacobian

pytestmark = pytest.mark.gpu


def _reference(orc, m, n, colptr, rowidx, vals, E, lam):
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, d = orc.damp_build(J, lam)
    hp, hi, hx = orc.ata(D).csc()
    g = orc.jt_vec_neg(J, E)
    return hp, hi, hx, g, d


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def _check(handle, orc, m, n, colptr, rowidx, vals, E, lam):
    handle.analyze(m, n, colptr, rowidx)
    foo = handle.assemble(vals, E, lam)
    hp, hi, hx, g, d = _reference(orc, m, n, colptr, rowidx, vals, E, lam)
    cp, ri = handle.h_pattern()
    assert (cp == hp).all() and (ri == hi).all(), "H pattern differs"
    hv = handle.h_values()
    assert (_bits(hv) == _bits(hx)).all(), "H values not bit-exact: max rel %g" % np.nanmax(np.abs(hv - hx) / (np.abs(hx) + 1e-300))
    assert (_bits(handle.rhs()) == _bits(g)).all(), "rhs not bit-exact"
    assert (_bits(handle.damping()) == _bits(d)).all(), "damping vector not bit-exact"
    finite = np.abs(g[~np.isnan(g)])
    assert foo == (finite.max() if len(finite) else 0.0)


@pytest.mark.parametrize("seed,m,n,k", [(0, 64, 20, 3), (1, 1000, 400, 4), (2, 5000, 5000, 6), (3, 300, 900, 2)])
def test_random_patterns(handle, orc, seed, m, n, k):
    rng = np.random.default_rng(seed)
    colptr, rowidx, vals = random_jacobian(rng, m, n, k, explicit_zero_frac=0.05)
    _check(handle, orc, m, n, colptr, rowidx, vals, rng.standard_normal(m), 0.01)


def test_empty_columns_and_rows(handle, orc):
    rng = np.random.default_rng(7)
    m, n = 400, 150
    colptr, rowidx, vals = random_jacobian(rng, m, n, 3, empty_cols=[0, 31, 32, 149], empty_rows=[0, 1, 399])
    _check(handle, orc, m, n, colptr, rowidx, vals, rng.standard_normal(m), 0.5)


def test_rosenbrock_c1_with_explicit_zero(handle, orc):
    # config C1: 500 blocks, J triplets incl. the structural zero (benchmark/rosenbrock/main.cpp:57-60)
    nb = 500
    n = m = 2 * nb
    x = np.tile([-100.0, 100.0], nb)
    colptr = (2 * np.arange(n + 1)).astype(np.int32)
    rowidx = (np.repeat(2 * np.arange(nb), 4) + np.tile([0, 1, 0, 1], nb)).astype(np.int32)
    vals = np.empty(4 * nb)
    vals[0::4] = -50.0 * x[0::2]
    vals[1::4] = -1.0
    vals[2::4] = 25.0
    vals[3::4] = 0.0
    E = np.empty(m)
    E[0::2] = 25.0 * (x[1::2] - x[0::2] ** 2)
    E[1::2] = 1.0 - x[0::2]
    _check(handle, orc, m, n, colptr, rowidx, vals, E, 0.01)
    assert handle.h_nnz() == 2000


@pytest.mark.parametrize("g", [8, 33, 100])
def test_lf_workload(handle, orc, g):
    J = orc.lf_matrix(g, 2, 20211)
    colptr, rowidx, vals = J.csc()
    m, n = J.shape
    E = J.scipy() @ np.ones(n) - 1.0
    _check(handle, orc, m, n, colptr, rowidx, vals, E, 0.01)


def test_dense_columns_use_wide_records_and_multi_tile_slots(handle, orc):
    # columns with >128 entries (32-bit records) and slots with >2048 pairs (multi-tile path);
    # this is the shape of the reference's dense linear_function Jacobian (m=1000, n=250 there)
    rng = np.random.default_rng(11)
    m, n = 3000, 5
    A = rng.standard_normal((m, n))
    A[rng.random((m, n)) < 0.1] = 0.0
    S = sp.csc_matrix(A)
    S.sort_indices()
    _check(handle, orc, m, n, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data, rng.standard_normal(m), 0.01)


def test_reference_linear_function_jacobian(handle, orc):
    m, n = 1000, 250
    A = np.vstack([np.eye(n) - 2.0 / m, -2.0 / m * np.ones((m - n, n))])
    colptr = (m * np.arange(n + 1)).astype(np.int32)
    rowidx = np.tile(np.arange(m), n).astype(np.int32)
    vals = np.ascontiguousarray(A.T).ravel()
    _check(handle, orc, m, n, colptr, rowidx, vals, A @ np.ones(n) - 1.0, 0.01)


def test_inf_and_nan_entries_propagate_like_the_reference(handle, orc):
    # the seamless barrier injects inf into J on purpose (SIInitialSolutionTraits.h:163,227-228)
    rng = np.random.default_rng(13)
    m, n = 200, 60
    colptr, rowidx, vals = random_jacobian(rng, m, n, 3)
    vals[5] = np.inf
    vals[17] = -np.inf
    vals[40] = np.nan
    handle.analyze(m, n, colptr, rowidx)
    E = rng.standard_normal(m)
    handle.assemble(vals, E, 0.01)
    hp, hi, hx, g, d = _reference(orc, m, n, colptr, rowidx, vals, E, 0.01)
    hv = handle.h_values()
    assert (np.isnan(hv) == np.isnan(hx)).all()
    ok = ~np.isnan(hx)
    assert (hv[ok] == hx[ok]).all()
    gg = handle.rhs()
    assert (np.isnan(gg) == np.isnan(g)).all() and (gg[~np.isnan(g)] == g[~np.isnan(g)]).all()


def test_lambda_changes_only_the_diagonal(handle, orc):
    rng = np.random.default_rng(17)
    m, n = 500, 200
    colptr, rowidx, vals = random_jacobian(rng, m, n, 4)
    E = rng.standard_normal(m)
    handle.analyze(m, n, colptr, rowidx)
    handle.assemble(vals, E, 0.01)
    H1 = handle.h_scipy()
    handle.assemble(vals, E, 10.0)
    H2 = handle.h_scipy()
    Dm = (H2 - H1).tocsc()
    Dm.eliminate_zeros()
    assert (Dm - sp.diags(Dm.diagonal())).nnz == 0
    _check(handle, orc, m, n, colptr, rowidx, vals, E, 10.0)


def test_device_pointers_and_reanalysis(handle, orc):
    import torch
    rng = np.random.default_rng(19)
    m, n = 800, 300
    colptr, rowidx, vals = random_jacobian(rng, m, n, 4)
    E = rng.standard_normal(m)
    handle.analyze(m, n, colptr, rowidx)
    assert handle.pattern_matches(m, n, colptr, rowidx)
    foo = handle.assemble(torch.tensor(vals, device="cuda"), torch.tensor(E, device="cuda"), 0.01)
    hp, hi, hx, g, d = _reference(orc, m, n, colptr, rowidx, vals, E, 0.01)
    assert (handle.h_values() == hx).all() and foo == np.abs(g).max()
    # a grown pattern (one more residual row) is detected and re-analysed
    colptr2, rowidx2, vals2 = random_jacobian(rng, m + 1, n, 4)
    assert not handle.pattern_matches(m + 1, n, colptr2, rowidx2)
    _check(handle, orc, m + 1, n, colptr2, rowidx2, vals2, rng.standard_normal(m + 1), 0.01)


def test_call_order_errors(spb):
    h = spb.Handle(0)
    with pytest.raises(spb.SpbError):
        h.assemble(np.zeros(3), np.zeros(3), 0.01)
    h.close()


def test_degenerate_shapes(spb, orc):
    # structurally empty Jacobian: H is the (zero) structural diagonal, the gradient is zero, and the
    # factorisation reports failure like the reference's LLT on a singular matrix (quirk F)
    for m, n in [(5, 3), (0, 4), (1, 1)]:
        h = spb.Handle(0)
        colptr = np.zeros(n + 1, np.int32)
        rowidx = np.zeros(0, np.int32)
        h.analyze(m, n, colptr, rowidx)
        foo = h.assemble(np.zeros(0), np.zeros(m), 0.01)
        assert foo == 0.0
        cp, ri = h.h_pattern()
        assert (cp == np.arange(n + 1)).all() and (ri == np.arange(n)).all()
        assert (h.h_values() == 0).all() and (h.rhs() == 0).all()
        for solver in (spb.SOLVER_CHOLESKY, spb.SOLVER_PCG_JACOBI):
            h.set_solver(solver)
            ok = h.factorize()
            st = spb.capi.SPB_NOT_SPD if not ok else h.solve()[2]
            assert st == spb.capi.SPB_NOT_SPD or (solver == spb.SOLVER_PCG_JACOBI and st == 0)  # PCG with rhs == 0 has nothing to iterate
        h.close()


def test_single_dense_row_and_single_column(handle, orc):
    rng = np.random.default_rng(23)
    # one residual row touching every unknown: H is dense rank-1 + damping
    n = 40
    colptr = np.arange(n + 1, dtype=np.int32)
    rowidx = np.zeros(n, np.int32)
    _check(handle, orc, 1, n, colptr, rowidx, rng.standard_normal(n), rng.standard_normal(1), 0.3)
    # one unknown, many residuals: H is 1 x 1
    m = 500
    _check(handle, orc, m, 1, np.array([0, m], np.int32), np.arange(m, dtype=np.int32), rng.standard_normal(m), rng.standard_normal(m), 0.01)


def test_streamed_host_assembly_equals_device_assembly(spb):
    """spb_assemble from PINNED host memory uploads J in parts on a copy stream and assembles every part as soon
    as its values are on the device; the result must be the same bits as the assembly from device buffers."""
    import torch
    h = spb.Handle(0)
    g = 384  # nnzJ = 2.36 M: above the streaming threshold
    prob = spb.BuiltinProblem(h, "lf", g=g, rows_mult=2, seed=20211)
    m, n, colptr, rowidx = prob.pattern()
    T = prob.c_traits()
    xd = torch.ones(n, dtype=torch.float64, device="cuda")
    Ed = torch.empty(m, dtype=torch.float64, device="cuda")
    Jd = torch.empty(int(colptr[-1]), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    T.objective_jacobian(T.user, xd.data_ptr(), Ed.data_ptr(), Jd.data_ptr(), None)
    torch.cuda.synchronize()
    h.analyze(m, n, colptr, rowidx)
    foo_dev = h.assemble(Jd, Ed, 0.01)
    H_dev, g_dev, d_dev = h.h_values(), h.rhs(), h.damping()
    Jh, Eh = Jd.cpu().pin_memory(), Ed.cpu().pin_memory()
    for _ in range(2):  # twice: the second call reuses the staging buffers behind the first call's kernels
        foo_host = h.assemble(Jh.numpy(), Eh.numpy(), 0.01)
        assert foo_host == foo_dev
        assert np.array_equal(h.h_values(), H_dev) and np.array_equal(h.rhs(), g_dev) and np.array_equal(h.damping(), d_dev)
    # pageable host memory takes the plain path and gives the same result
    foo_pg = h.assemble(Jd.cpu().numpy(), Ed.cpu().numpy(), 0.01)
    assert foo_pg == foo_dev and np.array_equal(h.h_values(), H_dev)
    prob.close()
    h.close()


def test_stream_form_is_the_default_and_general_kernels_agree(handle, orc):
    """The packed stream form (assemble_pairs_stream_kernel) runs wherever it applies; the general kernels take over for a
    misaligned device pointer and for long J columns -- every form reproduces the oracle's bits (LMSolver.h:136)."""
    import torch
    J = orc.lf_matrix(48, 2, 20211)
    colptr, rowidx, vals = J.csc()
    m, n = J.shape
    rng = np.random.default_rng(3)
    E = rng.standard_normal(m)
    _check(handle, orc, m, n, colptr, rowidx, vals, E, 0.01)
    assert handle.assembly_form() == 3
    hx = handle.h_values().copy()
    # same problem, device pointer shifted by 8 bytes: no 16-byte aligned bulk copies -> general kernel, same bits
    buf = torch.empty(len(vals) + 1, dtype=torch.float64, device="cuda")
    buf[1:] = torch.from_numpy(vals)
    handle.assemble(buf[1:], torch.from_numpy(E).cuda(), 0.01)
    assert handle.assembly_form() in (1, 2)
    assert (_bits(handle.h_values()) == _bits(hx)).all()
    # J columns longer than 128 entries: 32-bit pair records, general kernel
    m2, n2 = 400, 6
    A = (rng.random((m2, n2)) < 0.7) * rng.standard_normal((m2, n2))
    S = sp.csc_matrix(A)
    S.sort_indices()
    _check(handle, orc, m2, n2, S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data, rng.standard_normal(m2), 0.3)
    assert handle.assembly_form() in (1, 2)


@pytest.mark.parametrize("seed,m,n,k", [(10, 2000, 700, 4), (11, 300, 900, 2), (12, 6000, 6000, 5)])
def test_stream_form_forced_on_scattered_patterns(spb, orc, seed, m, n, k):
    """Scattered patterns are left to the general kernels by default (short bulk-copy runs); forced, the stream form
    reproduces the oracle's bits on them too (explicit zeros, empty columns, odd value counts)."""
    import os
    rng = np.random.default_rng(seed)
    colptr, rowidx, vals = random_jacobian(rng, m, n, k, explicit_zero_frac=0.05, empty_cols=[0, n - 1] if seed == 11 else ())
    os.environ["SPB_ASM_V2_MIN_RUN"] = "0"
    try:
        h = spb.Handle(0)
        _check(h, orc, m, n, colptr, rowidx, vals, rng.standard_normal(m), 0.02)
        assert h.assembly_form() == 3
        h.close()
    finally:
        del os.environ["SPB_ASM_V2_MIN_RUN"]


def test_stream_form_odd_value_count_tail(spb, orc):
    """An odd number of J values: the last bulk-copy window would end one element past the value array, so the producer
    shortens it and fetches the last element by hand (ChunkHdr2::tail_src).  The device array is exactly nnz long."""
    J = orc.lf_matrix(40, 2, 20211)
    colptr, rowidx, vals = J.csc()
    m, n = J.shape
    colptr = colptr.copy()
    colptr[-1] -= 1
    rowidx, vals = rowidx[:-1].copy(), vals[:-1].copy()
    assert colptr[-1] % 2 == 1
    rng = np.random.default_rng(5)
    h = spb.Handle(0)
    _check(h, orc, m, n, colptr, rowidx, vals, rng.standard_normal(m), 0.01)
    assert h.assembly_form() == 3
    h.close()


def test_stream_form_inf_nan_and_explicit_zeros(spb, orc):
    """inf / NaN / signed zeros through the stream form (LF pattern, so the form applies by itself): the first product
    is assigned, not added to 0, and every later one is a separately rounded multiply and add -- NaN positions and all
    other bits equal the oracle's (SIInitialSolutionTraits.h:163,227-228 injects inf on purpose)."""
    J = orc.lf_matrix(32, 2, 20211)
    colptr, rowidx, vals = J.csc()
    m, n = J.shape
    vals = vals.copy()
    rng = np.random.default_rng(21)
    pick = rng.choice(len(vals), 40, replace=False)
    vals[pick[:10]] = np.inf
    vals[pick[10:20]] = -np.inf
    vals[pick[20:25]] = np.nan
    vals[pick[25:33]] = 0.0
    vals[pick[33:]] = -0.0
    E = rng.standard_normal(m)
    h = spb.Handle(0)
    h.analyze(m, n, colptr, rowidx)
    h.assemble(vals, E, 0.01)
    assert h.assembly_form() == 3
    hp, hi, hx, g, d = _reference(orc, m, n, colptr, rowidx, vals, E, 0.01)
    hv = h.h_values()
    assert (np.isnan(hv) == np.isnan(hx)).all()
    ok = ~np.isnan(hx)
    assert (_bits(hv[ok]) == _bits(hx[ok])).all()   # including the sign of zeros
    h.close()


def test_stream_form_randomised_sweep(spb, orc):
    """Forced stream form over a sweep of small random shapes (single columns, one row, more columns than rows, odd
    counts): every case bit-exact against the oracle."""
    import os
    os.environ["SPB_ASM_V2_MIN_RUN"] = "0"
    try:
        rng = np.random.default_rng(99)
        h = spb.Handle(0)
        for case in range(24):
            m = int(rng.integers(1, 400))
            n = int(rng.integers(1, 300))
            k = int(rng.integers(1, 5))
            colptr, rowidx, vals = random_jacobian(rng, m, n, k, explicit_zero_frac=0.1)
            _check(h, orc, m, n, colptr, rowidx, vals, rng.standard_normal(m), float(rng.uniform(0.001, 2.0)))
        h.close()
    finally:
        del os.environ["SPB_ASM_V2_MIN_RUN"]
