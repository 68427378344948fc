"""CPU tier: the C-ABI library loads, exports every symbol the header declares, refuses to run
without a GPU (no CPU fallback), and its host-only symbolic phase matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import random_jacobian

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(spb):
    L = spb.capi.load()
    header = open(os.path.join(ROOT, "include", "saddlepoint_b200.h")).read()
    declared = set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", header))
    declared -= {"spb_handle"}
    bound = {name for name, _, _ in spb.capi.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.spb_version()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu(spb):
    L = spb.capi.load()
    h = C.c_void_p()
    assert L.spb_create(0, None, C.byref(h)) == spb.capi.SPB_ERR_CUDA and not h.value
    with pytest.raises(spb.SpbError):
        spb.Handle(0)


def test_product_does_not_reference_oracle():
    # the product tree must not import/link anything under oracle/
    for base, _, files in os.walk(os.path.join(ROOT, "saddlepoint_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "sp_oracle" not in src and "oracle/" not in src, os.path.join(base, f)


def _h_pattern(L, m, n, colptr, rowidx):
    nnz = C.c_int64(0)
    assert L.spb_symbolic_h_pattern(m, n, colptr.ctypes.data, rowidx.ctypes.data, C.byref(nnz), None, None) == 0
    hp = np.empty(n + 1, np.int32)
    hi = np.empty(nnz.value, np.int32)
    assert L.spb_symbolic_h_pattern(m, n, colptr.ctypes.data, rowidx.ctypes.data, C.byref(nnz), hp.ctypes.data, hi.ctypes.data) == 0
    return hp, hi


@pytest.mark.parametrize("case", ["random", "empty_cols_rows", "rosenbrock", "lf", "dense_cols"])
def test_symbolic_pattern_bit_exact_vs_oracle(spb, orc, case):
    L = spb.capi.load()
    rng = np.random.default_rng(5)
    if case == "random":
        m, n = 500, 200
        colptr, rowidx, vals = random_jacobian(rng, m, n, 4, explicit_zero_frac=0.2)
    elif case == "empty_cols_rows":
        m, n = 200, 90
        colptr, rowidx, vals = random_jacobian(rng, m, n, 3, empty_cols=[0, 17, 89], empty_rows=[0, 5, 199])
    elif case == "rosenbrock":
        nb = 50
        m = n = 2 * nb
        colptr = (2 * np.arange(n + 1)).astype(np.int32)
        rowidx = np.repeat(2 * np.arange(nb), 4).astype(np.int32) + np.tile([0, 1, 0, 1], nb).astype(np.int32)
        vals = rng.standard_normal(4 * nb)
        vals[3::4] = 0.0  # the reference's explicit structural zero
    elif case == "lf":
        J = orc.lf_matrix(20, 2, 20211)
        colptr, rowidx, vals = J.csc()
        m, n = J.shape
    else:  # columns longer than 128 entries => 32-bit pair records
        m, n = 400, 6
        A = (rng.random((m, n)) < 0.7) * rng.standard_normal((m, n))
        import scipy.sparse as sp
        S = sp.csc_matrix(A)
        S.sort_indices()
        colptr, rowidx, vals = S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, _ = orc.damp_build(J, 0.01)
    hp_ref, hi_ref, _ = orc.ata(D).csc()
    hp, hi = _h_pattern(L, m, n, colptr, rowidx)
    assert (hp == hp_ref).all() and (hi == hi_ref).all()
    stats = np.zeros(8, np.int64)
    assert L.spb_symbolic_plan_check(m, n, colptr.ctypes.data, rowidx.ctypes.data, stats.ctypes.data) == 0
    assert stats[7] == 0, "assembly plan invariants violated"
    assert stats[5] == len(hi_ref)
    assert stats[3] == (32 if case == "dense_cols" else 16)
    # every strictly-lower slot has at least one contributing pair
    assert stats[1] >= stats[0]


def test_symbolic_rejects_bad_patterns(spb):
    L = spb.capi.load()
    nnz = C.c_int64(0)
    colptr = np.array([0, 2], np.int32)
    rowidx = np.array([1, 0], np.int32)  # not ascending
    assert L.spb_symbolic_h_pattern(2, 1, colptr.ctypes.data, rowidx.ctypes.data, C.byref(nnz), None, None) == spb.capi.SPB_ERR_ARG
    rowidx = np.array([0, 5], np.int32)  # out of range
    assert L.spb_symbolic_h_pattern(2, 1, colptr.ctypes.data, rowidx.ctypes.data, C.byref(nnz), None, None) == spb.capi.SPB_ERR_ARG


def test_symbolic_results_do_not_depend_on_the_thread_count():
    """The threaded host phases (H pattern, pair stream, chunk packing; nested dissection over subtrees)
    must produce exactly what the single-threaded run produces.  The thread count is latched per process,
    so every count runs in its own interpreter and reports a digest."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from oracle import sp_oracle as orc
orc.build()
import saddlepoint_b200 as spb
L = spb.capi.load()
g = 224                                    # 50 176 unknowns: above both parallel thresholds
A = orc.lf_matrix(g, 2, 20211)
p, i, x = A.csc()
n, m = g * g, 2 * g * g
hsh = hashlib.sha256()
nnz = C.c_int64(0)
hp = np.zeros(n + 1, np.int32)
assert L.spb_symbolic_h_pattern(m, n, p.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), C.byref(nnz), hp.ctypes.data_as(C.c_void_p), None) == 0
hi = np.zeros(nnz.value, np.int32)
assert L.spb_symbolic_h_pattern(m, n, p.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), C.byref(nnz), hp.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p)) == 0
hsh.update(hp.tobytes()); hsh.update(hi.tobytes())
st = (C.c_int64 * 8)()
assert L.spb_symbolic_plan_check(m, n, p.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), st) == 0 and st[7] == 0
hsh.update(bytes(str(list(st)), "ascii"))
perm = np.zeros(n, np.int32)
assert L.spb_symbolic_chol_stats(m, n, p.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), st, perm.ctypes.data_as(C.c_void_p)) == 0 and st[7] == 0
hsh.update(bytes(str(list(st)), "ascii")); hsh.update(perm.tobytes())
print("DIGEST", hsh.hexdigest())
''' % ROOT
    digests = []
    for threads in ("1", "5"):
        env = dict(os.environ, SPB_HOST_THREADS=threads)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
        digests.append(re.search(r"DIGEST (\w+)", out.stdout).group(1))
    assert digests[0] == digests[1]


@pytest.mark.parametrize("case", ["random", "empty_cols_rows", "rosenbrock", "lf", "lf_odd_tail", "dense_cols"])
def test_pairs_stream_form_bit_exact_vs_oracle(spb, orc, case):
    """The packed 'stream' form of the pair assembly (pairs_v2.cpp), executed on the host in the kernel's order,
    reproduces every off-diagonal entry of the oracle's dampJ^T*dampJ bit for bit (LMSolver.h:136)."""
    L = spb.capi.load()
    rng = np.random.default_rng(11)
    if case == "random":
        m, n = 700, 300
        colptr, rowidx, vals = random_jacobian(rng, m, n, 4, explicit_zero_frac=0.2)
    elif case == "empty_cols_rows":
        m, n = 200, 90
        colptr, rowidx, vals = random_jacobian(rng, m, n, 3, empty_cols=[0, 17, 89], empty_rows=[0, 5, 199])
    elif case == "rosenbrock":
        nb = 50
        m = n = 2 * nb
        colptr = (2 * np.arange(n + 1)).astype(np.int32)
        rowidx = np.repeat(2 * np.arange(nb), 4).astype(np.int32) + np.tile([0, 1, 0, 1], nb).astype(np.int32)
        vals = rng.standard_normal(4 * nb)
        vals[3::4] = 0.0
    elif case in ("lf", "lf_odd_tail"):
        J = orc.lf_matrix(40, 2, 20211)
        colptr, rowidx, vals = J.csc()
        m, n = J.shape
        if case == "lf_odd_tail":
            # drop the last entry: an odd number of values, so the last bulk-copy window would end past the array
            colptr = colptr.copy()
            colptr[-1] -= 1
            rowidx, vals = rowidx[:-1].copy(), vals[:-1].copy()
            assert colptr[-1] % 2 == 1
    else:
        m, n = 400, 6
        A = (rng.random((m, n)) < 0.7) * rng.standard_normal((m, n))
        import scipy.sparse as sp
        S = sp.csc_matrix(A)
        S.sort_indices()
        colptr, rowidx, vals = S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data
    vals = np.ascontiguousarray(vals, np.float64)
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, _ = orc.damp_build(J, 0.01)
    hp_ref, hi_ref, hx_ref = orc.ata(D).csc()
    stats = np.zeros(8, np.int64)
    out = np.zeros(len(hi_ref), np.float64)
    # scattered patterns have short bulk-copy runs and are left to the general kernels by default: force the form
    os.environ["SPB_ASM_V2_MIN_RUN"] = "0"
    try:
        assert L.spb_symbolic_pairs_v2(m, n, colptr.ctypes.data, rowidx.ctypes.data, vals.ctypes.data, out.ctypes.data, stats.ctypes.data) == 0
    finally:
        del os.environ["SPB_ASM_V2_MIN_RUN"]
    if case == "dense_cols":
        assert stats[0] == 0  # 32-bit pair records: the general kernels stay in charge
        return
    assert stats[0] == 1, "stream form not applicable"
    cols = np.repeat(np.arange(n), np.diff(hp_ref))
    off = hi_ref != cols
    assert (out[off].view(np.int64) == hx_ref[off].view(np.int64)).all()
    assert (out[~off] == 0).all()
    assert stats[4] <= 3072 * 8 and stats[5] % 16 == 0
    if case == "random":  # default policy: short runs -> not applicable
        st2 = np.zeros(8, np.int64)
        assert L.spb_symbolic_pairs_v2(m, n, colptr.ctypes.data, rowidx.ctypes.data, vals.ctypes.data, out.ctypes.data, st2.ctypes.data) == 0
        assert st2[0] == 0


def test_cholesky_ordering_condenses_satellite_groups(spb):
    """Static condensation (chol_symbolic.cpp): on the seamless-integration structure the 8 field unknowns of every face are
    eliminated first, one leaf supernode per face, and the dissection runs on the vertex unknowns only -- far less fill than
    level-structure dissection of the full graph; the symbolic invariants hold either way."""
    import scipy.sparse as sp
    sys_path_tests = os.path.join(ROOT, "tests")
    import sys
    if sys_path_tests not in sys.path:
        sys.path.insert(0, sys_path_tests)
    import c3_blocks
    L = spb.capi.load()
    P = c3_blocks.build(24, 24)
    A = P["A"]
    n = A.shape[1]
    bi = P["bar_idx"]
    nb = len(bi)
    B = sp.csr_matrix((np.ones(4 * nb), (np.repeat(np.arange(nb), 4), bi.ravel())), shape=(nb, n))
    J = sp.vstack([sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape), B]).tocsc()
    J.sort_indices()
    m = J.shape[0]
    colptr, rowidx = J.indptr.astype(np.int32), J.indices.astype(np.int32)

    def stats(deg):
        os.environ["SPB_CHOL_CONDENSE_DEG"] = str(deg)
        try:
            st = np.zeros(8, np.int64)
            perm = np.zeros(n, np.int32)
            assert L.spb_symbolic_chol_stats(m, n, colptr.ctypes.data, rowidx.ctypes.data, st.ctypes.data, perm.ctypes.data) == 0
            assert st[7] == 0, "structural violations"
            assert sorted(perm.tolist()) == list(range(n))
            return st
        finally:
            del os.environ["SPB_CHOL_CONDENSE_DEG"]

    plain, cond = stats(0), stats(16)
    assert cond[0] >= 0.8 * P["nF"]      # about one leaf supernode per face (boundary faces may merge or stay)
    assert cond[2] * 4 < plain[2]        # nnz(L)
    assert cond[5] * 20 < plain[5]       # flops
