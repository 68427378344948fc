"""GPU tier: subdomain (halo) mode of the one-large-problem case.  On one GPU: the local problems of a
column partition are exactly the rows of the global LF problem (same bits), and the sum of the local
normal matrices restricted to their owners equals the global one.  With >= 2 GPUs a torchrun job runs the
real path (ncclSend/ncclRecv halo exchanges) against the single-GPU solve."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _eval(prob, x):
    import torch
    T = prob.c_traits()
    m, n, colptr, rowidx = prob.pattern()
    xd = torch.tensor(x, dtype=torch.float64, device="cuda")
    Ed = torch.empty(m, dtype=torch.float64, device="cuda")
    Jd = torch.empty(int(colptr[-1]), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    T.objective_jacobian(T.user, xd.data_ptr(), Ed.data_ptr(), Jd.data_ptr(), None)
    torch.cuda.synchronize()
    return Ed.cpu().numpy(), sp.csc_matrix((Jd.cpu().numpy(), rowidx, colptr), shape=(m, n))


def test_local_problems_are_the_rows_of_the_global_problem(spb, handle):
    from saddlepoint_b200 import dist as sd
    g, world = 16, 3
    n, halo = g * g, 2 * g
    rng = np.random.default_rng(8)
    xg = rng.standard_normal(n)
    glob = spb.BuiltinProblem(handle, "lf", g=g, rows_mult=2, seed=20211)
    Eg, Jg = _eval(glob, xg)
    Jg = Jg.tocsr()
    Hsum = sp.csr_matrix((n, n))
    for rank in range(world):
        cb, ce = sd.column_partition(n, world, rank, halo)
        own = ce - cb
        loc = spb.BuiltinProblem(handle, "lf_cols", g=g, rows_mult=2, seed=20211, col_begin=cb, col_end=ce)
        assert loc.xSize == own + 2 * halo and loc.ESize == 2 * own
        gidx = np.arange(cb - halo, ce + halo) % n          # global index of every local unknown
        El, Jl = _eval(loc, xg[gidx])
        rows_g = np.concatenate([np.arange(cb, ce), n + np.arange(cb, ce)])   # local row k*own + (b-cb)  <->  global k*n + b
        assert np.array_equal(El, Eg[rows_g])                # same bits: same entries, same summation order
        # local Jacobian scattered to global columns equals the global rows
        P = sp.csr_matrix((np.ones(len(gidx)), (np.arange(len(gidx)), gidx)), shape=(len(gidx), n))
        assert abs(Jl.tocsr() @ P - Jg[rows_g]).max() == 0.0
        Hsum = Hsum + (P.T @ (Jl.T @ Jl) @ P)
        loc.close()
    Hglob = (Jg.T @ Jg).tocsr()
    assert abs(Hsum - Hglob).max() <= 1e-12 * abs(Hglob).max()   # H = sum of the local normal matrices
    glob.close()


def test_halo_mode_needs_a_communicator(spb):
    h = spb.Handle(0)
    with pytest.raises(spb.SpbError):
        h.ring_halo(5, 25, 5, 1, 1)   # other ranks as neighbours, but no communicator
    with pytest.raises(spb.SpbError):
        h.ring_halo(5, 12, 5, 0, 0)   # fewer than 2*halo owned unknowns
    h.close()


@pytest.fixture
def peer_mode(request, monkeypatch):
    """1: the fused persistent kernel with peer-memory windows (pcg_peer.cu; the default); 0: the launch-per-phase
    path with NCCL / device-copy exchanges (pcg_solve_halo)."""
    monkeypatch.setenv("SPB_HALO_PEER", str(request.param))
    return request.param


@pytest.mark.parametrize("peer_mode", [1, 0], indirect=True)
@pytest.mark.parametrize("g", [24, 64])
def test_self_ring_halo_mode_matches_the_plain_solve(spb, orc, g, peer_mode):
    """One rank that is its own neighbour on the ring: the whole halo-mode code path (local problem with ghost
    copies, halo sums in the assembly, halo exchanges in the PCG, LM loop on local vectors).  peer_mode 1 runs the
    fused kernel against the rank's own window (band slabs, flags and dot-product slots all go through the same
    loads/stores as over NVLink); peer_mode 0 the NCCL path with device copies standing in for ncclSend/ncclRecv.
    The result must agree with the ordinary single-handle solve."""
    n, halo = g * g, 2 * g
    h = spb.Handle(0)
    loc = spb.BuiltinProblem(h, "lf_cols", g=g, rows_mult=2, seed=20211, col_begin=0, col_end=n)
    assert loc.xSize == n + 2 * halo
    h.ring_halo(halo, halo + n, halo, 0, 0)
    lm = spb.LMSolver()
    lm.init(spb.B200SolverWrapper(h, spb.SOLVER_PCG_JACOBI, 1e-13, 10000), loc, spb.DiagonalDamping(0.01), 100)
    ret = lm.solve(True, dedupe_evaluations=True)
    x = lm.x
    # ghosts are copies of their owners
    assert np.array_equal(x[:halo], x[halo + n - halo:halo + n]) and np.array_equal(x[halo + n:], x[halo:2 * halo])
    ref = orc.lm_solve("lf", iparam0=g, iparam1=2, seed=20211, maxIterations=100)
    assert ret == ref.ret and lm.exit_path == ref.exit_path
    assert lm.currIter == ref.currIter and lm.factorizations == ref.factorizations
    xo = x[halo:halo + n]
    assert np.linalg.norm(xo - ref.x) <= 1e-10 * np.linalg.norm(ref.x)
    assert abs(lm.energy - ref.energy) <= 1e-10 * max(abs(ref.energy), 1e-16 * ref.trace["prevEnergy2"][0])
    assert lm.trace["foo"][0] == ref.trace["foo"][0] or abs(lm.trace["foo"][0] - ref.trace["foo"][0]) <= 1e-12 * ref.trace["foo"][0]
    # granular: assembled gradient and damping (halo sums) against the plain handle
    hp = spb.Handle(0)
    glob = spb.BuiltinProblem(hp, "lf", g=g, rows_mult=2, seed=20211)
    import torch
    T = glob.c_traits()
    m, nn, colptr, rowidx = glob.pattern()
    xd = torch.ones(nn, dtype=torch.float64, device="cuda")
    Ed = torch.empty(m, dtype=torch.float64, device="cuda")
    Jd = torch.empty(int(colptr[-1]), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    T.objective_jacobian(T.user, xd.data_ptr(), Ed.data_ptr(), Jd.data_ptr(), None)
    torch.cuda.synchronize()
    hp.analyze(m, nn, colptr, rowidx)
    hp.assemble(Jd, Ed, 0.01)
    Tl = loc.c_traits()
    ml, nl, lcp, lri = loc.pattern()
    xl = torch.ones(nl, dtype=torch.float64, device="cuda")
    El = torch.empty(ml, dtype=torch.float64, device="cuda")
    Jl = torch.empty(int(lcp[-1]), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    Tl.objective_jacobian(Tl.user, xl.data_ptr(), El.data_ptr(), Jl.data_ptr(), None)
    torch.cuda.synchronize()
    h.assemble(Jl, El, 0.01)
    np.testing.assert_allclose(h.rhs()[halo:halo + n], hp.rhs(), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(h.damping()[halo:halo + n], hp.damping(), rtol=1e-13, atol=1e-300)
    loc.close(); glob.close(); h.close(); hp.close()


def test_fused_and_launch_per_phase_halo_solves_agree(spb, monkeypatch):
    """Same local system, same right-hand side: the fused peer-window kernel and the launch-per-phase path must
    take the same number of PCG iterations and agree on the step to rounding (they differ only in the order the
    per-block partial dot products are added); consecutive fused solves reuse the windows (sequence numbers)."""
    import torch
    g = 48
    n, halo = g * g, 2 * g
    xs, its = [], []
    for mode in ("1", "0"):
        monkeypatch.setenv("SPB_HALO_PEER", mode)
        h = spb.Handle(0)
        loc = spb.BuiltinProblem(h, "lf_cols", g=g, rows_mult=2, seed=20211, col_begin=0, col_end=n)
        h.ring_halo(halo, halo + n, halo, 0, 0)
        h.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
        Tl = loc.c_traits()
        ml, nl, lcp, lri = loc.pattern()
        xl = torch.ones(nl, dtype=torch.float64, device="cuda")
        El = torch.empty(ml, dtype=torch.float64, device="cuda")
        Jl = torch.empty(int(lcp[-1]), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        Tl.objective_jacobian(Tl.user, xl.data_ptr(), El.data_ptr(), Jl.data_ptr(), None)
        torch.cuda.synchronize()
        h.analyze(ml, nl, lcp, lri)
        for lam in (0.01, 1.0, 0.01):   # three solves on the same windows
            h.assemble(Jl, El, lam)
            h.factorize()
            x, it, st = h.solve(None)
            assert st == spb.capi.SPB_OK
            xs.append(x)
            its.append(it)
        loc.close(); h.close()
    for a, b, ia, ib in zip(xs[:3], xs[3:], its[:3], its[3:]):
        assert abs(ia - ib) <= 1 and ia > 10
        assert np.linalg.norm(a - b) <= 1e-12 * np.linalg.norm(b)
        # ghosts are bit-identical copies of their owners in both paths
        assert np.array_equal(a[:halo], a[n:halo + n]) and np.array_equal(a[halo + n:], a[halo:2 * halo])
    assert np.linalg.norm(xs[0] - xs[2]) == 0.0   # deterministic: same system, same bits


def test_two_gpu_halo_job():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "scripts", "halo_check.py"), "--g", "96"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "HALO_CHECK_OK" in out.stdout
