"""GPU tier: the row-partitioned (multi-GPU) path.  On one GPU the same code runs with a
communicator-free handle per row block (partial assembly, deferred damping, distributed-PCG code
path with world size 1); with >= 2 GPUs a torchrun job checks the real NCCL path against the
single-GPU solve."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import random_jacobian

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_full_range_partition_is_bit_exact(spb, orc):
    rng = np.random.default_rng(3)
    m, n = 900, 300
    colptr, rowidx, vals = random_jacobian(rng, m, n, 4, cover_cols=True)
    E = rng.standard_normal(m)
    h = spb.Handle(0)
    h.analyze_partitioned(m, n, colptr, rowidx, 0, m)
    lp, li = h.local_pattern()
    assert (lp == colptr).all() and (li == rowidx).all()
    foo = h.assemble(vals, E, 0.01)
    J = orc.from_csc(m, n, colptr, rowidx, vals)
    D, d = orc.damp_build(J, 0.01)
    hp, hi, hx = orc.ata(D).csc()
    assert (h.h_values() == hx).all() and (h.rhs() == orc.jt_vec_neg(J, E)).all() and (h.damping() == d).all()
    assert foo == np.abs(h.rhs()).max()
    h.close()


def test_row_blocks_sum_to_the_full_normal_equations(spb, orc):
    from saddlepoint_b200 import dist as sd
    rng = np.random.default_rng(4)
    m, n = 1200, 400
    colptr, rowidx, vals = random_jacobian(rng, m, n, 4, cover_cols=True)
    E = rng.standard_normal(m)
    full = spb.Handle(0)
    full.analyze(m, n, colptr, rowidx)
    full.assemble(vals, E, 0.0)
    Hf, gf = full.h_values(), full.rhs()
    Hs, gs, ds = 0.0, 0.0, 0.0
    for rank in range(3):
        rb, re = sd.row_partition(m, 3, rank)
        h = spb.Handle(0)
        h.analyze_partitioned(m, n, colptr, rowidx, rb, re)
        cp, ri = h.h_pattern()
        assert (cp == full.h_pattern()[0]).all() and (ri == full.h_pattern()[1]).all()  # identical H pattern on every rank
        h.assemble(sd.local_values(colptr, rowidx, vals, rb, re), E[rb:re], 0.0)
        Hs = Hs + h.h_values()
        gs = gs + h.rhs()
        h.close()
    np.testing.assert_allclose(Hs, Hf, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(gs, gf, rtol=1e-13, atol=1e-13)
    full.close()


def test_partitioned_lm_world1_matches_plain(spb, orc):
    h = spb.Handle(0)
    g = 24
    m = 2 * g * g
    ls = spb.B200SolverWrapper(h)
    lm = spb.LMSolver()
    prob = spb.BuiltinProblem(h, "lf_rows", g=g, rows_mult=2, seed=20211, row_begin=0, row_end=m)
    lm.init(ls, prob, spb.DiagonalDamping(0.01), 100)
    ret = lm.solve(True)
    ref = orc.lm_solve("lf", iparam0=g, iparam1=2, maxIterations=100)
    assert ret == ref.ret and lm.currIter == ref.currIter and lm.exit_path == ref.exit_path
    assert np.linalg.norm(lm.x - ref.x) <= 1e-10 * np.linalg.norm(ref.x)
    h.close()


def test_two_gpu_nccl_job():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "scripts", "dist_check.py"), "--g", "96"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_CHECK_OK" in out.stdout
