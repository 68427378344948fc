"""Config C3 stand-in at scale: the seamless-integration Jacobian structure (regular grid,
N=4 field per face, no seams) with device-resident traits (spb_problem_blocks), LM with the direct
solver (the reference's SimplicialLLT role).  Reports one-time costs and the per-iteration budget.

    python scripts/c3_bench.py [nx ny [iterations]]      (500 500 = 498 002 faces, 4.48 M unknowns)
"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import saddlepoint_b200 as spb
import c3_blocks

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ny = int(sys.argv[2]) if len(sys.argv) > 2 else nx
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 8
solver = {"chol": spb.SOLVER_CHOLESKY, "pcg": spb.SOLVER_PCG_JACOBI, "ic0": spb.SOLVER_PCG_IC0}[sys.argv[4] if len(sys.argv) > 4 else "chol"]
t0 = time.time()
P = c3_blocks.build(nx, ny)
t_build = time.time() - t0
ls = spb.B200SolverWrapper(solver=solver)
h = ls.handle
prob = spb.BuiltinProblem(h, "blocks", **{k: P[k] for k in ("A", "b", "bar_idx", "bar_vol", "s", "w_bar", "x0")})
m, n, colptr, rowidx = prob.pattern()
print("C3 stand-in: %d faces, n=%d unknowns, m=%d residuals, nnzJ=%d (problem built on the host in %.1f s)" % (P["nF"], n, m, colptr[-1], t_build), flush=True)
lm = spb.LMSolver()
lm.init(ls, prob, spb.DiagonalDamping(0.01), iters)
h.stage_reset(True)
t0 = time.time()
ret = lm.solve(True, dedupe_evaluations=True)
t_first = time.time() - t0
st = h.stage_ms()
nit = max(1, lm.factorizations)
info = h.factor_info() if solver == spb.SOLVER_CHOLESKY else None
print("first solve: %.2f s total, %.2f s of it pattern analysis; %d LM iterations (%d factorisations), exit %s, energy %.6e" % (
    t_first, lm.seconds_analyze, lm.currIter, lm.factorizations, lm.exit_path, lm.energy), flush=True)
print("nnzH=%d" % h.h_nnz(), ("nnzL=%.3g flops/factorisation=%.3g supernodes=%d levels=%d" % (info["nnzL"], info["flops"], info["nsuper"], info["nlevels"])) if info else "")
print("per LM iteration (ms): " + "  ".join("%s %.2f" % (k, v[0] / nit) for k, v in st.items() if v[0] > 0))
# steady state: second solve on the analysed handle
lm2 = spb.LMSolver()
lm2.init(ls, prob, spb.DiagonalDamping(0.01), iters)
h.stage_reset(True)
t0 = time.time()
lm2.solve(True, dedupe_evaluations=True)
t2 = time.time() - t0
print("second solve (analysis + symbolic factorisation cached): %.3f s for %d iterations = %.1f ms/iteration, linear iterations %d" % (
    t2, lm2.factorizations, 1e3 * t2 / max(1, lm2.factorizations), lm2.linear_iterations))
