"""Times the direct (supernodal Cholesky) path on LF problems: analysis, factor, solve; residual check."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import saddlepoint_b200 as spb

for g in [int(a) for a in sys.argv[1:]] or [128, 256]:
    h = spb.Handle(0, torch.cuda.current_stream().cuda_stream)
    prob = spb.BuiltinProblem(h, "lf", g=g, rows_mult=2, seed=20211)
    m, n, colptr, rowidx = prob.pattern()
    T = prob.c_traits()
    x = torch.ones(n, dtype=torch.float64, device="cuda")
    E = torch.empty(m, dtype=torch.float64, device="cuda")
    Jv = torch.empty(int(colptr[-1]), dtype=torch.float64, device="cuda")
    T.objective_jacobian(T.user, x.data_ptr(), E.data_ptr(), Jv.data_ptr(), torch.cuda.current_stream().cuda_stream)
    t0 = time.time(); h.analyze(m, n, colptr, rowidx); t_an = time.time() - t0
    h.assemble(Jv, E, 0.01)
    h.set_solver(spb.SOLVER_CHOLESKY)
    t0 = time.time(); ok = h.factorize(); torch.cuda.synchronize(); t_first = time.time() - t0
    info = h.factor_info()
    h.stage_reset(True)
    reps = 3
    for _ in range(reps):
        h.assemble(Jv, E, 0.01); h.factorize()
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        xs, _, _ = h.solve(None, out=b)
    st = h.stage_ms()
    g_ = torch.tensor(h.rhs(), device="cuda")
    r = h.spmv(b) - g_
    h.set_solver(spb.SOLVER_PCG_JACOBI, 1e-13, 10000); h.factorize(); xp, it, _ = h.solve(None, out=torch.empty_like(b))
    print("g=%d n=%d ok=%s analysis(H)=%.2fs first_factor(incl symbolic)=%.2fs nnzL=%.3g (%.0f/col) flops=%.3g nsuper=%d levels=%d | factor %.2f ms (%.1f GFLOP/s, %d launches) solve %.2f ms (%d launches) | relres %.2e | vs pcg %.2e (pcg its %d)" % (
        g, n, ok, t_an, t_first, info["nnzL"], info["nnzL"] / n, info["flops"], info["nsuper"], info["nlevels"],
        st["factor"][0] / reps, info["flops"] / (st["factor"][0] / reps * 1e-3) / 1e9, st["factor"][1] // reps,
        st["trisolve"][0] / reps, st["trisolve"][1] // reps, float(r.norm() / g_.norm()), float((b - xp).norm() / xp.norm()), it))
    prob.close(); h.close()
