"""Config C4 throughput: independent 10k-unknown problems per second on one GPU, for 1..W concurrent workers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saddlepoint_b200 as spb
from saddlepoint_b200 import batch

nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = sys.argv[2] if len(sys.argv) > 2 else "lf"
spec = (lambda i: batch.c4_problem_spec(2 * i + 1)) if kind == "lf" else (lambda i: batch.c4_problem_spec(2 * i))
for workers in (1, 2, 4, 8, 16):
    batch.solve_batch(min(nprob, 2 * workers), make_spec=spec, solver=spb.SOLVER_PCG_JACOBI, workers=workers)  # warm-up
    t0 = time.time()
    res = batch.solve_batch(nprob, make_spec=spec, solver=spb.SOLVER_PCG_JACOBI, workers=workers)
    dt = time.time() - t0
    its = sum(r["currIter"] for r in res)
    print("C4 %s n=10000: %d problems, %d workers: %.2f s, %.1f problems/s, %.0f LM iterations/s" % (kind, nprob, workers, dt, nprob / dt, its / dt), flush=True)
