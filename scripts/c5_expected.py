"""One-GPU run of config C5's problem (LF g=4472, seed 20214, plain handle -- no halo mode, no communicator): the LM solve
with the reference's stop tests on.  Writes profiles/r2_c5_expected.json, which bench.py's N>1 runs (the same problem
over N ranks) compare their iteration count / final energy / checksums against.
    python scripts/c5_expected.py [--g 4472]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import saddlepoint_b200 as spb

ap = argparse.ArgumentParser()
ap.add_argument("--g", type=int, default=4472)
ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_c5_expected.json"))
args = ap.parse_args()
g = args.g
n = g * g
h = spb.Handle(0, torch.cuda.current_stream().cuda_stream)
prob = spb.BuiltinProblem(h, "lf", g=g, rows_mult=2, seed=20214)
ls = spb.B200SolverWrapper(h, spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
x = torch.empty(n, dtype=torch.float64, device="cuda")
# timing: fixed loop bodies, solves of 5 from x0 (as bench.py)
lm = spb.LMSolver()
lm.init(ls, prob, spb.DiagonalDamping(0.01), 100, funcTolerance=0.0, fooTolerance=0.0)
lm.solve(False, dedupe_evaluations=True, fixed_iterations=5, x_out=x)
analyze_s = lm.seconds_analyze
lm.DT.currLambda = 0.01
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lm.solve(False, dedupe_evaluations=True, fixed_iterations=5, x_out=x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
pcg_per_step = lm.linear_iterations / 5
# the checked solve: stop tests on
lm2 = spb.LMSolver()
lm2.init(ls, prob, spb.DiagonalDamping(0.01), 100)
ret = lm2.solve(True, dedupe_evaluations=True, x_out=x)
out = {"g": g, "n": n, "seed": 20214, "gpus": 1, "ms_per_lm_iteration": ms, "pcg_iterations_per_step": pcg_per_step, "analysis_seconds": analyze_s,
       "lm_solve": {"ret": bool(ret), "currIter": int(lm2.currIter), "exit_path": str(lm2.exit_path), "factorizations": int(lm2.factorizations),
                    "energy": float(lm2.energy), "x_sum": float(x.sum().item()), "x_sumsq": float((x * x).sum().item()),
                    "pcg_iterations": int(lm2.linear_iterations)}}
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(out, open(args.out, "w"), indent=1)
print(json.dumps(out))
