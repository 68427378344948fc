"""torchrun job: row-partitioned LM solve over all ranks (NCCL) vs the single-GPU solve on rank 0.
    python -m torch.distributed.run --nproc-per-node N scripts/dist_check.py --g 256 [--iters K]
Prints DIST_CHECK_OK on success and the device time per LM iteration."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import saddlepoint_b200 as spb
from saddlepoint_b200 import dist as sd

ap = argparse.ArgumentParser()
ap.add_argument("--g", type=int, default=128)
ap.add_argument("--iters", type=int, default=0, help=">0: fixed iterations (timing mode)")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = args.g
n, m = g * g, 2 * g * g
h = spb.Handle(local, torch.cuda.current_stream().cuda_stream)
sd.attach_comm(h, dist)
rb, re = sd.row_partition(m, world, rank)
prob = spb.BuiltinProblem(h, "lf_rows", g=g, rows_mult=2, seed=20211, row_begin=rb, row_end=re)
ls = spb.B200SolverWrapper(h, spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
lm = spb.LMSolver()
fixed = args.iters
lm.init(ls, prob, spb.DiagonalDamping(0.01), 100, *( (0.0, 0.0) if fixed else () ))
x = torch.empty(n, dtype=torch.float64, device="cuda")
lm.solve(not fixed, dedupe_evaluations=True, fixed_iterations=fixed, x_out=x)  # warm-up incl. analysis
lm.DT.currLambda = 0.01  # the solve wrote the final lambda back (like the reference's DT object)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ret = lm.solve(not fixed, dedupe_evaluations=True, fixed_iterations=fixed, x_out=x)
e1.record(); dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
bodies = lm.factorizations
# every rank must hold the same iterate bit for bit
xs = [torch.empty_like(x) for _ in range(world)]
dist.all_gather(xs, x)
same = all(bool((xs[0] == t).all()) for t in xs)
ok = same
if rank == 0:
    h1 = spb.Handle(local)
    p1 = spb.BuiltinProblem(h1, "lf", g=g, rows_mult=2, seed=20211)
    l1 = spb.LMSolver()
    l1.init(spb.B200SolverWrapper(h1, spb.SOLVER_PCG_JACOBI, 1e-13, 10000), p1, spb.DiagonalDamping(0.01), 100, *((0.0, 0.0) if fixed else ()))
    x1 = torch.empty(n, dtype=torch.float64, device="cuda")
    r1 = l1.solve(not fixed, dedupe_evaluations=True, fixed_iterations=fixed, x_out=x1)
    rel = float((x - x1).norm() / x1.norm())
    ok = ok and (r1 == ret) and (l1.currIter == lm.currIter) and (l1.factorizations == lm.factorizations) and rel <= 1e-10
    print("world=%d g=%d n=%d bodies=%d currIter=%d/%d ret=%s/%s rel_diff_vs_1gpu=%.2e identical_across_ranks=%s ms_per_iteration=%.3f pcg_iters=%d collectives=%d"
          % (world, g, n, bodies, lm.currIter, l1.currIter, ret, r1, rel, same, float(ms.item()) / max(1, bodies), lm.linear_iterations, h.collective_count()))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK_OK" if int(flag.item()) == 1 else "DIST_CHECK_FAILED")
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
