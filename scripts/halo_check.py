"""torchrun job: LM solve of ONE LF problem in subdomain (halo) mode over all ranks vs the single-GPU solve.
Every rank analyses and assembles only its local problem (owned unknowns + halo); per PCG iteration two halo
slabs and two scalar allreduces cross NVLink.
    python -m torch.distributed.run --nproc-per-node N scripts/halo_check.py --g 256 [--iters K]
Prints HALO_CHECK_OK on success and the device time per LM iteration."""
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
ap.add_argument("--no-reference", action="store_true", help="skip the single-GPU comparison (sizes that do not fit / take long)")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = args.g
n = g * g
halo = 2 * g
h = spb.Handle(local, torch.cuda.current_stream().cuda_stream)
sd.attach_comm(h, dist)
cb, ce = sd.column_partition(n, world, rank, halo)
left, right = sd.ring_neighbours(world, rank)
prob = spb.BuiltinProblem(h, "lf_cols", g=g, rows_mult=2, seed=20211, col_begin=cb, col_end=ce)
own = ce - cb
h.ring_halo(halo, halo + own, halo, left, right)
ls = spb.B200SolverWrapper(h, spb.SOLVER_PCG_JACOBI, 1e-13, 10000)
lm = spb.LMSolver()
fixed = args.iters
lm.init(ls, prob, spb.DiagonalDamping(0.01), 100, *((0.0, 0.0) if fixed else ()))
x = torch.empty(prob.xSize, dtype=torch.float64, device="cuda")
lm.solve(not fixed, dedupe_evaluations=True, fixed_iterations=fixed, x_out=x)  # warm-up incl. the (local) analysis
t_analyze = lm.seconds_analyze
lm.DT.currLambda = 0.01
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ret = lm.solve(not fixed, dedupe_evaluations=True, fixed_iterations=fixed, x_out=x)
e1.record(); dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
bodies = lm.factorizations
# ghost copies must equal the owners' values: gather the owned parts, compare the ghosts against them
owned = x[halo:halo + own].contiguous()
sizes = [sd.shard_range(n, world, r) for r in range(world)]
parts = [torch.empty(hi - lo, dtype=torch.float64, device="cuda") for lo, hi in sizes]
dist.all_gather(parts, owned) if len(set(hi - lo for lo, hi in sizes)) == 1 else None
ok = True
xg = None
if len(set(hi - lo for lo, hi in sizes)) == 1:
    xg = torch.cat(parts)
    idx = (torch.arange(cb - halo, ce + halo, device="cuda") % n)
    ok = bool((xg[idx] == x).all())  # ghosts consistent with their owners, bit for bit
if rank == 0 and xg is not None and not args.no_reference:
    h1 = spb.Handle(local)
    p1 = spb.BuiltinProblem(h1, "lf", g=g, rows_mult=2, seed=20211)
    l1 = spb.LMSolver()
    l1.init(spb.B200SolverWrapper(h1, spb.SOLVER_PCG_JACOBI, 1e-13, 10000), p1, spb.DiagonalDamping(0.01), 100, *((0.0, 0.0) if fixed else ()))
    x1 = torch.empty(n, dtype=torch.float64, device="cuda")
    r1 = l1.solve(not fixed, dedupe_evaluations=True, fixed_iterations=fixed, x_out=x1)
    rel = float((xg - x1).norm() / x1.norm())
    erel = abs(lm.energy - l1.energy) / max(abs(l1.energy), 1e-300)
    ok = ok and (r1 == ret) and (l1.currIter == lm.currIter) and (l1.factorizations == lm.factorizations) and rel <= 1e-10
    print("world=%d g=%d n=%d (local %d + 2x%d) bodies=%d currIter=%d/%d ret=%s/%s rel_diff_vs_1gpu=%.2e energy_rel=%.2e ghosts_consistent=%s"
          % (world, g, n, own, halo, bodies, lm.currIter, l1.currIter, ret, r1, rel, erel, ok))
if rank == 0:
    print("halo mode: %.3f ms per LM iteration, %d PCG iterations total, %d collectives, local analysis %.2f s"
          % (float(ms.item()) / max(1, bodies), lm.linear_iterations, h.collective_count(), t_analyze))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("HALO_CHECK_OK" if int(flag.item()) == 1 else "HALO_CHECK_FAILED")
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
