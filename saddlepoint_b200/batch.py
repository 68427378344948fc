"""Config C4: batches of independent LM problems, split per GPU.

Every problem is a complete LMSolver::solve with its own lambda, acceptance and stop test, so
problems finish after different iteration counts.  The batch is sharded in contiguous index blocks
(one handle per GPU, no data-path collective); results are gathered at the end.  Inside a rank the
problems run back to back on one handle: the Jacobian pattern of same-shaped problems is analysed
once (pattern fingerprint), buffers are reused.  With workers > 1 the rank keeps that many handles
(each with its own CUDA stream) and host threads: the problems are small (their persistent PCG
kernels occupy a few dozen CTAs, their LM loops are latency- and host-bound), so several of them
share the GPU; every problem still runs the unmodified per-problem loop, results are identical.
"""
from concurrent.futures import ThreadPoolExecutor
import threading

import numpy as np

from . import dist as sd
from .capi import SOLVER_CHOLESKY, SOLVER_PCG_JACOBI
from .solver import B200SolverWrapper, BuiltinProblem, DiagonalDamping, Handle, LMSolver

_GOLD = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


def _splitmix64_at(seed, ctr):
    z = (seed + (ctr + 1) * _GOLD) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def c4_problem_spec(index, n=10000, seed=20213):
    """Problem `index` of config C4: even indices are extended Rosenbrock with a start
    jittered per problem, odd indices LF on a sqrt(n) x sqrt(n) torus with a per-problem seed."""
    if index % 2 == 0:
        nb = n // 2
        u = np.array([_splitmix64_at(seed, index * 4 * nb + k) >> 11 for k in range(2 * nb)], dtype=np.float64) / 9007199254740992.0
        x0 = np.tile([-100.0, 100.0], nb) * (1.0 + 0.01 * (2.0 * u - 1.0))
        return {"kind": "rosenbrock", "nblocks": nb, "x0": x0, "maxIterations": 1000}
    g = int(round(n ** 0.5))
    return {"kind": "lf", "g": g, "rows_mult": 2, "seed": seed + index, "maxIterations": 100}


def solve_problem(handle, spec, solver=SOLVER_CHOLESKY, verbose=True):
    prob = BuiltinProblem(handle, spec["kind"], **{k: v for k, v in spec.items() if k not in ("kind", "maxIterations")})
    ls = B200SolverWrapper(handle, solver)
    lm = LMSolver()
    lm.init(ls, prob, DiagonalDamping(0.01), spec.get("maxIterations", 100))
    ret = lm.solve(verbose, dedupe_evaluations=True)
    out = {"ret": ret, "currIter": lm.currIter, "factorizations": lm.factorizations, "energy": lm.energy,
           "exit_path": lm.exit_path, "x": lm.x.copy(), "seconds": lm.seconds_total}
    prob.close()
    return out


def solve_batch(nproblems, rank=0, world=1, handle=None, make_spec=c4_problem_spec, solver=SOLVER_CHOLESKY, dist=None, keep_x=False,
                workers=1):
    """Solve this rank's block of the batch; with world > 1 gather every rank's results."""
    lo, hi = sd.shard_range(nproblems, world, rank)

    def one(h, i):
        r = solve_problem(h, make_spec(i), solver)
        if not keep_x:
            r["xnorm"] = float(np.linalg.norm(r.pop("x")))
        return (i, r)

    if workers <= 1:
        own = handle is None
        handle = handle or Handle(0)
        mine = [one(handle, i) for i in range(lo, hi)]
        if own:
            handle.close()
    else:
        # one handle (and stream) per worker thread; ctypes releases the GIL inside the library
        local = threading.local()
        handles = []
        lock = threading.Lock()

        def task(i):
            if not hasattr(local, "h"):
                local.h = Handle(handle.device if handle is not None else 0)
                with lock:
                    handles.append(local.h)
            return one(local.h, i)

        with ThreadPoolExecutor(max_workers=workers) as pool:
            mine = list(pool.map(task, range(lo, hi)))
        for h in handles:
            h.close()
    if world == 1:
        return [r for _, r in mine]
    return sd.merge_problem_results(sd.gather_objects(mine, dist))
