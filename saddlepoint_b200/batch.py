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


EXIT = {1: "foo", 2: "factorize_failed", 3: "direction_small", 4: "func_tol", 5: "traits_stop", 6: "max_iter"}


def solve_batch_fused(handle, indices, n=10000, seed=20213, verbose=True, keep_x=False, pcg_rtol=1e-13):
    """The C4 problems `indices` (all Rosenbrock = even indices, or all LF = odd indices) as ONE block-diagonal
    problem through spb_lm_solve_batch: one fused assembly, one persistent PCG kernel and one controller launch per
    stage for all of them; per-problem lambda / acceptance / stop tests on the device.  Returns one dict per problem
    (same keys as solve_problem)."""
    import ctypes as C
    from . import capi
    from .solver import _c_traits
    indices = list(indices)
    B = len(indices)
    kinds = {i % 2 for i in indices}
    if len(kinds) != 1:
        raise ValueError("a fused batch holds problems of one kind (even indices: Rosenbrock, odd: LF)")
    if 0 in kinds:
        specs = [c4_problem_spec(i, n=n, seed=seed) for i in indices]
        nb = specs[0]["nblocks"]
        prob = BuiltinProblem(handle, "rosenbrock", nblocks=nb * B, x0=np.concatenate([s["x0"] for s in specs]))
        max_it = specs[0]["maxIterations"]
    else:
        stride = indices[1] - indices[0] if B > 1 else 1
        if any(indices[k + 1] - indices[k] != stride for k in range(B - 1)):
            raise ValueError("LF batch indices must be equally spaced")
        spec0 = c4_problem_spec(indices[0], n=n, seed=seed)
        prob = BuiltinProblem(handle, "lf_batch", g=spec0["g"], rows_mult=2, seed=spec0["seed"], seed_stride=stride, nproblems=B)
        max_it = spec0["maxIterations"]
    handle.set_solver(SOLVER_PCG_JACOBI, pcg_rtol, 10000)
    T, keep = _c_traits(prob)
    O = capi.LMOptions()
    handle.L.spb_lm_default_options(C.byref(O))
    O.maxIterations = max_it
    O.verbose = int(bool(verbose))
    O.dedupe_evaluations = 1
    R = (capi.LMResult * B)()
    x = np.empty(T.xSize, np.float64)
    handle._check(handle.L.spb_lm_solve_batch(handle.h, C.byref(T), B, C.byref(O), R, C.c_void_p(x.ctypes.data), capi.SPB_HOST))
    n_per = T.xSize // B
    out = []
    for p in range(B):
        r = {"ret": bool(R[p].ret), "currIter": R[p].currIter, "factorizations": R[p].factorizations, "energy": R[p].energy,
             "exit_path": EXIT.get(R[p].exit_path, R[p].exit_path), "lambda": R[p].lambda_, "linear_iterations": R[p].linear_iterations,
             "seconds": R[p].seconds_total, "seconds_analyze": R[p].seconds_analyze}
        xp = x[p * n_per:(p + 1) * n_per]
        if keep_x:
            r["x"] = xp.copy()
        else:
            r["xnorm"] = float(np.linalg.norm(xp))
        out.append(r)
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
