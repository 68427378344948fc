#!/usr/bin/env python
"""bench.py -- LM iterations/s at the 1M-unknown sparse least-squares workload (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--g 1024]

A *step* is one Levenberg-Marquardt loop body (LMSolver.h:110-188): fused assembly of
rhs / ||rhs||_inf / damping / H, the damped sparse solve, the trial evaluation, acceptance and
the lambda update.  Workload: config C2, "LF-1M" (sparse analogue of
benchmark/linear_function: g x g torus, n = g^2 unknowns, m = 2n residuals, 8 nnz/row; g=1024).
  value : steps/s with the traits evaluated on the device (inputs resident in HBM), K loop
          bodies of one solve started at x0 (lambda follows the LM rule, stop tests off).
  e2e   : the same hot path through the C ABI with HOST buffers -- every step uploads the
          Jacobian values and the residual from pinned host memory, runs assemble + factorize +
          solve, and downloads the step vector and ||rhs||_inf.
  roofline : the dominant kernel (the PCG's SpMV) re-timed per launch with CUDA events.
  cpu_baseline / --impl reference : the CPU oracle (Eigen restatement, 1 thread) on a bounded
          sample of the same workload.
N>1: one process per GPU (torchrun), each rank solves its own independent instance of the same LF
configuration: the "batches of independent problems split per GPU" sharding -- no data-path collective;
plus, as an extra key, ONE C5-size problem solved by all ranks together (subdomain / halo mode, NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _label(g):
    """BASELINE.json's name for the configuration quoted (C2) or the sizes of the other LF configs."""
    return "LF-1M (C2)" if g == 1024 else ("LF-20M (C5 size, one GPU)" if g == 4472 else "LF (g=%d, not a BASELINE configuration)" % g)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ncu --set full capture of pcg_persistent_kernel on LF-1M (profiles/r1c_ncu_raw_metrics.txt, end of round 1):
# dram__bytes_read.sum + dram__bytes_write.sum = 21.380 GB + 0.289 GB for an 81-iteration solve (5.31 ms)
# = 267.5 MB per PCG iteration (r1b: 25.394 GB / 96 iterations = 264.6 MB): the 10-byte/entry H stream (FP64
# value + 16-bit column delta, 26.2 M padded entries) plus the solve's first and last vector passes; the PCG
# vectors never leave the 126 MB L2.
NCU_DRAM_BYTES_PER_PCG_ITERATION = 21.669e9 / 81.0


def lf_sizes(g):
    n = g * g
    return n, 2 * n, 16 * n


# ------------------------------------------------------------------------------------------
def cpu_reference(args, sample_g=None, steps=None, warmup=None):
    """The reference's own CPU path (oracle: Eigen restatement, SimplicialLLT with AMD, 1 thread)
    on a bounded sample of the workload.  Returns (steps_per_s_scaled, info)."""
    from oracle import sp_oracle as orc
    g_full = args.g
    gs = sample_g or min(g_full, 160)
    steps = steps or 2
    warmup = 0 if warmup is None else warmup
    n_s, m_s, _ = lf_sizes(gs)
    t0 = time.time()
    r = orc.lm_solve("lf", iparam0=gs, iparam1=2, seed=20211, maxIterations=max(0, warmup + steps - 1), funcTolerance=0.0,
                     fooTolerance=0.0, verbose=False, solver_kind=0)
    wall = time.time() - t0
    bodies = max(1, r.factorizations)
    per_iter = (r.times["assembly"] + r.times["factor"] + r.times["solve"]) / bodies
    n_full = g_full * g_full
    scale = n_s / float(n_full)  # linear extrapolation in the unknown count: optimistic for the CPU
    info = {"kind": "port", "cores": 1,
            "sample": "oracle (Eigen restatement: conservative SpGEMM + AMD + scalar up-looking LLT), LF g=%d (n=%d, 1/%.0f of the "
                      "unknowns), %d loop bodies, assembly+factor+solve %.3f s/iteration, scaled linearly in n to g=%d "
                      "(factorisation cost grows faster than n, so this over-states the CPU rate)" % (gs, n_s, 1.0 / scale, bodies, per_iter, g_full),
            "seconds_per_iteration_sample": per_iter, "wall_s": wall,
            "split_s": {k: v / bodies for k, v in r.times.items()}}
    return (1.0 / per_iter) * scale, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, info = cpu_reference(args, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    n, m, nnzJ = lf_sizes(args.g)
    info["value"] = val
    line = {"impl": "reference", "metric": "LM iterations/s", "value": val, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "benchmark/linear_function %s: g=%d torus, n=%d, m=%d, nnzJ=%d" % (_label(args.g), args.g, n, m, nnzJ)},
            "cpu_baseline": info,
            "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import saddlepoint_b200 as spb

    K, W, g = args.steps, max(args.warmup, 3), args.g
    n, m, nnzJ = lf_sizes(g)
    stream = torch.cuda.current_stream().cuda_stream
    h = spb.Handle(local, stream)
    solver_id = spb.SOLVER_PCG_IC0 if args.solver == "pcg_ic0" else spb.SOLVER_PCG_JACOBI
    h.set_solver(solver_id, args.pcg_rtol, 10000)
    # every rank solves its own instance of the SAME configuration (seed 20211): per-GPU work is identical, so the
    # max-over-ranks time measures scaling and not the spread of PCG iteration counts between different problems
    prob = spb.BuiltinProblem(h, "lf", g=g, rows_mult=2, seed=20211)
    ls = spb.B200SolverWrapper(h, solver_id, args.pcg_rtol, 10000)
    lm = spb.LMSolver()
    lm.init(ls, prob, spb.DiagonalDamping(0.01), 100, funcTolerance=0.0, fooTolerance=0.0)
    x_dev = torch.empty(n, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident run (value) -------------------------------------------------------
    # The LF problem is linear: LM converges after a handful of loop bodies and later bodies of the same solve
    # are cheap (the gradient is ~0, the PCG stops at once).  So that a step costs the same whatever --steps is,
    # the solve is restarted from x0 every BODIES_PER_SOLVE loop bodies (the default --steps).
    BODIES_PER_SOLVE = 5

    def run_steps(k):
        iters = 0
        while k > 0:
            lm.DT.currLambda = 0.01  # every solve starts at the reference's lambda0 (the solve writes the final lambda back)
            lm.solve(False, dedupe_evaluations=True, fixed_iterations=min(k, BODIES_PER_SOLVE), x_out=x_dev)
            iters += int(lm.linear_iterations)
            k -= BODIES_PER_SOLVE
        return iters

    run_steps(W)  # warm-up (+ one-time analysis)
    analyze_s = lm.seconds_analyze
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = h.launch_count()
    lin_box = []
    ms = timed(lambda: lin_box.append(run_steps(K)))
    launches = h.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    lin_iters = int(lin_box[0])
    value = world * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel: same K steps, every launch bracketed by CUDA events
    h.stage_reset(True)
    lin_iters_roof = run_steps(K)
    stages = h.stage_ms()
    h.stage_reset(False)
    nnzH = h.h_nnz()
    peaks, peak_kind = _peaks()
    # B_spmv = 12*nnzH + 4(n+1) + 8n + 8n bytes per product; a PCG iteration moves
    # B_spmv + 10*8n (x, r, p, q, dinv reads/writes around the product)
    b_spmv = 12 * nnzH + 4 * (n + 1) + 16 * n
    b_pcg = b_spmv + 80 * n
    total_ms = max(1e-9, sum(v[0] for v in stages.values()))
    fused_ms, fused_launches = stages["pcg_fused"]
    if fused_ms > 0:
        kname = "pcg_persistent_kernel (whole Jacobi-PCG solve: SpMV + vector updates + reductions, one cooperative launch)"
        k_ms, k_launches, k_bytes_total = fused_ms, fused_launches, b_pcg * max(1, lin_iters)
    else:
        kname = "spmv_sell_kernel<true> (PCG SpMV + p.q)"
        k_ms, k_launches, k_bytes_total = stages["spmv"][0], max(1, lin_iters), b_spmv * max(1, lin_iters)
    achieved = k_bytes_total / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    asm_ms, asm_launches = stages["assembly"]
    b_asm = 12 * nnzJ + 4 * (m + 1) + 8 * m + 8 * nnzH + 16 * n
    # DRAM bytes per launch from the committed ncu --set full capture (profiles/): per PCG iteration
    # dram__bytes_read+write, scaled by this launch's iteration count; only valid for the g=1024 workload
    traffic = NCU_DRAM_BYTES_PER_PCG_ITERATION * max(1, lin_iters) / max(1, k_launches) if (fused_ms > 0 and g == 1024) else None
    roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
            "bytes_per_launch": k_bytes_total / max(1, k_launches), "avg_launch_us": 1e3 * k_ms / max(1, k_launches),
            "launches_timed": int(k_launches), "pcg_iterations_timed": lin_iters, "us_per_pcg_iteration": 1e3 * k_ms / max(1, lin_iters),
            "share_of_step": k_ms / total_ms,
            "assembly": {"achieved": b_asm * K / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 else 0.0, "bytes_per_step": b_asm,
                         "ms_per_step": asm_ms / K},
            "stage_ms_per_step": {k: v[0] / K for k, v in stages.items()}}
    roof["assembly"]["frac"] = roof["assembly"]["achieved"] / peaks["hbm_gbs"]

    # ---- e2e: the C-ABI hot path with HOST buffers (H2D + D2H inside the timed region) -------
    mrows, ncols, colptr, rowidx = prob.pattern()
    Jv_host = torch.empty(nnzJ, dtype=torch.float64).pin_memory()
    E_host = torch.empty(m, dtype=torch.float64).pin_memory()
    dir_host = torch.empty(n, dtype=torch.float64).pin_memory()
    # contents: the problem's own J values and residual at x0 (generated on the device, copied out once)
    Tt = prob.c_traits()
    xd = torch.ones(n, dtype=torch.float64, device="cuda")
    Ed = torch.empty(m, dtype=torch.float64, device="cuda")
    Jd = torch.empty(nnzJ, dtype=torch.float64, device="cuda")
    Tt.objective_jacobian(Tt.user, xd.data_ptr(), Ed.data_ptr(), Jd.data_ptr(), stream)
    torch.cuda.synchronize()
    Jv_host.copy_(Jd)
    E_host.copy_(Ed)
    del xd, Ed, Jd
    Jn, En, dn_ = Jv_host.numpy(), E_host.numpy(), dir_host.numpy()

    def e2e_steps(k):
        for _ in range(k):
            h.assemble(Jn, En, 0.01)
            h.factorize()
            h.solve(None, out=dn_)

    e2e_steps(min(W, 3))
    ms_e2e = timed(lambda: e2e_steps(K))
    e2e_val = world * K / (ms_e2e * 1e-3)

    # ---- the direct path (supernodal Cholesky) on the same system, for the record (N=1 only) --
    direct = None
    if rank == 0 and world == 1 and not args.no_direct:
        h.set_solver(spb.SOLVER_CHOLESKY)
        h.assemble(Jn, En, 0.01)
        h.factorize()  # includes the one-time symbolic phase (ordering, supernodes, maps)
        h.solve(None, out=dn_)
        h.stage_reset(True)
        reps = 3
        for _ in range(reps):
            h.assemble(Jn, En, 0.01)
            h.factorize()
            h.solve(None, out=dn_)
        st2 = h.stage_ms()
        h.stage_reset(False)
        info = h.factor_info()
        direct = {"factor_ms": st2["factor"][0] / reps, "solve_ms": st2["trisolve"][0] / reps, "nnzL": info["nnzL"], "flops": info["flops"],
                  "factor_tflops_fp64": info["flops"] / (st2["factor"][0] / reps * 1e-3) / 1e12, "supernodes": info["nsuper"],
                  "tree_levels": info["nlevels"],
                  "note": "multifrontal LL^T, nested-dissection order; SYRK/GEMM on FP64 tensor cores (DMMA); LM steps/s with this "
                          "solver = 1000/(assembly+factor+solve ms)"}
        h.set_solver(solver_id, args.pcg_rtol, 10000)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cpu = cpu_reference(args)
        cpu["value"] = v
        cpu["unit"] = "iterations/s"

    # ---- N > 1 extra (not the headline): ONE problem of config C5's size over all ranks, subdomain (halo) mode
    one_problem = None
    if world > 1 and not args.no_one_problem:
        one_problem = one_problem_all_gpus(spb, dist, torch, local, rank, world, args)

    if rank == 0:
        line = {"metric": "LM iterations/s", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "benchmark/linear_function %s: g=%d torus, n=%d, m=%d, nnzJ=%d, nnzH=%d" % (_label(g), g, n, m, nnzJ, nnzH),
                           "per_gpu": "one independent problem per GPU (every rank its own instance of the same configuration, seed 20211)", "solver": "%s rtol=%g" % (args.solver, args.pcg_rtol),
                           "l2": "inputs larger than L2 (H %d MB + J %d MB per step vs 126 MB L2)" % (12 * nnzH // 2**20, 12 * nnzJ // 2**20),
                           "pcg_iterations_per_step": lin_iters / K, "analysis_seconds_one_time": analyze_s,
                           "lm_bodies_per_solve": BODIES_PER_SOLVE},
                "e2e": {"value": e2e_val, "unit": "iterations/s", "h2d_bytes_per_step": 8 * (nnzJ + m), "d2h_bytes_per_step": 8 * n + 8,
                        "ms_per_step": ms_e2e / K},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "direct_solver": direct}
        if one_problem is not None:
            line["one_problem_all_gpus"] = one_problem
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def one_problem_all_gpus(spb, dist, torch, local, rank, world, args, iters=3):
    """Strong-scaling data point next to the (weak-scaling) headline: one LF problem of config C5's size
    (g=4472, n ~ 20 M; the 1-GPU time of the same problem is in DESIGN.md) solved by all ranks together in
    subdomain (halo) mode -- local analysis per rank, one halo exchange and two scalar allreduces per PCG
    iteration over NVLink.  Never fails the bench line: errors are reported as a string."""
    try:
        from saddlepoint_b200 import dist as sd
        g = args.one_problem_g
        n = g * g
        halo = 2 * g
        h = spb.Handle(local, torch.cuda.current_stream().cuda_stream)
        sd.attach_comm(h, dist)
        cb, ce = sd.column_partition(n, world, rank, halo)
        left, right = sd.ring_neighbours(world, rank)
        prob = spb.BuiltinProblem(h, "lf_cols", g=g, rows_mult=2, seed=20214, col_begin=cb, col_end=ce)
        h.ring_halo(halo, halo + (ce - cb), halo, left, right)
        ls = spb.B200SolverWrapper(h, spb.SOLVER_PCG_JACOBI, args.pcg_rtol, 10000)
        lm = spb.LMSolver()
        lm.init(ls, prob, spb.DiagonalDamping(0.01), 100, funcTolerance=0.0, fooTolerance=0.0)
        x = torch.empty(prob.xSize, dtype=torch.float64, device="cuda")
        lm.solve(False, dedupe_evaluations=True, fixed_iterations=1, x_out=x)  # warm-up incl. the local analysis
        analyze_s = lm.seconds_analyze
        lm.DT.currLambda = 0.01
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lm.solve(False, dedupe_evaluations=True, fixed_iterations=iters, x_out=x)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = {"workload": "LF g=%d (n=%d, m=%d) as ONE problem over %d GPUs, subdomain (halo) mode" % (g, n, 2 * n, world),
               "ms_per_lm_iteration": float(t.item()) / iters, "lm_iterations_per_s": iters / (float(t.item()) * 1e-3),
               "pcg_iterations_per_lm_iteration": lm.linear_iterations / iters, "local_analysis_seconds": analyze_s,
               "collectives": int(h.collective_count()), "scaling": "strong"}
        prob.close()
        h.close()
        return out
    except Exception as e:  # noqa: BLE001 -- an extra must not break the bench line
        return {"error": "%s: %s" % (type(e).__name__, e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--g", type=int, default=1024)
    ap.add_argument("--pcg-rtol", type=float, default=1e-13)
    ap.add_argument("--solver", default="pcg_jacobi", choices=["pcg_jacobi", "pcg_ic0"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-direct", action="store_true", help="skip the supernodal-Cholesky timing block")
    ap.add_argument("--no-one-problem", action="store_true", help="N>1: skip the one-problem-over-all-GPUs (halo mode) extra")
    ap.add_argument("--one-problem-g", type=int, default=4472, help="torus size of that extra (config C5: 4472)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
