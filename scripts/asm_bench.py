"""Times the fused assembly (column kernel + pair kernel) on an LF problem with device-resident inputs.
usage: python scripts/asm_bench.py [g=1024] [reps=20]        LF problem on a g x g torus
       python scripts/asm_bench.py c3 [nx=300] [reps=20]     seamless-integration structure (tests/c3_blocks.py), nx x nx x 2 faces
(SPB_ASM_V2=0 selects the general pair kernels)
Prints one JSON line: ms per assembly (CUDA events over `reps` back-to-back calls), the algorithmic bytes of
SURVEY section 8(d) and the fraction of the measured HBM peak."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saddlepoint_b200 as spb  # noqa: E402


def main():
    h = spb.Handle(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c3":
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        import c3_blocks
        g = int(sys.argv[2]) if len(sys.argv) > 2 else 300
        reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
        P = c3_blocks.build(g, g)
        prob = spb.BuiltinProblem(h, "blocks", **{k: P[k] for k in ("A", "b", "bar_idx", "bar_vol", "s", "w_bar", "x0")})
    else:
        g = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
        reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
        prob = spb.BuiltinProblem(h, "lf", g=g, rows_mult=2, seed=20211)
    m, n, colptr, rowidx = prob.pattern()
    h.analyze(m, n, colptr, rowidx)
    nnzJ = int(colptr[-1])
    rng = np.random.default_rng(1)
    Jv = torch.from_numpy(rng.uniform(-1, 1, nnzJ)).cuda()
    E = torch.from_numpy(rng.standard_normal(m)).cuda()
    for _ in range(3):
        h.assemble(Jv, E, 0.01)
    torch.cuda.synchronize()
    h.stage_reset(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        h.assemble(Jv, E, 0.01)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    st = h.stage_ms()["assembly"]
    nnzH = h.h_nnz()
    b_asm = 12 * nnzJ + 4 * (m + 1) + 8 * m + 8 * nnzH + 16 * n
    peak = 6541.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    kern_ms = st[0] / max(1, reps)
    print(json.dumps({"g": g, "n": n, "nnzJ": nnzJ, "nnzH": nnzH, "form": h.assembly_form(),
                      "ms_per_assembly_wall": ms, "ms_per_assembly_kernels": kern_ms, "launches": st[1] / max(1, reps),
                      "algorithmic_MB": b_asm / 1e6, "GBs": b_asm / (kern_ms * 1e-3) / 1e9, "frac_of_peak": b_asm / (kern_ms * 1e-3) / 1e9 / peak}))


if __name__ == "__main__":
    main()
