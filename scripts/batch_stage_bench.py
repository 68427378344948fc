"""Stage breakdown of the fused batch path (spb_lm_solve_batch) on B extended-Rosenbrock problems of 10 000 unknowns:
ms per LM loop body by stage (CUDA events), PCG iterations per body.
    python scripts/batch_stage_bench.py [B=512]        (SPB_BATCH_BLOCK2=0: plain Jacobi preconditioner)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saddlepoint_b200 as spb
from saddlepoint_b200 import batch

h = spb.Handle(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
idx = list(range(0, 2 * B, 2))
batch.solve_batch_fused(h, idx[:8], keep_x=False, verbose=False)
h.stage_reset(True)
t0 = time.time()
r = batch.solve_batch_fused(h, idx, keep_x=False, verbose=False)
dt = time.time() - t0
st = h.stage_ms()
bodies = max(x["currIter"] for x in r)
print("B", B, "total %.3f s" % dt, "solve %.3f s" % r[0]["seconds"], "analyze %.3f s" % r[0]["seconds_analyze"], "bodies", bodies,
      "ms/body %.3f" % (1e3 * r[0]["seconds"] / bodies), "PCG iterations per body %.2f" % (sum(x["linear_iterations"] for x in r) / B / bodies))
print({k: (round(v[0] / bodies, 3), v[1] / bodies) for k, v in st.items() if v[1] > 0})
