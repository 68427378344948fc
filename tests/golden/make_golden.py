"""Generates tests/golden/trajectories.json from the CPU oracle.

The reference ships no golden vectors and cannot be compiled here (no Eigen), so these are
REGRESSION ANCHORS produced by the oracle itself (parity unpinned; see oracle/sp_oracle.cpp).
They agree with the indicative trajectories recorded independently in SURVEY.md section 4
(691/692, 692/693, 98/99 iterations/factorizations).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import sp_oracle as orc  # noqa: E402

cases = {
    "rosenbrock_reference": dict(kind="rosenbrock", iparam0=1, iparam1=0, maxIterations=1000),
    "rosenbrock_c1_n1000": dict(kind="rosenbrock", iparam0=500, iparam1=0, maxIterations=1000),
    "rosenbrock_mgh_start": dict(kind="rosenbrock", iparam0=500, iparam1=0, maxIterations=1000, x0=list(np.tile([-1.2, 1.0], 500))),
    "lf_g32": dict(kind="lf", iparam0=32, iparam1=2, maxIterations=100),
}
out = {}
for name, c in cases.items():
    r = orc.lm_solve(c["kind"], iparam0=c["iparam0"], iparam1=c["iparam1"], maxIterations=c["maxIterations"],
                     x0=np.array(c["x0"]) if c.get("x0") else None)
    out[name] = dict(c, exit_path=r.exit_path, currIter=r.currIter, factorizations=r.factorizations,
                     energy=r.energy, **{"lambda": list(r.trace["lambda"]), "newEnergy2_head": list(r.trace["newEnergy2"][:8])})
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "trajectories.json"), "w") as f:
    json.dump(out, f)
print({k: (v["currIter"], v["factorizations"], v["exit_path"]) for k, v in out.items()})
