"""Incremental re-analysis at the C3 stand-in size: time spb_analyze (from scratch) against
spb_analyze_update after ONE constraint row is appended to the constant block of the seamless-integration Jacobian
(IterativeRoundingTraits.h:76-79; the barrier rows below move down by one), and check that the patched handle
assembles the same bits as a freshly analysed one.

    python scripts/analyze_update_bench.py [nx ny [rounds]]      (500 500 = 498 002 faces, 4.48 M unknowns)
"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import saddlepoint_b200 as spb
import c3_blocks

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 120
ny = int(sys.argv[2]) if len(sys.argv) > 2 else nx
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3


def insert_row(m, n, colptr, rowidx, rho, c):
    """CSC pattern with one single-entry row (column c) inserted before old row rho."""
    q0, q1 = colptr[c], colptr[c + 1]
    pos = q0 + int(np.searchsorted(rowidx[q0:q1], rho))
    new_rows = np.empty(len(rowidx) + 1, np.int32)
    new_rows[:pos] = rowidx[:pos]
    new_rows[pos + 1:] = rowidx[pos:]
    new_rows += (new_rows >= rho).astype(np.int32)  # (the slot at pos is overwritten below)
    new_rows[pos] = rho
    new_cp = colptr.copy()
    new_cp[c + 1:] += 1
    return m + 1, n, new_cp, new_rows, pos


P = c3_blocks.build(nx, ny)
h = spb.Handle(0)
h.set_solver(spb.SOLVER_CHOLESKY, 1e-13, 10000)
prob = spb.BuiltinProblem(h, "blocks", **{k: P[k] for k in ("A", "b", "bar_idx", "bar_vol", "s", "w_bar", "x0")})
m, n, colptr, rowidx = prob.pattern()
colptr = np.ascontiguousarray(colptr, np.int32)
rowidx = np.ascontiguousarray(rowidx, np.int32)
n_const = P["A"].shape[0]
print("C3 stand-in %dx%d: %d faces, n=%d, m=%d (constant block %d rows), nnzJ=%d" % (nx, ny, P["nF"], n, m, n_const, colptr[-1]), flush=True)
t0 = time.time()
h.analyze(m, n, colptr, rowidx)
t_full = time.time() - t0
rng = np.random.default_rng(1)
vals = rng.standard_normal(int(colptr[-1]))
E = rng.standard_normal(m)
h.assemble(vals, E, 0.01)
t0 = time.time()
ok = h.factorize()
h.solve(h.rhs())
t_first_factor = time.time() - t0
print("from-scratch analysis %.3f s; first factorisation incl. ordering / supernodes %.3f s (ok=%s)" % (t_full, t_first_factor, ok), flush=True)

out = {"workload": "C3 stand-in %dx%d, %d faces, n=%d, m=%d, nnzJ=%d" % (nx, ny, P["nF"], n, m, int(colptr[-1])),
       "full_analysis_s": t_full, "first_factorisation_incl_symbolic_s": t_first_factor, "updates": []}
pin = rng.choice(n, size=rounds, replace=False)
for r in range(rounds):
    c = int(pin[r])
    m, n, colptr, rowidx, pos = insert_row(m, n, colptr, rowidx, n_const, c)  # appended to the constraint block
    n_const += 1
    vals = np.insert(vals, pos, 1.0)
    E = np.insert(E, n_const - 1, 0.25)
    t0 = time.time()
    how = h.analyze_update(m, n, colptr, rowidx)
    t_upd = time.time() - t0
    h.assemble(vals, E, 0.01)
    t0 = time.time()
    ok = h.factorize()
    x, _, _ = h.solve(h.rhs())
    t_fac = time.time() - t0
    rec = {"pinned_column": c, "how": how, "update_s": t_upd, "factor_and_solve_after_update_s": t_fac, "factor_ok": bool(ok)}
    if r == rounds - 1:
        # referee: a fresh handle analysed from scratch assembles the same bits
        hv = h.h_values().view(np.uint64)
        g = h.rhs().view(np.uint64)
        d = h.damping().view(np.uint64)
        h2 = spb.Handle(0)
        t0 = time.time()
        h2.analyze(m, n, colptr, rowidx)
        rec["full_analysis_of_the_grown_pattern_s"] = time.time() - t0
        h2.assemble(vals, E, 0.01)
        cp1, ri1 = h.h_pattern()
        cp2, ri2 = h2.h_pattern()
        rec["h_pattern_equal"] = bool((cp1 == cp2).all() and (ri1 == ri2).all())
        rec["h_values_bit_equal"] = bool((hv == h2.h_values().view(np.uint64)).all())
        rec["rhs_bit_equal"] = bool((g == h2.rhs().view(np.uint64)).all())
        rec["damping_bit_equal"] = bool((d == h2.damping().view(np.uint64)).all())
        assert rec["h_pattern_equal"] and rec["h_values_bit_equal"] and rec["rhs_bit_equal"] and rec["damping_bit_equal"]
        h2.close()
    out["updates"].append(rec)
    print(json.dumps(rec), flush=True)
print("ANALYZE_UPDATE_JSON " + json.dumps(out))
