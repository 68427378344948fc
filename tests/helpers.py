"""Shared generators for the parity tests (numpy side; no product code here)."""
import numpy as np
import scipy.sparse as sp


def random_jacobian(rng, m, n, nnz_per_row, explicit_zero_frac=0.0, empty_cols=(), empty_rows=(), cover_cols=False):
    """Random sparse m x n CSC pattern+values with optional explicit zeros / empty rows / columns.
    cover_cols: give every allowed column at least one entry (so J^T J + D is positive definite)."""
    rows, cols = [], []
    allowed = np.setdiff1d(np.arange(n), np.asarray(empty_cols, int))
    seen = np.zeros(n, bool)
    for r in range(m):
        if r in empty_rows:
            continue
        k = min(len(allowed), max(1, int(rng.integers(1, 2 * nnz_per_row))))
        c = rng.choice(allowed, size=k, replace=False)
        rows += [r] * k
        cols += list(c)
        seen[c] = True
    if cover_cols:
        for c in allowed[~seen[allowed]]:
            r = int(rng.integers(0, m))
            while r in empty_rows:
                r = int(rng.integers(0, m))
            rows.append(r)
            cols.append(int(c))
    rows = np.array(rows, np.int32)
    cols = np.array(cols, np.int32)
    vals = rng.standard_normal(len(rows))
    if explicit_zero_frac > 0:
        vals[rng.random(len(vals)) < explicit_zero_frac] = 0.0
    order = np.lexsort((rows, cols))
    rows, cols, vals = rows[order], cols[order], vals[order]
    colptr = np.zeros(n + 1, np.int32)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr).astype(np.int32)
    return colptr, rows.astype(np.int32), vals


def splitmix64_at(seed, ctr):
    """numpy twin of the counter-based generator used by the LF workload."""
    M = (1 << 64) - 1
    z = (seed + (ctr + 1) * 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def lf_matrix_numpy(g, rows_mult, seed):
    """Pure-Python/numpy construction of the LF Jacobian (independent of both C++ generators)."""
    n = g * g
    m = rows_mult * n
    dy = [-1, -1, -1, 0, 0, 1, 1, 1]
    dx = [-1, 0, 1, -1, 1, -1, 0, 1]
    R, Cc, V = [], [], []
    for r in range(m):
        c0 = r % n
        y0, x0 = divmod(c0, g)
        h0 = splitmix64_at(seed, r * 8)
        drop = h0 & 7
        u = lambda h: 2.0 * ((h >> 11) * (1.0 / 9007199254740992.0)) - 1.0
        R.append(r); Cc.append(c0); V.append(2.0 + u(h0))
        s = 1
        for k in range(8):
            if k == drop:
                continue
            R.append(r); Cc.append(((y0 + dy[k]) % g) * g + (x0 + dx[k]) % g); V.append(u(splitmix64_at(seed, r * 8 + s)))
            s += 1
    A = sp.coo_matrix((V, (R, Cc)), shape=(m, n)).tocsc()
    A.sort_indices()
    return A
