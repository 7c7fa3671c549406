"""Synthetic sparse LPs in free-format MPS (BASELINE config 4: "Netlib-shaped sparse LP via MPS").

The reference ships no generator: these are this repo's workload definitions.  They are emitted as MPS
TEXT so that both sides of a parity test — the oracle's restated parser and `minilp_b200.mps` — enter through the
MPS path the config names.  numpy only; every number is written with repr() (shortest round-trip form).

  netlib_like(m, n, ...)  Minimize c x, rows 50 % L / 25 % G / 25 % E, power-law column counts, 60 % of the coefficients
                          +-1 and the rest lognormal, 30 % of the variables with a finite upper bound, a few negative
                          costs on bounded variables, a few RANGES; rhs = A x0 + slack for a hidden feasible x0.
                          Starts dual feasible and primal infeasible: the dual simplex loop (restore_feasibility).
  sparse_pos(m, n, ...)   Maximize c x, A x <= b, A >= 0, x >= 0: the primal loop with primal steepest edge — the sparse
                          counterpart of dense_pos.
"""
import numpy as np


def _column_counts(rng, m, n, mean, cmin=2):
    """Power-law (Pareto, shape 1.5) column counts with the requested mean, clipped to [cmin, m]."""
    mean = max(float(mean), cmin + 0.5)
    raw = (rng.pareto(1.5, size=n) + 1.0)  # mean 3
    cnt = cmin + (raw - 1.0) * (mean - cmin) / 2.0
    return np.clip(np.rint(cnt).astype(np.int64), cmin, m)


def _pattern(rng, m, n, mean_col_nnz):
    cnt = _column_counts(rng, m, n, mean_col_nnz)
    cols = np.repeat(np.arange(n), cnt)
    rows = np.empty(cols.shape[0], dtype=np.int64)
    off = 0
    for j in range(n):
        c = int(cnt[j])
        rows[off:off + c] = rng.choice(m, size=c, replace=False) if c * 4 > m else _distinct(rng, m, c)
        off += c
    # every row needs at least one entry (an empty row is dropped by try_new and would shift the row numbering)
    missing = np.setdiff1d(np.arange(m), rows)
    if missing.size:
        rows = np.concatenate([rows, missing])
        cols = np.concatenate([cols, rng.integers(0, n, size=missing.size)])
    order = np.lexsort((rows, cols))
    rows, cols = rows[order], cols[order]
    keep = np.ones(rows.shape[0], dtype=bool)
    keep[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    return rows[keep], cols[keep]


def _distinct(rng, m, c):
    out = np.unique(rng.integers(0, m, size=c))
    while out.size < c:
        out = np.unique(np.concatenate([out, rng.integers(0, m, size=c - out.size)]))
    return rng.permutation(out)


def _emit(name, m, n, rows, cols, vals, obj, row_types, rhs, ranges, ubs, lbs=None):
    out = [f"NAME {name}", "ROWS", " N COST"]
    out += [f" {row_types[i]} R{i}" for i in range(m)]
    out.append("COLUMNS")
    start = np.searchsorted(cols, np.arange(n + 1))
    for j in range(n):
        if obj[j] != 0.0:
            out.append(f" X{j} COST {float(obj[j])!r}")
        for t in range(start[j], start[j + 1]):
            out.append(f" X{j} R{rows[t]} {float(vals[t])!r}")
    out.append("RHS")
    out += [f" RHS R{i} {float(rhs[i])!r}" for i in range(m) if rhs[i] != 0.0]
    if ranges:
        out.append("RANGES")
        out += [f" RNG R{i} {float(r)!r}" for i, r in sorted(ranges.items())]
    bl = []
    for j in range(n):
        if lbs is not None and lbs[j] != 0.0:
            bl.append(f" LO BND X{j} {float(lbs[j])!r}")
        if np.isfinite(ubs[j]):
            bl.append(f" UP BND X{j} {float(ubs[j])!r}")
    if bl:
        out.append("BOUNDS")
        out += bl
    out.append("ENDATA")
    return "\n".join(out) + "\n"


def netlib_like(m, n, mean_col_nnz=8.0, seed=1):
    """Returns (mps_text, direction) with direction 0 = Minimize."""
    rng = np.random.default_rng(seed)
    rows, cols = _pattern(rng, m, n, mean_col_nnz)
    nz = rows.shape[0]
    unit = rng.random(nz) < 0.6
    sign = np.where(rng.random(nz) < 0.5, -1.0, 1.0)
    vals = np.where(unit, 1.0, np.exp(rng.normal(0.0, 1.0, size=nz))) * sign
    vals = np.round(vals, 6)
    vals[vals == 0.0] = 1.0
    ubs = np.where(rng.random(n) < 0.3, np.round(1.0 + 9.0 * rng.random(n), 3), np.inf)
    obj = np.round(0.5 + rng.random(n), 6)
    neg = np.isfinite(ubs) & (rng.random(n) < 0.2)
    obj[neg] = -obj[neg]  # negative cost only where an upper bound keeps the LP bounded
    x0 = np.where(rng.random(n) < 0.5, rng.random(n) * np.where(np.isfinite(ubs), ubs, 5.0), 0.0)
    act = np.zeros(m)
    np.add.at(act, rows, vals * x0[cols])
    kind = rng.random(m)
    row_types = np.where(kind < 0.5, "L", np.where(kind < 0.75, "G", "E"))
    slack = np.round(rng.random(m) * 2.0, 6)
    rhs = np.where(row_types == "L", act + slack, np.where(row_types == "G", act - slack, act))
    rhs = np.round(rhs, 9)
    ranges = {}
    for i in rng.choice(m, size=max(1, m // 50), replace=False):
        if row_types[i] in ("L", "G"):
            ranges[int(i)] = float(np.round(slack[i] + 1.0 + 3.0 * rng.random(), 6))
    return _emit(f"NETLIKE_{m}x{n}_s{seed}", m, n, rows, cols, vals, obj, row_types, rhs, ranges, ubs), 0


def sparse_pos(m, n, mean_col_nnz=8.0, seed=1):
    """Returns (mps_text, direction) with direction 1 = Maximize."""
    rng = np.random.default_rng(seed)
    rows, cols = _pattern(rng, m, n, mean_col_nnz)
    nz = rows.shape[0]
    vals = np.round(np.where(rng.random(nz) < 0.6, 1.0, np.exp(rng.normal(0.0, 0.7, size=nz))), 6)
    vals[vals == 0.0] = 1.0
    obj = np.round(0.5 + rng.random(n), 6)
    rhs = np.round((0.5 + rng.random(m)) * max(1.0, mean_col_nnz * n / m / 4.0), 6)
    return _emit(f"SPPOS_{m}x{n}_s{seed}", m, n, rows, cols, vals, obj, np.full(m, "L"), rhs, {}, np.full(n, np.inf)), 1
