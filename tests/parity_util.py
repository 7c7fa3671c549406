"""Shared assertions of the GPU parity tests.

Tie rule.  Pass 2 of the two Harris ratio tests keeps the FIRST maximal |coeff| in list order (solver.rs:804-823, 982-1002
over ScatteredVec.nonzero, sparse.rs:75-80); the engine keeps the lowest row / variable index.  The oracle runs with the
REFERENCE's rule (tie_lowest_index=False) and records every pivot whose winner was contested — exactly, or within 1e-9
relative, i.e. within the rounding of a re-ordered sum.  Where no pivot was contested the two rules cannot differ and the
engine must take the oracle's sequence pivot for pivot; where one was, the sequences must agree up to that pivot and the
engine must have flagged a contested winner no later than there (it reports the same counts through mlp_leaving.ties /
mlp_dual_entering.ties / mlp_solver_tie_stats).  The two selections (pricing, dual row) follow the same rule on both sides;
the oracle records a pivot whose runner-up scored within 1e-9 of the winner (first_sel_near_tie_pivot) and equality is
required up to there as well — beyond it the order of the two scores is a property of rounding, not of the LP.
"""
import numpy as np

REL = 1e-8


def close(a, b, rel=REL):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    fin = np.isfinite(a) & np.isfinite(b)
    return bool(np.array_equal(np.isfinite(a), np.isfinite(b)) and np.all(np.abs(a - b)[fin] <= rel * scale[fin])
                and np.array_equal(a[~fin], b[~fin]))


def uncontested_prefix(ref, n_pivots):
    """Number of leading pivots of the oracle's run in which no decision was contested: neither a ratio-test winner (pass 2,
    where the rules differ) nor a selection (pricing / dual row: same rule on both sides, but a runner-up within 1e-9 of the
    winner is ordered by the rounding of the sums behind the scores, not by the LP)."""
    firsts = [int(x) for x in (ref.first_near_tie_pivot, ref.first_sel_near_tie_pivot) if x >= 0]
    return n_pivots if not firsts else min(min(firsts), n_pivots)


def assert_sequence_parity(tg, tr, ref=None, gpu=None, cols=(0, 1, 2, 3, 4), values=True):
    """tg / tr: traces (13 doubles per pivot) of the engine and of the oracle (reference tie rule).  ref / gpu: the live
    solvers, for the tie bookkeeping; without `ref` full equality is required."""
    k = tr.shape[0] if ref is None else uncontested_prefix(ref, tr.shape[0])
    contested = k < tr.shape[0]
    if not contested:
        assert tg.shape[0] == tr.shape[0], f"pivot counts differ: gpu {tg.shape[0]} oracle {tr.shape[0]}"
    else:
        assert tg.shape[0] >= k, f"engine stopped after {tg.shape[0]} pivots, before the first contested pivot {k}"
    cols = list(cols)
    seq_g, seq_r = tg[:k, cols], tr[:k, cols]
    if not np.array_equal(seq_g, seq_r):
        bad = int(np.argmax(np.any(seq_g != seq_r, axis=1)))
        raise AssertionError(f"basis sequence diverges at pivot {bad} (uncontested prefix {k}): gpu {tg[bad]} oracle {tr[bad]}")
    if values:
        for col in (5, 6, 7):  # pivot_coeff, entering_diff, obj_after
            assert close(tg[:k, col], tr[:k, col]), f"trace column {col} differs"
    if gpu is not None and ref is not None:
        st = gpu.tie_stats()
        if not contested:
            assert st["tied_pivots"] == 0, f"engine reports exact ties {st} where the oracle met none"
        elif 0 <= ref.first_near_tie_pivot == k:  # contested ratio-test winner: the engine must have flagged it too
            assert 0 <= st["first_near_tie_pivot"] <= k, f"oracle: first contested pivot {k}; engine: {st}"
    return contested
