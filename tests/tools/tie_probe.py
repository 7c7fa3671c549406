"""Probe (not a test): small LPs with small-integer data, where exact ties in the pricing and ratio tests are common
("measure-zero on dense random data but common on +-1-structured LPs").

Three solvers per instance: the oracle with the REFERENCE's tie rule (first in list order), the oracle with the engine's
rule (lowest index), and the engine.  Reports, per instance, whether a ratio-test winner was contested (exactly / within
1e-9), whether the engine follows the reference-rule oracle, whether it follows the lowest-index oracle, and — where it
leaves even that one — WHY: the state of both at the diverging pivot is replayed and the deciding quantities are printed
(pricing scores of the two entering candidates, or |alpha| / step of the two leaving candidates).
Run under gpurun; prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import minilp_b200 as mb
import oracle

N_INST = int(sys.argv[1]) if len(sys.argv) > 1 else 40


def make(rng, inst):
    m, n = int(rng.integers(8, 28)), int(rng.integers(10, 36))
    a = rng.integers(-2, 3, size=(m, n)).astype(float)
    a[rng.random((m, n)) < 0.4] = 0.0
    a[:, a.any(axis=0) == 0] = 1.0
    a[a.any(axis=1) == 0, :] = 1.0
    obj = rng.integers(-3, 4, size=n).astype(float)
    lo = np.zeros(n)
    hi = rng.integers(1, 6, size=n).astype(float)
    x0 = np.array([rng.integers(0, h + 1) for h in hi], dtype=float)
    ops = rng.integers(0, 3, size=m).astype(np.int32)  # 0 Eq, 1 Le, 2 Ge
    act = a @ x0
    rhs = np.where(ops == 1, act + rng.integers(0, 3, size=m), np.where(ops == 2, act - rng.integers(0, 3, size=m), act)).astype(float)
    d = mb.OptimizationDirection.Minimize if inst % 2 else mb.OptimizationDirection.Maximize
    return mb.DenseLP(d, a, obj, lo, hi, ops, rhs)


def run_oracle(lp, lowest):
    ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs, tie_lowest_index=lowest)
    try:
        ref.continue_solve()
        return ref, "ok"
    except Exception as exc:  # Infeasible / Unbounded
        return ref, type(exc).__name__


def explain(lp, p, tg_row, tr_row):
    """Replay both to just before pivot p and show what decided it."""
    g = mb.Solver.from_dense(lp)
    r = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs, tie_lowest_index=True)
    if p > 0:
        g.run(p)
        r.continue_solve(p)
    out = {"pivot": p, "phase": int(tr_row[0]), "gpu": [int(x) for x in tg_row[:5]], "oracle": [int(x) for x in tr_row[:5]]}
    nb = g.nb_vars()
    if int(tg_row[0]) == 1 and int(tg_row[1]) != int(tr_row[1]):  # primal pricing chose another column
        dg, gg = g.nb_var_obj_coeffs(), g.primal_edge_sq_norms()
        dr, gr = np.asarray(r.nb_var_obj_coeffs), np.asarray(r.primal_edge_sq_norms)
        cg, cr = int(tg_row[2]), int(tr_row[2])
        pse = bool(r.enable_primal_steepest_edge)
        sc = (lambda d_, g_, c: float(d_[c] * d_[c] / g_[c]) if pse else float(abs(d_[c])))
        out["what"] = "pricing"
        out["gpu_scores(gpu col, oracle col)"] = [sc(dg, gg, cg), sc(dg, gg, cr)]
        out["oracle_scores(gpu col, oracle col)"] = [sc(dr, gr, cg), sc(dr, gr, cr)]
        a, b = out["oracle_scores(gpu col, oracle col)"]
        out["rel_gap_in_oracle"] = abs(a - b) / max(abs(a), abs(b), 1e-300)
    elif int(tg_row[0]) == 0 and int(tg_row[3]) != int(tr_row[3]):  # dual row selection
        xb_g, w_g = g.basic_var_vals(), g.dual_edge_sq_norms()
        xb_r, w_r = np.asarray(r.basic_var_vals), np.asarray(r.dual_edge_sq_norms)
        lo_b, hi_b = g.engine.download(8), g.engine.download(9)

        def score(xb, w, row):
            v = xb[row]
            inf = lo_b[row] - v if v < lo_b[row] - 1e-8 else (v - hi_b[row] if v > hi_b[row] + 1e-8 else 0.0)
            return float(inf * inf / w[row])
        rg, rr = int(tg_row[3]), int(tr_row[3])
        out["what"] = "dual row"
        out["gpu_scores(gpu row, oracle row)"] = [score(xb_g, w_g, rg), score(xb_g, w_g, rr)]
        out["oracle_scores(gpu row, oracle row)"] = [score(xb_r, w_r, rg), score(xb_r, w_r, rr)]
        a, b = out["oracle_scores(gpu row, oracle row)"]
        out["rel_gap_in_oracle"] = abs(a - b) / max(abs(a), abs(b), 1e-300)
    else:
        out["what"] = "ratio test (leaving row of the primal loop / entering column of the dual loop)"
        out["gpu_pivot_coeff"], out["oracle_pivot_coeff"] = float(tg_row[5]), float(tr_row[5])
        out["rel_gap_abs_coeff"] = abs(abs(tg_row[5]) - abs(tr_row[5])) / max(abs(tg_row[5]), abs(tr_row[5]), 1e-300)
    g.close()
    return out


rng = np.random.default_rng(7)
res = []
for inst in range(N_INST):
    lp = make(rng, inst)
    rec = {"inst": inst, "m": int(lp.a.shape[0]), "n": int(lp.a.shape[1])}
    ref, rec["oracle_reference_rule"] = run_oracle(lp, False)
    low, rec["oracle_lowest_index"] = run_oracle(lp, True)
    try:
        g = mb.Solver.from_dense(lp)
        g.run()
        rec["gpu"] = "ok"
    except Exception as exc:
        rec["gpu"] = type(exc).__name__
    tg = g.trace()
    st = g.tie_stats()
    rec["gpu_ties"] = st
    rec["oracle_ties"] = {"tied_pivots": ref.tied_pivots, "near_tie_pivots": ref.near_tie_pivots,
                          "first_tied_pivot": ref.first_tied_pivot, "first_near_tie_pivot": ref.first_near_tie_pivot}
    for name, o in (("reference_rule", ref), ("lowest_index", low)):
        tr = o.trace()
        k = min(tg.shape[0], tr.shape[0])
        same = np.all(tg[:k, :5] == tr[:k, :5], axis=1)
        div = -1 if (same.all() and tg.shape[0] == tr.shape[0]) else (int(np.argmin(same)) if not same.all() else k)
        rec[f"first_divergence_vs_{name}"] = div
    contested = ref.first_near_tie_pivot
    dref = rec["first_divergence_vs_reference_rule"]
    rec["follows_reference_where_uncontested"] = bool(dref < 0 or (contested >= 0 and dref >= contested))
    if rec["gpu"] == "ok" and rec["oracle_lowest_index"] == "ok":
        rec["obj_equal"] = bool(abs(g.cur_obj_val - low.cur_obj_val) <= 1e-8 * max(1.0, abs(low.cur_obj_val)))
    dl = rec["first_divergence_vs_lowest_index"]
    if dl >= 0 and dl < min(tg.shape[0], low.trace().shape[0]):
        rec["why"] = explain(lp, dl, tg[dl], low.trace()[dl])
    res.append(rec)
    g.close()
print(json.dumps({
    "instances": len(res),
    "status_agree": sum(r["gpu"] == r["oracle_reference_rule"] for r in res),
    "contested_exact": sum(r["oracle_ties"]["tied_pivots"] > 0 for r in res),
    "contested_near": sum(r["oracle_ties"]["near_tie_pivots"] > 0 for r in res),
    "follows_reference_where_uncontested": sum(r["follows_reference_where_uncontested"] for r in res),
    "same_sequence_as_reference_rule": sum(r["first_divergence_vs_reference_rule"] < 0 for r in res),
    "same_sequence_as_lowest_index": sum(r["first_divergence_vs_lowest_index"] < 0 for r in res),
    "gpu_first_tie_matches_oracle": sum(r["gpu_ties"]["first_tied_pivot"] == r["oracle_ties"]["first_tied_pivot"] for r in res
                                        if r["first_divergence_vs_reference_rule"] < 0 or r["first_divergence_vs_reference_rule"] >= max(r["oracle_ties"]["first_tied_pivot"], 0)),
    "obj_equal": sum(bool(r.get("obj_equal")) for r in res), "detail": res}))
