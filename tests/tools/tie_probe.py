"""Probe (not a test): small LPs with small-integer data, where exact ties in the pricing and ratio tests are common
("measure-zero on dense random data but common on +-1-structured LPs").  The oracle runs with the engine's
tie rule (lowest index) and counts exact ties; reports how often the engine's pivot sequence still equals the oracle's and
whether the end state does.  Run under gpurun; prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import minilp_b200 as mb
import oracle

rng = np.random.default_rng(7)
res = []
for inst in range(40):
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
    lp = mb.DenseLP(d, a, obj, lo, hi, ops, rhs)
    rec = {"inst": inst, "m": m, "n": n}
    try:
        ref = oracle.DenseSolver(d, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs, tie_lowest_index=True)
        r_ok = ref.continue_solve()
        rec["oracle"] = "ok"
    except Exception as exc:  # Infeasible / Unbounded
        rec["oracle"] = type(exc).__name__
        ref = None
    try:
        g = mb.Solver.from_dense(lp)
        g.run()
        rec["gpu"] = "ok"
    except Exception as exc:
        rec["gpu"] = type(exc).__name__
        g = None
    if ref is not None and g is not None:
        tg, tr = g.trace(), ref.trace()
        rec["ties"] = int(ref.tie_events)
        rec["pivots"] = [int(tg.shape[0]), int(tr.shape[0])]
        rec["same_sequence"] = bool(tg.shape == tr.shape and np.array_equal(tg[:, :5], tr[:, :5]))
        rec["obj_equal"] = bool(abs(g.cur_obj_val - ref.cur_obj_val) <= 1e-8 * max(1.0, abs(ref.cur_obj_val)))
    res.append(rec)
    if g is not None:
        g.close()
both = [r for r in res if r.get("oracle") == "ok" and r.get("gpu") == "ok"]
print(json.dumps({"instances": len(res), "both_solved": len(both), "status_agree": sum(r["oracle"] == r["gpu"] for r in res),
                  "with_ties": sum(r["ties"] > 0 for r in both), "same_sequence": sum(r["same_sequence"] for r in both),
                  "same_sequence_among_tied": sum(r["same_sequence"] for r in both if r["ties"] > 0),
                  "obj_equal": sum(r["obj_equal"] for r in both), "detail": res}))
