"""Deep-run parity at mid size (run under gpurun): the engine against the oracle over thousands of pivots, far enough for
the basis to hold more than 512 structural columns and the eta file more than 512 columns, i.e. past the limits of the
fused chain (production thresholds, not the MLP_FUSED_MAX test knob).  Prints one JSON line per case."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import minilp_b200 as mb
import oracle

cases = [(1, 1500, 2500, 3, 2500), (3, 1800, 1800, 2, 2500), (0, 2500, 2500, 1, 2500), (2, 1500, 2000, 5, 2500)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in c.split(",")) for c in sys.argv[1:]]
for kind, m, n, seed, budget in cases:
    lp = mb.synth_dense(kind, m, n, seed, threads=os.cpu_count() or 1)
    s = mb.Solver.from_dense(lp)
    t0 = time.perf_counter()
    dg = s.run(budget)
    s.engine.sync()
    tg = time.perf_counter() - t0
    ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
    t0 = time.perf_counter()
    dr = ref.continue_solve(budget)
    tc = time.perf_counter() - t0
    a, b = s.trace(), ref.trace()
    kk = min(a.shape[0], b.shape[0])
    same = np.all(a[:kk, :5] == b[:kk, :5], axis=1)
    c = s.engine.counters()
    rel = abs(s.cur_obj_val - ref.cur_obj_val) / max(1.0, abs(ref.cur_obj_val))
    print(json.dumps({"kind": kind, "m": m, "n": n, "seed": seed, "gpu_pivots": int(a.shape[0]), "cpu_pivots": int(b.shape[0]),
                      "done": [bool(dg), bool(dr)], "first_divergence": int(np.argmin(same)) if not same.all() else -1,
                      "obj_rel_diff": rel, "k_structural": c["k_structural"], "max_eta_count": int(a[:, 8].max()) if a.shape[0] else 0,
                      "refactors": c["refactors"], "oracle_ties": ref.tie_events, "gpu_pivots_per_s": a.shape[0] / tg,
                      "cpu_pivots_per_s": b.shape[0] / tc}), flush=True)
    s.close()
