"""ms/pivot as the basis fills up (BASELINE config 3, run under gpurun): the 50k x 50k LP from the slack basis towards the
optimum in segments, one JSON line per segment {pivots_done, k, K_end, ms_per_pivot (CUDA events), share of the dense N^T v
price-out, its GB/s, refactorizations} and a final line with the totals.
   python scripts/deep_curve.py [--segment 500] [--max-pivots 40000] [--m 50000 --n 50000]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=50000); ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--kind", type=int, default=0); ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--segment", type=int, default=500); ap.add_argument("--max-pivots", type=int, default=40000)
ap.add_argument("--max-seconds", type=float, default=400.0)
ap.add_argument("--skip", type=int, default=0, help="pivots to run before the first measured segment")
ap.add_argument("--workload", default="dense", choices=["dense", "netlib_like", "sparse_pos"])
ap.add_argument("--col-nnz", type=float, default=100.0)
ap.add_argument("--refactor-factor", type=float, default=1.0)
a = ap.parse_args()
if a.workload == "dense":
    s, setup = bench.build_solver(a, 0)
else:  # BASELINE config 4: through MPS text
    import numpy as np
    import minilp_b200 as mb
    from minilp_b200 import mps
    t0 = time.perf_counter()
    text, d = bench.sparse_text(a)
    p = mps.MpsFile.parse(text, d).problem
    rp, ci, va, ops, rhs = p.to_csr()
    s = mb.Solver(len(ops), len(p.obj_coeffs), csr=(rp, ci, va))
    s.init(np.array(p.obj_coeffs), np.array(p.var_mins), np.array(p.var_maxs), ops, rhs)
    setup = {"total_s": round(time.perf_counter() - t0, 2), "nnz": int(len(va))}
s.set_refactor_factor(a.refactor_factor)
e = s.engine
s.set_record_trace(True)
def infeasible_rows(e):
    """rows whose basic variable is outside its bounds by more than EPS (what choose_pivot_row_dual still has to repair)"""
    import numpy as np
    xb, lo, hi = e.download(3), e.download(8), e.download(9)
    return int(np.sum((xb < lo - 1e-8) | (xb > hi + 1e-8)))


if a.skip > 0:
    s.run(a.skip)
t_start = time.perf_counter()
tot_ms = 0.0
done = False
while not done and s.pivots_done < a.max_pivots and time.perf_counter() - t_start < a.max_seconds:
    c0 = e.counters(); p0 = s.pivots_done; rf0 = s.timers()[1]
    e.profile_enable(True)
    e.sync(); w0 = time.perf_counter(); e.event_mark(0)
    done = s.run(a.segment)
    e.event_mark(1); e.sync(); w1 = time.perf_counter()
    ms = e.event_elapsed_ms(0, 1)
    e.profile_enable(False)
    pr = e.profile(); c1 = e.counters()
    piv = s.pivots_done - p0
    tot_ms += ms
    nv = max(pr["price_v_launches"], 1)
    print(json.dumps({"pivots_done": s.pivots_done, "segment_pivots": piv, "k": c1["k_structural"], "K_end": c1["eta_count"],
                      "ms_per_pivot": ms / max(piv, 1), "wall_ms_per_pivot": (w1 - w0) * 1e3 / max(piv, 1),
                      "price_v_ms_per_launch": pr["price_v_ms"] / nv, "price_v_share": pr["price_v_ms"] / ms if ms else 0,
                      "price_v_GBps": pr["price_v_bytes"] / max(pr["price_v_ms"], 1e-9) / 1e6,
                      "price_rho_ms_per_launch": pr["price_rho_ms"] / max(pr["price_rho_launches"], 1),
                      "other_ms_per_pivot": (ms - pr["price_v_ms"]) / max(piv, 1),
                      "launches_per_pivot": (c1["kernel_launches"] - c0["kernel_launches"]) / max(piv, 1),
                      "refactors": c1["refactors"] - c0["refactors"], "refactor_wall_ms_per_pivot": (s.timers()[1] - rf0) * 1e3 / max(piv, 1),
                      "lu_nnz": c1["lu_nnz"], "obj": s.cur_obj_val, "primal_infeasible_rows": infeasible_rows(e), "done": bool(done)}), flush=True)
print(json.dumps({"summary": True, "workload": bench.workload_name(a), "pivots": s.pivots_done, "optimal": bool(done),
                  "device_seconds": tot_ms / 1e3, "objective": s.cur_obj_val, "ties": s.tie_stats(), "setup": setup}), flush=True)
s.close()
