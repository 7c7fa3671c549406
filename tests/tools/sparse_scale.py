"""BASELINE config 4 at scale: netlib-like sparse LP via MPS on one B200 vs the oracle port on one host core.
   python tests/tools/sparse_scale.py [--m 100000 --n 100000 --col-nnz 100 --pivots 2000]
Prints one JSON line (not a bench.py line: config 4 is a parity case, this is its measured footnote)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import minilp_b200 as mb
from minilp_b200 import mps, synth
import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=100000); ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--col-nnz", type=float, default=100.0); ap.add_argument("--pivots", type=int, default=2000)
ap.add_argument("--seed", type=int, default=1); ap.add_argument("--family", default="netlib_like")
ap.add_argument("--cpu-seconds", type=float, default=60.0)
a = ap.parse_args()
t0 = time.time()
text, d = getattr(synth, a.family)(a.m, a.n, a.col_nnz, a.seed)
t1 = time.time()
p = mps.MpsFile.parse(text, d).problem
t2 = time.time()
rp, ci, va, ops, rhs = p.to_csr()
s = mb.Solver(len(ops), len(p.obj_coeffs), csr=(rp, ci, va))
s.init(np.array(p.obj_coeffs), np.array(p.var_mins), np.array(p.var_maxs), ops, rhs)
s.set_record_trace(True)
t3 = time.time()
e = s.engine
s.run(20)
e.sync(); c0 = e.counters(); w0 = time.perf_counter(); e.event_mark(0)
done = s.run(a.pivots)
e.event_mark(1); e.sync(); w1 = time.perf_counter()
piv = s.pivots_done - 20
c1 = e.counters()
out = {"workload": f"{a.family} {a.m}x{a.n} nnz {len(va)} seed {a.seed} via MPS", "gen_s": round(t1 - t0, 2), "mps_parse_s": round(t2 - t1, 2),
       "setup_s": round(t3 - t2, 2), "mps_bytes": len(text), "gpu_pivots": piv, "gpu_ms_per_pivot_device": e.event_elapsed_ms(0, 1) / max(piv, 1),
       "gpu_pivots_per_s_e2e": piv / (w1 - w0), "done": bool(done), "k_structural": c1["k_structural"], "refactors": c1["refactors"] - c0["refactors"],
       "launches_per_pivot": (c1["kernel_launches"] - c0["kernel_launches"]) / max(piv, 1), "obj": s.cur_obj_val}
tg = s.trace()
# oracle on the same text, same pivot budget (time-capped)
t4 = time.time()
op = oracle.MpsFile.parse(text, d).problem
ref = op.init_only()
t5 = time.time()
ref.continue_solve(20)
n_ref, sec = 0, 0.0
while n_ref < a.pivots and sec < a.cpu_seconds:
    fin, dt = ref.continue_timed(min(50, a.pivots - n_ref))
    sec += dt; n_ref = ref.pivots_done - 20
    if fin: break
tr = ref.trace()
k = min(tg.shape[0], tr.shape[0])
same = np.all(tg[:k, :5] == tr[:k, :5], axis=1)
out.update({"cpu_setup_s": round(t5 - t4, 2), "cpu_pivots": n_ref, "cpu_pivots_per_s": n_ref / sec if sec else 0.0, "cpu_cores": 1,
            "compared_pivots": int(k), "first_divergence": int(np.argmin(same)) if not same.all() else -1, "oracle_ties": ref.tie_events})
print(json.dumps(out))
