"""BASELINE config 2: 1000 x 1000 dense LP, FTRAN / BTRAN / price kernels only — device time per call (CUDA events
through the ABI's marks, median of 50) next to the oracle port's probes on one host core."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import minilp_b200 as mb
import oracle

m = n = 1000
lp = mb.synth_dense(0, m, n, 1)
gpu = mb.Solver.from_dense(lp)
ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
gpu.run(60); ref.continue_solve(60)
e = gpu.engine
nb = gpu.nb_vars()

def dev(f, reps=50):
    ts = []
    for i in range(reps):
        e.sync(); e.event_mark(0); f(i); e.event_mark(1); e.sync()
        ts.append(e.event_elapsed_ms(0, 1) * 1e3)
    return float(np.median(ts))

def cpu(f, reps=20):
    ts = []
    for i in range(reps):
        t = time.perf_counter(); f(i); ts.append((time.perf_counter() - t) * 1e6)
    return float(np.median(ts))

cols = [int(c) for c in np.random.default_rng(0).integers(0, n, 64)]
rows = [int(r) for r in np.random.default_rng(1).integers(0, m, 64)]
out = {"workload": "dense_pos 1000x1000 seed 1 after 60 pivots", "counters": e.counters(),
       "gpu_us": {"ftran_col": dev(lambda i: e.ftran_col(int(nb[cols[i % 64]]))),
                  "btran_unit": dev(lambda i: e.btran_unit(rows[i % 64])),
                  "price_row": dev(lambda i: e.price_row()),
                  "calc_row_coeffs": dev(lambda i: e.calc_row_coeffs(rows[i % 64]))},
       "cpu_us_1core": {"ftran_col": cpu(lambda i: ref.probe_ftran_col(cols[i % 64])),
                        "calc_row_coeffs": cpu(lambda i: ref.probe_row_coeffs(rows[i % 64]))}}
ms, by = e.bench_price_dense(20)
out["price_dense_all_rows"] = {"us": ms * 1e3, "GBps": by / (ms * 1e-3) / 1e9, "note": "8 MB matrix: L2-resident, not an HBM number"}
t = time.perf_counter(); dg = gpu.run(); e.sync(); tg = time.perf_counter() - t
pg = gpu.pivots_done - 60
t = time.perf_counter(); ref.continue_solve(); tc = time.perf_counter() - t
pc = ref.pivots_done - 60
out["full_solve"] = {"gpu_pivots": pg, "gpu_pivots_per_s": pg / tg, "cpu_pivots": pc, "cpu_pivots_per_s": pc / tc,
                     "same_sequence": bool(np.array_equal(gpu.trace()[:, :5], ref.trace()[:, :5]))}
print(json.dumps(out))
