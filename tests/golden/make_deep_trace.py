"""Generates tests/golden/deep_trace_*.npz: the ORACLE's pivot sequence over thousands of pivots of a mid-size dense LP
(minutes of CPU), so that the GPU suite can check a deep run — basis with more than 512 structural columns, eta file
longer than 512 — in seconds, without running the oracle on the GPU box.
  python tests/golden/make_deep_trace.py            (from the repo root; needs only the CPU)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import minilp_b200 as mb
import oracle

CASES = [(3, 1800, 1800, 2, 5000), (1, 1500, 2500, 3, 6000)]
for kind, m, n, seed, budget in CASES:
    lp = mb.synth_dense(kind, m, n, seed, threads=os.cpu_count() or 1)
    ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
    t0 = time.perf_counter()
    done = ref.continue_solve(budget)
    tr = ref.trace()
    out = os.path.join(ROOT, "tests", "golden", f"deep_trace_k{kind}_{m}x{n}_s{seed}.npz")
    np.savez_compressed(out, kind=kind, m=m, n=n, seed=seed, budget=budget, done=done, seq=tr[:, :5].astype(np.int32),
                        obj_every_100=tr[99::100, 7].copy(), eta_count=tr[:, 8].astype(np.int32), obj_final=ref.cur_obj_val,
                        tie_events=ref.tie_events, k_end=int((np.asarray(ref.basic_vars) < n).sum()))
    print(out, "pivots", tr.shape[0], "done", done, "max eta", int(tr[:, 8].max()), "ties", ref.tie_events,
          f"{time.perf_counter() - t0:.0f}s", os.path.getsize(out), "bytes")
