"""Generates tests/golden/fullsize_*.npz: the ORACLE's first pivots on the LPs the benchmark numbers are quoted on
(BASELINE configs 3, 4 and a config-5-shaped LP), so that the driver-run GPU suite compares the engine with the oracle at
those sizes without running minutes of oracle on the GPU box.  The oracle runs with the reference's own tie rule.

  python tests/golden/make_fullsize_traces.py [cfg3] [cfg4] [cfg5]      (repo root; CPU only; needs ~30 GB of host RAM)

  cfg3  dense_pos 50 000 x 50 000 seed 1 (the bench.py workload), first 24 pivots                      ~4 min
  cfg4  netlib_like 100 000 x 100 000, 0.1 % non-zeros, seed 1, through MPS text, first 3 000 pivots    ~6 min
  cfg4opt  netlib_like 8 000 x 8 000 (the same 0.1 % density), seed 1, through MPS text, TO THE OPTIMUM: 18 058 pivots   ~3 min
  cfg5  dense_pos 30 000 x 120 000 seed 1 — config 5's 1:4 shape at the largest size this host (62 GB) holds in the
        oracle's dense storage; first 12 pivots                                                         ~5 min
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import oracle

OUT = os.path.join(ROOT, "tests", "golden")
which = set(sys.argv[1:]) or {"cfg3", "cfg4", "cfg5"}


def save(name, meta, ref, budget, done, secs):
    tr = ref.trace()
    path = os.path.join(OUT, name)
    np.savez_compressed(path, seq=tr[:, :5].astype(np.int32), obj=tr[:, 7].copy(), pivot_coeff=tr[:, 5].copy(),
                        eta_count=tr[:, 8].astype(np.int32), lu_nnz=tr[:, 9].astype(np.int64), refactored=tr[:, 12].astype(np.int8),
                        budget=budget, done=done, tied_pivots=ref.tied_pivots, near_tie_pivots=ref.near_tie_pivots,
                        first_near_tie_pivot=ref.first_near_tie_pivot, tie_events=ref.tie_events, oracle_seconds=secs, **meta)
    print(path, "pivots", tr.shape[0], "done", done, "tied", ref.tied_pivots, "near", ref.near_tie_pivots, f"{secs:.0f}s",
          os.path.getsize(path), "bytes", flush=True)


def dense(name, kind, m, n, seed, budget):
    t0 = time.perf_counter()
    ref = oracle.DenseSolver.synth(kind, m, n, seed, threads=os.cpu_count() or 1)
    print(name, "generated + try_new", f"{time.perf_counter() - t0:.0f}s", flush=True)
    t0 = time.perf_counter()
    done = ref.continue_solve(budget)
    save(name, dict(kind=kind, m=m, n=n, seed=seed), ref, budget, done, time.perf_counter() - t0)


if "cfg3" in which:
    dense("fullsize_cfg3_dense_pos_50000x50000_s1.npz", 0, 50000, 50000, 1, 24)
if "cfg5" in which:
    dense("fullsize_cfg5_dense_pos_30000x120000_s1.npz", 0, 30000, 120000, 1, 12)
if "cfg4opt" in which:
    # the same family, same 0.1 % density, at a size the oracle SOLVES: 8 000 x 8 000 to the optimum (18 058 pivots, ~3 min)
    from minilp_b200 import synth
    m = n = 8000
    text, d = synth.netlib_like(m, n, 8.0, 1)
    ref = oracle.MpsFile.parse(text, d).problem.init_only()
    t0 = time.perf_counter()
    done = ref.continue_solve(-1)
    assert done and ref.sel_near_tie_pivots == 0
    save("fullsize_cfg4opt_netlib_like_8000x8000_s1.npz", dict(m=m, n=n, seed=1, col_nnz=8.0), ref, -1, done, time.perf_counter() - t0)
if "cfg4" in which:
    from minilp_b200 import synth
    m = n = 100000
    t0 = time.perf_counter()
    text, d = synth.netlib_like(m, n, 100.0, 1)
    ref = oracle.MpsFile.parse(text, d).problem.init_only()
    print("cfg4 generated + parsed + try_new", f"{time.perf_counter() - t0:.0f}s", flush=True)
    t0 = time.perf_counter()
    done = ref.continue_solve(3000)
    save("fullsize_cfg4_netlib_like_100000x100000_s1.npz", dict(m=m, n=n, seed=1, col_nnz=100.0), ref, 3000, done,
         time.perf_counter() - t0)
