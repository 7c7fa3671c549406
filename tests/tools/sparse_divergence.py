"""Diagnosis (not a test): where and why a sparse solve leaves the oracle's pivot sequence.
   python tests/tools/sparse_divergence.py netlib_like 500 350 7.0 5
Replays the oracle (reference tie rule) and the engine to the pivot before the first difference and prints the deciding
quantities on both sides: the dual row scores infeas^2 / w of the two rows, or the |coeff| / step of the two entering
candidates, plus the refactorization history up to there (cadence differences only change rounding)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minilp_b200 as mb
import oracle
from minilp_b200 import mps, synth
from test_sparse_gpu import solver_from_problem

fam, m, n, cn, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
text, d = getattr(synth, fam)(m, n, cn, seed)
p = mps.MpsFile.parse(text, d).problem
ref = oracle.MpsFile.parse(text, d).problem.solve()
g = solver_from_problem(p, "sparse")
g.run()
tg, tr = g.trace(), ref.trace()
k = min(len(tg), len(tr))
same = np.all(tg[:k, :5] == tr[:k, :5], axis=1)
out = {"workload": f"{fam} {m}x{n} {cn} seed {seed}", "gpu_pivots": len(tg), "oracle_pivots": len(tr),
       "obj": [g.cur_obj_val, ref.cur_obj_val], "oracle_ties": [ref.tied_pivots, ref.near_tie_pivots], "gpu_ties": g.tie_stats()}
if same.all() and len(tg) == len(tr):
    out["first_divergence"] = -1
else:
    q = int(np.argmin(same)) if not same.all() else k
    out["first_divergence"] = q
    out["gpu_row"], out["oracle_row"] = tg[q].tolist(), tr[q].tolist()
    out["refactor_pivots_gpu"] = np.flatnonzero(tg[:q + 1, 12]).tolist()[-8:]
    out["refactor_pivots_oracle"] = np.flatnonzero(tr[:q + 1, 12]).tolist()[-8:]
    out["nnz_col_mismatch_before"] = int(np.sum(tg[:q, 10] != tr[:q, 10]))
    out["lu_nnz_mismatch_before"] = int(np.sum(tg[:q, 9] != tr[:q, 9]))
    out["first_cadence_difference"] = int(np.argmax(tg[:q, 12] != tr[:q, 12])) if np.any(tg[:q, 12] != tr[:q, 12]) else -1
    g2 = solver_from_problem(p, "sparse")
    r2 = oracle.MpsFile.parse(text, d).problem.solve(max_pivots=q)
    g2.run(q)
    e = g2.engine
    if int(tr[q, 0]) == 0:  # dual loop
        xb_g, w_g, lo, hi = e.download(3), e.download(4), e.download(8), e.download(9)
        xb_r, w_r = np.asarray(r2.basic_var_vals), np.asarray(r2.dual_edge_sq_norms)

        def score(xb, w, row):
            v = xb[row]
            inf = lo[row] - v if v < lo[row] - 1e-8 else (v - hi[row] if v > hi[row] + 1e-8 else 0.0)
            return float(inf * inf / w[row])
        rg, rr = int(tg[q, 3]), int(tr[q, 3])
        out["dual_rows(gpu, oracle)"] = [rg, rr]
        out["row_scores_on_gpu(gpu row, oracle row)"] = [score(xb_g, w_g, rg), score(xb_g, w_g, rr)]
        out["row_scores_on_oracle(gpu row, oracle row)"] = [score(xb_r, w_r, rg), score(xb_r, w_r, rr)]
        out["max_rel_diff_xB"] = float(np.max(np.abs(xb_g - xb_r) / np.maximum(1.0, np.abs(xb_r))))
        out["max_rel_diff_w"] = float(np.max(np.abs(w_g - w_r) / np.maximum(1.0, np.abs(w_r))))
        if rg == rr:
            e.calc_row_coeffs(rg)
            rho_r, rc_r = r2.probe_row_coeffs(rr)
            nb = g2.nb_vars()
            rc_g = e.download(7)[nb]
            dg, dr = g2.nb_var_obj_coeffs(), np.asarray(r2.nb_var_obj_coeffs)
            cg, cr = int(tg[q, 2]), int(tr[q, 2])
            out["entering_cols(gpu, oracle)"] = [cg, cr]
            out["row_coeffs_on_gpu"] = [float(rc_g[cg]), float(rc_g[cr])]
            out["row_coeffs_on_oracle"] = [float(rc_r[cg]), float(rc_r[cr])]
            out["d_on_gpu"] = [float(dg[cg]), float(dg[cr])]
            out["d_on_oracle"] = [float(dr[cg]), float(dr[cr])]
    else:
        dg, gg = g2.nb_var_obj_coeffs(), g2.primal_edge_sq_norms()
        dr, gr = np.asarray(r2.nb_var_obj_coeffs), np.asarray(r2.primal_edge_sq_norms)
        cg, cr = int(tg[q, 2]), int(tr[q, 2])
        out["entering_cols(gpu, oracle)"] = [cg, cr]
        out["d(gpu side; oracle side)"] = [[float(dg[cg]), float(dg[cr])], [float(dr[cg]), float(dr[cr])]]
        out["gamma(gpu side; oracle side)"] = [[float(gg[cg]), float(gg[cr])], [float(gr[cg]), float(gr[cr])]]
print(json.dumps(out))
