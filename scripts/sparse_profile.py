"""Short sparse-engine run for profiling (ncu launch list / timing): family m n col_nnz pivots."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minilp_b200 as mb
from minilp_b200 import mps, synth

family = sys.argv[1] if len(sys.argv) > 1 else "netlib_like"
m, n = int(sys.argv[2]) if len(sys.argv) > 2 else 30000, int(sys.argv[3]) if len(sys.argv) > 3 else 30000
cn = float(sys.argv[4]) if len(sys.argv) > 4 else 30.0
piv = int(sys.argv[5]) if len(sys.argv) > 5 else 300
warm = int(sys.argv[6]) if len(sys.argv) > 6 else 200
text, d = getattr(synth, family)(m, n, cn, 1)
p = mps.MpsFile.parse(text, d).problem
rp, ci, va, ops, rhs = p.to_csr()
s = mb.Solver(len(ops), len(p.obj_coeffs), csr=(rp, ci, va))
s.init(np.array(p.obj_coeffs), np.array(p.var_mins), np.array(p.var_maxs), ops, rhs)
s.set_record_trace(False)
e = s.engine
s.run(warm)
e.sync(); c0 = e.counters(); p0 = s.pivots_done; w0 = time.perf_counter(); e.event_mark(0)
done = s.run(piv)
e.event_mark(1); e.sync(); w1 = time.perf_counter()
c1 = e.counters()
k = s.pivots_done - p0
run_s, refac_s = s.timers()
print(json.dumps({"family": family, "m": m, "n": n, "nnz": len(va), "pivots": k, "ms_per_pivot_device": e.event_elapsed_ms(0, 1) / max(k, 1),
                  "ms_per_pivot_wall": (w1 - w0) * 1e3 / max(k, 1), "launches_per_pivot": (c1["kernel_launches"] - c0["kernel_launches"]) / max(k, 1),
                  "refactors": c1["refactors"] - c0["refactors"], "k_structural": c1["k_structural"], "eta_count": c1["eta_count"],
                  "refactor_wall_s_total": refac_s, "d2h_per_pivot": (c1["d2h_bytes"] - c0["d2h_bytes"]) / max(k, 1), "done": bool(done)}))
