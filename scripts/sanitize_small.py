"""Small solves for compute-sanitizer (memcheck / racecheck / synccheck): dense primal + dual + mixed, sharded (2 logical
shards), sparse via MPS, incremental ops."""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minilp_b200 as mb
from minilp_b200 import mps, synth

for kind in (0, 3, 1):
    lp = mb.synth_dense(kind, 48, 64, 2)
    s = mb.Solver.from_dense(lp)
    assert s.run()
    print("dense", kind, s.pivots_done, s.cur_obj_val)
    s.close()
text, d = synth.netlib_like(80, 100, 4.0, 1)
sol = mps.MpsFile.parse(text, d).problem.solve(storage="sparse")
print("sparse", sol.objective())
p = mb.Problem(mb.OptimizationDirection.Minimize)
v1 = p.add_var(2.0, (0.0, float("inf"))); v2 = p.add_var(1.0, (0.0, float("inf")))
p.add_constraint([(v1, 1.0), (v2, 1.0)], 1, 4.0); p.add_constraint([(v1, 1.0), (v2, 1.0)], 2, 2.0)
sol = p.solve().fix_var(v2, 1.5).add_constraint([(v1, -1.0), (v2, 1.0)], 1, 0.0)
print("incremental", sol.objective(), sol.clone().objective())
lp = mb.synth_dense(3, 40, 64, 3)
g = mb.LocalGroup(2)
outs = [None, None]
def work(r):
    s = mb.Solver.from_dense(lp, rank=r, world=2, comm=g)
    s.run(); outs[r] = s.cur_obj_val; s.close()
th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
[t.start() for t in th]; [t.join() for t in th]
print("sharded", outs)
# round 2 paths: blocked inverse (forced at small k), periodic recomputation, a sparse LP with refactorizations and the compact
# core rows, sharded sparse, row-capacity growth
os.environ["MLP_INV_BLOCKED_MIN"] = "8"
lp = mb.synth_dense(3, 64, 80, 4)
s = mb.Solver.from_dense(lp)
s.set_recalc_period(25)
assert s.run()
print("blocked inverse + recalc", s.pivots_done, s.recalcs_done, s.cur_obj_val, s.tie_stats())
s.close()
del os.environ["MLP_INV_BLOCKED_MIN"]
text, d = synth.netlib_like(160, 200, 6.0, 3)
p = mps.MpsFile.parse(text, d).problem
rp, ci, va, ops, rhs = p.to_csr()
g2 = mb.LocalGroup(2)
res = [None, None]
def work2(r):
    s = mb.Solver(len(ops), len(p.obj_coeffs), rank=r, world=2, comm=g2, csr=(rp, ci, va))
    s.init(np.array(p.obj_coeffs), np.array(p.var_mins), np.array(p.var_maxs), ops, rhs)
    s.run(); res[r] = (s.pivots_done, s.cur_obj_val, s.engine.counters()["refactors"]); s.close()
th = [threading.Thread(target=work2, args=(r,)) for r in range(2)]
[t.start() for t in th]; [t.join() for t in th]
print("sharded sparse", res)
q = mb.Problem(mb.OptimizationDirection.Maximize)
xs = [q.add_var(1.0, (0.0, 10.0)) for _ in range(6)]
q.add_constraint([(x, 1.0) for x in xs], 1, 30.0)
sol = q.solve()
for t in range(70):
    sol = sol.add_constraint([(xs[t % 6], 1.0), (xs[(t + 1) % 6], 0.5)], 1, 14.0 - 0.01 * t)
print("row growth", sol.objective())
