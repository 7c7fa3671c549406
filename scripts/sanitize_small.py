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
