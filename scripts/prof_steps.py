"""Ad-hoc timing of individual engine calls at bench scale (run under gpurun)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=50000); ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--pivots", type=int, default=60)
a0 = ap.parse_args()
a = argparse.Namespace(m=a0.m, n=a0.n, kind=0, seed=1)
s, setup = bench.build_solver(a, 0)
print("setup", setup)
e = s.engine
s.run(a0.pivots)
print("counters", e.counters())
def T(f, reps=5):
    e.sync(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); f(); e.sync(); ts.append(time.perf_counter() - t)
    return min(ts) * 1e3, np.median(ts) * 1e3
nb = s.nb_vars()
print("refactor ms (min, med):", T(lambda: e.refactor()))
s.run(10)
print("counters", e.counters())
print("select_entering", T(lambda: e.select_entering_primal()))
print("ftran_col", T(lambda: e.ftran_col(int(nb[5]))))
print("ratio_primal", T(lambda: e.ratio_primal(1, 1e30)))
print("btran_unit", T(lambda: e.btran_unit(7)))
print("price_row", T(lambda: e.price_row()))
print("select_row_dual", T(lambda: e.select_row_dual()))
t = time.perf_counter(); s.run(20); e.sync(); print("20 pivots ms/pivot", (time.perf_counter() - t) * 1e3 / 20, e.counters())
