"""Tile-width sweep of the bulk-copy price-out kernel at the per-GPU shapes of the 50k x 50k bench (run under gpurun).
For every local width n_loc (= 50 000 / N GPUs) and tile width: isolated dense N^T v bandwidth (all m rows), whether the
result is bit-identical to the 512-column tiling, and the step time of a short pivot run."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=50000)
ap.add_argument("--widths", default="50000,25000,12500,6250")
ap.add_argument("--tiles", default="512,1024,2048,256")
ap.add_argument("--pivots", type=int, default=40)
a0 = ap.parse_args()
out = []
for n in [int(x) for x in a0.widths.split(",")]:
    a = argparse.Namespace(m=a0.m, n=n, kind=0, seed=1)
    s, setup = bench.build_solver(a, 0)
    e = s.engine
    s.run(8)
    ref = None
    for tile in [int(x) for x in a0.tiles.split(",")]:
        e.set_tuning("price_tile", tile)
        ms, by = e.bench_price_dense(10)
        h = e.download(10)
        if ref is None:
            ref = h
        p0 = s.pivots_done
        e.sync(); e.event_mark(0)
        t0 = time.perf_counter()
        s.run(a0.pivots)
        e.event_mark(1); e.sync()
        rec = {"n_loc": n, "tile": tile, "isolated_ms": ms, "isolated_GBps": by / (ms * 1e-3) / 1e9,
               "same_bits_as_first": bool(np.array_equal(h, ref)), "pivots": s.pivots_done - p0,
               "ms_per_pivot": e.event_elapsed_ms(0, 1) / max(1, s.pivots_done - p0), "k": e.counters()["k_structural"]}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    s.close()
