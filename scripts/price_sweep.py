"""Tile-width / tail-split sweep of the bulk-copy price-out kernel at the per-GPU shapes of the 50k x 50k bench (run
under gpurun).  For every local width n_loc (= 50 000 / N GPUs): isolated dense N^T v bandwidth (all m rows) for a grid
of (tile, split), all at the same engine state (so the products must be bit-identical), then the step time of a short
pivot run with the automatic choice and with the old fixed 512-column tiling."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=50000)
ap.add_argument("--widths", default="50000,25000,12500,6250")
ap.add_argument("--tiles", default="512,640,768,960,1024,1088,1280,1344,1536,2048")
ap.add_argument("--pivots", type=int, default=40)
ap.add_argument("--splits", type=int, default=1, help="also sweep a few (tile, split) combinations")
a0 = ap.parse_args()
for n in [int(x) for x in a0.widths.split(",")]:
    a = argparse.Namespace(m=a0.m, n=n, kind=0, seed=1)
    s, setup = bench.build_solver(a, 0)
    e = s.engine
    s.run(8)
    auto = (e.get_tuning("price_tile"), e.get_tuning("price_split"))
    ref = None
    grid = [(t, 1) for t in [int(x) for x in a0.tiles.split(",")]]
    if a0.splits:
        grid += [(1024, 2), (1024, 4), (2048, 2), (2048, 4), (1280, 2), (1536, 4)]
    grid.append(auto)
    for tile, split in grid:
        e.set_tuning("price_tile", tile)
        e.set_tuning("price_split", split)
        ms, by = e.bench_price_dense(10)
        h = e.download(10)
        if ref is None:
            ref = h
        print(json.dumps({"n_loc": n, "tile": tile, "split": split, "auto": (tile, split) == auto, "isolated_ms": round(ms, 5),
                          "isolated_GBps": round(by / (ms * 1e-3) / 1e9, 1), "same_bits": bool(np.array_equal(h, ref))}), flush=True)
    for tile, split in [auto, (512, 1), auto]:
        e.set_tuning("price_tile", tile)
        e.set_tuning("price_split", split)
        p0 = s.pivots_done
        e.sync(); e.event_mark(0)
        s.run(a0.pivots)
        e.event_mark(1); e.sync()
        print(json.dumps({"n_loc": n, "tile": tile, "split": split, "pivots": s.pivots_done - p0,
                          "ms_per_pivot": round(e.event_elapsed_ms(0, 1) / max(1, s.pivots_done - p0), 5),
                          "k": e.counters()["k_structural"]}), flush=True)
    s.close()
