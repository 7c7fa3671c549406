"""The deep-run golden traces (tests/golden/deep_trace_*.npz) are the oracle's own output: re-running the oracle reproduces
their first pivots bit for bit (CPU only; the whole traces take minutes, see tests/golden/make_deep_trace.py)."""
import glob
import os

import numpy as np
import pytest

import minilp_b200 as mb
import oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deep_trace_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_the_golden_prefix(path):
    g = np.load(path)
    kind, m, n, seed = (int(g[k]) for k in ("kind", "m", "n", "seed"))
    lp = mb.synth_dense(kind, m, n, seed, threads=os.cpu_count() or 1)
    ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
    ref.continue_solve(300)
    tr = ref.trace()
    assert tr.shape[0] == 300
    assert np.array_equal(tr[:, :5].astype(np.int32), g["seq"][:300])
    assert np.array_equal(tr[99::100, 7], g["obj_every_100"][:3])
