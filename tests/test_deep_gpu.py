"""Deep runs against golden traces of the oracle (tests/golden/deep_trace_*.npz, made by tests/golden/make_deep_trace.py):
thousands of pivots of a mid-size dense LP, far enough for the basis to hold more than 512 structural columns and for
the eta file to grow past 512 columns — the regime beyond the fused chain's limits, where the separate kernels with
their column-group splits run.  Index work (entering variable, position, leaving row, leaving variable, phase) must match
pivot for pivot; the objective to 1e-8 relative."""
import glob
import os

import numpy as np
import pytest

import minilp_b200 as mb

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deep_trace_*.npz")))


@pytest.mark.parametrize("fused_max", [None, 128], ids=["fused<=512", "fused<=128"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_deep_run_follows_the_oracle_trace(path, fused_max, monkeypatch):
    """fused_max=None: the fused FTRAN -> BTRAN chain serves k, K up to 512; 128: the separate kernels with their
    column-group splits take over from 128 on — both hand-over points must follow the oracle."""
    if fused_max is not None:
        monkeypatch.setenv("MLP_FUSED_MAX", str(fused_max))
    g = np.load(path)
    kind, m, n, seed, budget = (int(g[k]) for k in ("kind", "m", "n", "seed", "budget"))
    assert int(g["tie_events"]) == 0  # the sequence is well-defined: the oracle met no exact tie
    lp = mb.synth_dense(kind, m, n, seed, threads=os.cpu_count() or 1)
    s = mb.Solver.from_dense(lp)
    done = s.run(budget)
    tr = s.trace()
    want = g["seq"]
    k = min(tr.shape[0], want.shape[0])
    same = np.all(tr[:k, :5].astype(np.int64) == want[:k], axis=1)
    assert same.all(), f"basis sequence leaves the oracle's at pivot {int(np.argmin(same))}: gpu {tr[int(np.argmin(same)), :5]} oracle {want[int(np.argmin(same))]}"
    assert tr.shape[0] == want.shape[0] and bool(done) == bool(g["done"])
    obj = tr[99::100, 7]
    ref = g["obj_every_100"]
    assert np.all(np.abs(obj - ref) <= 1e-8 * np.maximum(1.0, np.abs(ref)))
    assert abs(s.cur_obj_val - float(g["obj_final"])) <= 1e-8 * max(1.0, abs(float(g["obj_final"])))
    if kind == 3:
        assert s.engine.counters()["k_structural"] > 512 and int(tr[:, 8].max()) > 512
    s.close()


def test_deep_run_with_the_blocked_inverse(monkeypatch):
    """Cores of 2048+ columns get their explicit inverse by blocked substitution (dense_block.cuh) instead of the
    per-column kernels; MLP_INV_BLOCKED_MIN=48 forces that path on a golden run whose core grows past 700 columns:
    the pivot sequence must not notice."""
    monkeypatch.setenv("MLP_INV_BLOCKED_MIN", "48")
    path = [p for p in GOLDEN if "k3_1800x1800" in p][0]
    g = np.load(path)
    kind, m, n, seed = (int(g[k]) for k in ("kind", "m", "n", "seed"))
    lp = mb.synth_dense(kind, m, n, seed, threads=os.cpu_count() or 1)
    s = mb.Solver.from_dense(lp)
    budget = 2500
    s.run(budget)
    tr = s.trace()
    want = g["seq"][:budget]
    same = np.all(tr[:, :5].astype(np.int64) == want, axis=1)
    assert same.all(), f"leaves the oracle's sequence at pivot {int(np.argmin(same))}"
    assert s.engine.counters()["k_structural"] > 300
    # the two inverse algorithms agree on B^-1 a_q to rounding
    var = int(s.nb_vars()[5])
    s.engine.refactor()
    s.engine.ftran_col(var)
    a_blocked = s.engine.download(5)
    s.close()
    monkeypatch.setenv("MLP_INV_BLOCKED_MIN", "1000000")
    s2 = mb.Solver.from_dense(lp)
    s2.run(budget)
    assert np.array_equal(s2.trace()[:, :5], tr[:, :5])
    s2.engine.refactor()
    s2.engine.ftran_col(var)
    a_col = s2.engine.download(5)
    assert np.all(np.abs(a_blocked - a_col) <= 1e-9 * np.maximum(1.0, np.abs(a_col)))
    s2.close()


def test_golden_traces_exist():
    assert len(GOLDEN) >= 2
