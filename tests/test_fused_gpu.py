"""The fused FTRAN -> BTRAN chain (csrc/chain_fused.cuh, one cooperative kernel) against the separate kernels it replaces
and against the oracle.  MLP_FUSED / MLP_FUSED_MAX / MLP_LANE1_LDG are read when an engine is created."""
import os

import numpy as np
import pytest

import minilp_b200 as mb
import oracle

pytestmark = pytest.mark.gpu


def close(a, b, rel):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    return bool(np.all(np.abs(a - b) <= rel * scale))


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update({k: str(v) for k, v in self.kw.items()})

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def make(lp, **kw):
    with env(**kw):
        return mb.Solver.from_dense(lp)


def state(s):
    e = s.engine
    return dict(d=e.download(0), gam=e.download(1), xnb=e.download(2), xb=e.download(3), w=e.download(4), alpha=e.download(5),
                helper=e.download(10), basic=s.basic_vars(), nb=s.nb_vars())


@pytest.mark.parametrize("kind,m,n,seed,steps", [(0, 200, 300, 1, 60), (3, 150, 400, 4, 80), (0, 1000, 1000, 1, 120),
                                                 (3, 97, 131, 9, 50), (0, 33, 1, 1, 5), (0, 1, 40, 1, 3)])
def test_fused_chain_equals_separate_kernels_per_pivot(kind, m, n, seed, steps):
    """Same LP on two engines, one pivot at a time: identical basis sequence, all state vectors within 1e-9 — including
    alpha_q (the FTRAN result) and the helper N^T v (which is built from the chain's v list)."""
    lp = mb.synth_dense(kind, m, n, seed)
    f, u = make(lp, MLP_FUSED=1), make(lp, MLP_FUSED=0)
    for it in range(steps):
        df, du = f.run(1), u.run(1)
        assert df == du, f"termination differs at pivot {it}"
        sf, su = state(f), state(u)
        assert np.array_equal(sf["basic"], su["basic"]) and np.array_equal(sf["nb"], su["nb"]), f"basis differs at pivot {it}"
        for key in ("d", "gam", "xnb", "xb", "w", "alpha"):
            assert close(sf[key], su[key], 1e-9), f"{key} differs at pivot {it}"
        # N^T v sums ~m products of size |v| |A|: entries that cancel to ~0 carry the absolute error of the big ones,
        # so the helper is compared relative to its largest entry
        hs = max(1.0, float(np.abs(su["helper"]).max()))
        assert np.all(np.abs(sf["helper"] - su["helper"]) <= 1e-9 * hs), f"helper differs at pivot {it}"
        if df:
            break
    cf, cu = f.engine.counters(), u.engine.counters()
    assert cf["kernel_launches"] < cu["kernel_launches"]  # the fused engine really took the fused path
    f.close()
    u.close()


@pytest.mark.parametrize("kind", [0, 3])
def test_hand_over_between_fused_and_separate_paths(kind):
    """MLP_FUSED_MAX=6: the chain is fused while k, K <= 6 and falls back to the separate kernels beyond (in production the
    limit is 512) — the solve must not notice the hand-over, in either direction (K drops to 0 at every refactorization)."""
    lp = mb.synth_dense(kind, 120, 160, 2)
    s = make(lp, MLP_FUSED=1, MLP_FUSED_MAX=6)
    ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
    assert s.run() and ref.continue_solve()
    tg, tr = s.trace(), ref.trace()
    assert tg.shape == tr.shape and np.array_equal(tg[:, :5], tr[:, :5])
    assert close(s.cur_obj_val, ref.cur_obj_val, 1e-8)
    assert s.engine.counters()["k_structural"] > 6
    s.close()


def test_probe_after_fused_pivots_matches_oracle():
    """mlp_ftran_col of an arbitrary variable goes through the fused chain too; its alpha against the oracle's probe."""
    lp = mb.synth_dense(0, 400, 500, 3)
    s = make(lp, MLP_FUSED=1)
    ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
    s.run(45)
    ref.continue_solve(45)
    assert np.array_equal(s.trace()[:, :5], ref.trace()[:, :5])
    nbv = s.nb_vars()
    for col in (0, 11, 250, 499):
        s.engine.ftran_col(int(nbv[col]))
        assert close(s.engine.download(5), ref.probe_ftran_col(col), 1e-9)
    assert s.run() and ref.continue_solve()
    assert np.array_equal(s.trace()[:, :5], ref.trace()[:, :5])
    s.close()


def test_lane1_price_kernel_choice_is_bit_identical():
    """The tableau-row price-out runs as the LDG kernel beside lane 0's bulk-copy kernel; both forms give the same bits."""
    lp = mb.synth_dense(0, 300, 700, 5)
    a, b = make(lp, MLP_LANE1_LDG=1), make(lp, MLP_LANE1_LDG=0)
    assert a.run(80) == b.run(80)
    assert np.array_equal(a.trace(), b.trace())
    for which in (0, 1, 3, 4, 7):
        assert np.array_equal(a.engine.download(which), b.engine.download(which))
    a.close()
    b.close()


def test_price_tilings_are_bit_identical():
    """Tile width and tail split of the bulk-copy price-out only change which CTA sums which columns: every column adds
    its rows in list order within the same support chunks, so the product has the same bits under every tiling."""
    lp = mb.synth_dense(0, 700, 3000, 4)
    s = make(lp)
    s.run(25)
    e = s.engine
    ref = None
    for tile, split in [(512, 1), (1024, 1), (1024, 4), (2048, 2), (704, 1), (1088, 1), (128, 1), (256, 2), (1536, 4)]:
        e.set_tuning("price_tile", tile)
        e.set_tuning("price_split", split)
        assert e.get_tuning("price_tile") == tile and e.get_tuning("price_split") == split
        e.bench_price_dense(1)
        h = e.download(10)
        assert np.any(h != 0.0)
        if ref is None:
            ref = h
            assert np.allclose(h[:3000], 0.5 * lp.a.sum(axis=0) * (h[:3000] != 0), rtol=1e-12)
        assert np.array_equal(h, ref), (tile, split)
    e.set_tuning("price_tile", 0)  # automatic again
    assert s.run()
    s.close()
