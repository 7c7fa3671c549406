"""periodic from-scratch recomputation of x_B and d (recalc_basic_var_vals solver.rs:1177-1197 — dead code
in the reference —, recalc_obj_coeffs 1199-1231; asked for by the TODO at 1024-1025).  Off by default: the default path is
the reference's and stays bit-unchanged; behind the flag the primal residual A x + s - rhs must not grow."""
import numpy as np
import pytest

import minilp_b200 as mb
import oracle
from parity_util import close

pytestmark = pytest.mark.gpu


def residual(s, lp):
    """max_i |rhs_i - a_i x - s_i| over the current basic solution."""
    e = s.engine
    n, m = lp.a.shape[1], lp.a.shape[0]
    fl, pos = e.var_state()
    xnb, xb = e.download(2), e.download(3)
    val = np.where((fl & 4) != 0, xb[np.clip(pos, 0, m - 1)], xnb)
    r = lp.rhs - lp.a @ val[:n] - val[n:]
    return float(np.max(np.abs(r))), float(max(1.0, np.max(np.abs(lp.rhs))))


def test_default_path_is_unchanged_and_flag_changes_nothing_until_set():
    lp = mb.synth_dense(3, 150, 200, 2)
    a, b = mb.Solver.from_dense(lp), mb.Solver.from_dense(lp)
    b.set_recalc_period(0)
    assert a.run() and b.run()
    assert np.array_equal(a.trace(), b.trace()) and a.recalcs_done == b.recalcs_done == 0
    for which in (0, 2, 3, 4):
        assert np.array_equal(a.engine.download(which), b.engine.download(which))
    a.close()
    b.close()


def test_recalc_basic_vals_matches_the_updated_values():
    lp = mb.synth_dense(3, 300, 400, 1)
    s = mb.Solver.from_dense(lp)
    s.run(150)
    before = s.engine.download(3)
    r0, scale = residual(s, lp)
    s.engine.recalc_basic_vals()
    after = s.engine.download(3)
    r1, _ = residual(s, lp)
    assert close(after, before, 1e-9)
    assert r1 <= max(r0, 1e-12 * scale) and r1 <= 1e-10 * scale
    assert s.run()  # the solve continues from the recomputed values
    ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
    assert ref.continue_solve()
    assert close(s.cur_obj_val, ref.cur_obj_val)
    s.close()


@pytest.mark.parametrize("kind,m,n,seed,period", [(3, 600, 600, 2, 100), (0, 400, 700, 1, 64), (1, 300, 500, 3, 50)])
def test_periodic_recalc_keeps_the_residual_down_and_reaches_the_same_optimum(kind, m, n, seed, period):
    lp = mb.synth_dense(kind, m, n, seed)
    plain, rec = mb.Solver.from_dense(lp), mb.Solver.from_dense(lp)
    rec.set_recalc_period(period)
    assert plain.run() and rec.run()
    assert rec.recalcs_done >= 1
    rp, scale = residual(plain, lp)
    rr, _ = residual(rec, lp)
    assert rr <= max(rp, 1e-12 * scale), (rr, rp)
    assert rr <= 1e-9 * scale
    assert close(rec.cur_obj_val, plain.cur_obj_val)
    assert close(rec.values(), plain.values(), 1e-7)
    plain.close()
    rec.close()


def test_recalc_on_sparse_storage():
    from minilp_b200 import mps, synth
    from test_sparse_gpu import solver_from_problem
    text, d = synth.netlib_like(400, 400, 6.0, 2)
    p = mps.MpsFile.parse(text, d).problem
    plain, rec = solver_from_problem(p, "sparse"), solver_from_problem(p, "sparse")
    rec.set_recalc_period(40)
    assert plain.run() and rec.run()
    assert rec.recalcs_done >= 1
    assert close(rec.cur_obj_val, plain.cur_obj_val)
    plain.close()
    rec.close()
