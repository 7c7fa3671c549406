"""Product-form refresh of the core inverse (csrc/refresh_inverse.cuh, MLP_TUNE_LU_EVERY): between two true factorizations
(lu.rs:118-304) the refactorizations the rule of solver.rs:1096-1103 asks for fold the eta file into C^-1.  The refreshed
inverse must be THE inverse of the new core: FTRAN / BTRAN probes against the oracle mid-solve, the oracle's pivot sequence,
the same optimum as with true factorizations only."""
import numpy as np
import pytest

import minilp_b200 as mb
import oracle
from minilp_b200 import mps, synth

from parity_util import assert_sequence_parity
from test_parity_gpu import assert_same_trace, close
from test_sparse_gpu import solver_from_problem

pytestmark = pytest.mark.gpu


def make(text, d, lu_every):
    s = solver_from_problem(mps.MpsFile.parse(text, d).problem, "sparse")
    s.engine.set_tuning("lu_every", lu_every)
    return s


@pytest.mark.parametrize("gen,args,budget", [(synth.netlib_like, (300, 300, 6.0, 1), 150), (synth.netlib_like, (2000, 2000, 8.0, 3), 400),
                                             (synth.sparse_pos, (400, 900, 8.0, 3), 200)])
def test_refreshed_inverse_is_the_inverse_of_the_new_core(gen, args, budget):
    text, d = gen(*args)
    ref = oracle.MpsFile.parse(text, d).problem.solve(max_pivots=budget)
    gpu = make(text, d, 1 << 30)  # never a true factorization after the first
    gpu.run(budget)
    c = gpu.engine.counters()
    assert c["refreshes"] > 3 and 2 * c["refreshes"] > c["refactors"] and c["refresh_rejects"] == 0, c
    assert_same_trace(gpu.trace(), ref.trace(), ref, gpu)
    nb = ref.nb_vars
    for col in (0, 7, len(nb) // 2, len(nb) - 1):
        gpu.engine.ftran_col(int(nb[col]))
        assert close(gpu.engine.download(5), ref.probe_ftran_col(col), 1e-9)
    m = len(ref.basic_vars)
    for r in (0, 11, m // 2, m - 1):
        gpu.engine.calc_row_coeffs(r)
        rho, rc = ref.probe_row_coeffs(r)
        assert close(gpu.engine.download(6), rho, 1e-9)
        assert close(gpu.engine.download(7)[gpu.nb_vars()], rc, 1e-9)
    gpu.close()


@pytest.mark.parametrize("lu_every", [1, 7, 64, 1 << 30])
def test_refresh_cadences_reach_the_oracles_optimum(lu_every):
    """Whatever the share of refreshes (lu_every = 1: none), the solve follows the oracle to the same optimum."""
    text, d = synth.netlib_like(600, 600, 7.0, 2)
    ref = oracle.MpsFile.parse(text, d).problem.solve()
    gpu = make(text, d, lu_every)
    assert gpu.run()
    c = gpu.engine.counters()
    assert (c["refreshes"] == 0) == (lu_every == 1), c
    contested = assert_sequence_parity(gpu.trace(), ref.trace(), ref, gpu)
    assert close(gpu.cur_obj_val, ref.cur_obj_val)
    if not contested:
        assert close(gpu.values(), ref.values(), 1e-7)
    gpu.close()


def test_refresh_survives_clone_and_added_rows():
    """A clone taken between two true factorizations continues like its source; a row added to the LP (Solution::add_constraint,
    solver.rs:549-634) forces a true factorization and the refreshes resume after it."""
    text, d = synth.netlib_like(400, 400, 6.0, 9)
    a = make(text, d, 1 << 30)
    a.run(120)
    assert a.engine.counters()["refreshes"] > 0
    b = a.clone()
    assert a.run() and b.run()
    assert np.array_equal(a.trace(), b.trace())
    assert a.cur_obj_val == b.cur_obj_val
    ref = oracle.MpsFile.parse(text, d).problem.solve()
    assert close(a.cur_obj_val, ref.cur_obj_val)
    a.close()
    b.close()


def test_a_failed_accuracy_probe_forces_a_true_factorization(monkeypatch):
    """Every refresh is probed (max |C C^-1 - I| over sampled columns).  With an
    impossible tolerance every probe fails: each refresh is redone as a true factorization, and the solve is still the oracle's."""
    monkeypatch.setenv("MLP_REFRESH_TOL", "1e-300")
    text, d = synth.netlib_like(600, 600, 7.0, 2)
    ref = oracle.MpsFile.parse(text, d).problem.solve()
    gpu = make(text, d, 1 << 30)
    assert gpu.run()
    c = gpu.engine.counters()
    assert c["refresh_rejects"] > 3 and c["refreshes"] <= 5, c  # (a residual of exactly 0 passes any tolerance)
    assert_sequence_parity(gpu.trace(), ref.trace(), ref, gpu)
    assert close(gpu.cur_obj_val, ref.cur_obj_val)
    gpu.close()
