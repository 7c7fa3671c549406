"""Solution::add_constraint / fix_var / unfix_var / add_gomory_cut (lib.rs:368-423) on the device
engine — the reference's own tests (lib.rs:544-645; `clone()` replaced by re-solving) and differential runs against the
oracle on mid-size LPs — for both storages of the matrix (dense rows in HBM; CSR + CSC)."""
import numpy as np
import pytest

import minilp_b200 as mb
import oracle

from test_parity_gpu import close

pytestmark = pytest.mark.gpu
INF = float("inf")
Le, Ge, Eq = mb.ComparisonOp.Le, mb.ComparisonOp.Ge, mb.ComparisonOp.Eq
STORAGES = pytest.mark.parametrize("storage", ["dense", "sparse"])


def fix_unfix_problem():
    p = mb.Problem(mb.OptimizationDirection.Maximize)
    v1 = p.add_var(1.0, (0.0, 3.0))
    v2 = p.add_var(2.0, (0.0, 3.0))
    p.add_constraint([(v1, 1.0), (v2, 1.0)], Le, 4.0)
    p.add_constraint([(v1, 1.0), (v2, 1.0)], Ge, 1.0)
    return p, v1, v2


@STORAGES
def test_lib_fix_unfix_var(storage):
    """lib.rs:544-576"""
    p, v1, v2 = fix_unfix_problem()
    _solve, p.solve = p.solve, lambda: _solve(storage=storage)
    sol = p.solve().fix_var(v1, 0.5)
    assert (sol[v1], sol[v2], sol.objective()) == (0.5, 3.0, 6.5)
    sol, was = sol.unfix_var(v1)
    assert was and (sol[v1], sol[v2], sol.objective()) == (1.0, 3.0, 7.0)
    sol, was = sol.unfix_var(v1)
    assert not was
    sol = p.solve().fix_var(v2, 2.5)
    assert (sol[v1], sol[v2], sol.objective()) == (1.5, 2.5, 6.5)
    sol, was = sol.unfix_var(v2)
    assert was and (sol[v1], sol[v2], sol.objective()) == (1.0, 3.0, 7.0)
    with pytest.raises(mb.Infeasible):
        p.solve().fix_var(v1, 3.5)  # outside the bounds, solver.rs:379-381


@STORAGES
def test_clone_continues_identically(storage):
    """Solution: Clone (lib.rs:313): the reference's tests branch off one solved problem with clone()."""
    p, v1, v2 = fix_unfix_problem()
    orig = p.solve(storage=storage)
    a = orig.clone().fix_var(v1, 0.5)
    assert (a[v1], a[v2], a.objective()) == (0.5, 3.0, 6.5)
    b = orig.clone().fix_var(v2, 2.5)
    assert (b[v1], b[v2], b.objective()) == (1.5, 2.5, 6.5)
    assert (orig[v1], orig[v2], orig.objective()) == (1.0, 3.0, 7.0)  # the source is untouched
    # mid-solve clone of a larger LP (factors + eta file copied): both continue to the same end, pivot for pivot
    lp = mb.synth_dense(3, 80, 120, 5)
    if storage == "sparse":
        rp = np.arange(0, 80 * 120 + 1, 120, dtype=np.int64)
        ci = np.tile(np.arange(120, dtype=np.int32), 80)
        s = mb.Solver.from_csr(lp.direction, 80, 120, rp, ci, lp.a.ravel(), lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
    else:
        s = mb.Solver.from_dense(lp)
    s.run(37)
    c = s.clone()
    assert s.run() and c.run()
    assert np.array_equal(s.trace(), c.trace())
    assert s.cur_obj_val == c.cur_obj_val and np.array_equal(s.values(), c.values())
    s.close()
    c.close()


def add_constraint_problem():
    p = mb.Problem(mb.OptimizationDirection.Minimize)
    v1 = p.add_var(2.0, (0.0, INF))
    v2 = p.add_var(1.0, (0.0, INF))
    p.add_constraint([(v1, 1.0), (v2, 1.0)], Le, 4.0)
    p.add_constraint([(v1, 1.0), (v2, 1.0)], Ge, 2.0)
    return p, v1, v2


@STORAGES
def test_lib_add_constraint(storage):
    """lib.rs:579-621"""
    p, v1, v2 = add_constraint_problem()
    _solve, p.solve = p.solve, lambda: _solve(storage=storage)
    sol = p.solve().add_constraint([(v1, -1.0), (v2, 1.0)], Le, 0.0)
    assert (sol[v1], sol[v2], sol.objective()) == (1.0, 1.0, 3.0)
    sol = p.solve().fix_var(v2, 1.5).add_constraint([(v1, -1.0), (v2, 1.0)], Le, 0.0)
    assert (sol[v1], sol[v2], sol.objective()) == (1.5, 1.5, 4.5)
    sol = p.solve().add_constraint([(v1, -1.0), (v2, 1.0)], Ge, 3.0)
    assert (sol[v1], sol[v2], sol.objective()) == (0.0, 3.0, 3.0)
    # empty expressions (lib.rs:485-526, the Solution half)
    sol = p.solve().add_constraint([], Eq, 0.0).add_constraint([], Ge, -1.0).add_constraint([], Le, 1.0)
    assert sol.objective() == 2.0
    for op, b in ((Eq, 12.0), (Ge, 34.0), (Le, -56.0)):
        with pytest.raises(mb.Infeasible):
            p.solve().add_constraint([], op, b)
    with pytest.raises(mb.Infeasible):
        p.solve().add_constraint([(v1, 1.0), (v2, 1.0)], Ge, 5.0)  # contradicts x + y <= 4


@STORAGES
def test_lib_gomory_cut(storage):
    """lib.rs:624-645"""
    p = mb.Problem(mb.OptimizationDirection.Minimize)
    v1 = p.add_var(0.0, (0.0, INF))
    v2 = p.add_var(-1.0, (0.0, INF))
    p.add_constraint([(v1, 3.0), (v2, 2.0)], Le, 6.0)
    p.add_constraint([(v1, -3.0), (v2, 2.0)], Le, 0.0)
    sol = p.solve(storage=storage)
    assert (sol[v1], sol[v2], sol.objective()) == (1.0, 1.5, -1.5)
    sol = sol.add_gomory_cut(v2)
    assert abs(sol[v1] - 2.0 / 3.0) < 1e-8 and abs(sol[v2] - 1.0) < 1e-12 and abs(sol.objective() + 1.0) < 1e-12
    sol = sol.add_gomory_cut(v1)
    assert abs(sol[v1] - 1.0) < 1e-8 and abs(sol[v2] - 1.0) < 1e-12 and abs(sol.objective() + 1.0) < 1e-12


def both(kind, m, n, seed, storage="dense"):
    lp = mb.synth_dense(kind, m, n, seed)
    p = mb.Problem(lp.direction)
    q = oracle.Problem(lp.direction)
    for j in range(n):
        p.add_var(lp.obj[j], (lp.mins[j], lp.maxs[j]))
        q.add_var(lp.obj[j], (lp.mins[j], lp.maxs[j]))
    for i in range(m):
        e = [(j, float(lp.a[i, j])) for j in range(n)]
        p.add_constraint(e, int(lp.ops[i]), lp.rhs[i])
        q.add_constraint(e, int(lp.ops[i]), lp.rhs[i])
    return lp, p.solve(storage=storage), q.solve()  # the oracle keeps the reference's own tie rule


def same(g, r):
    assert close(g.objective(), r.objective()), (g.objective(), r.objective())
    assert close(g.solver.values(), r.values())
    tg, tr = g.solver.trace(), r.trace()
    assert r.near_tie_pivots == 0, "contested ratio-test winner: the sequence would depend on the tie rule"
    assert tg.shape[0] == tr.shape[0] and np.array_equal(tg[:, [1, 3, 4]], tr[:, [1, 3, 4]]), "basis sequence differs"


@STORAGES
@pytest.mark.parametrize("kind,m,n,seed", [(0, 40, 60, 1), (3, 50, 50, 2), (1, 30, 45, 3)])
def test_incremental_ops_match_oracle(kind, m, n, seed, storage):
    lp, g, r = both(kind, m, n, seed, storage)
    same(g, r)
    rng = np.random.default_rng(seed)
    x = r.values()
    # fix a few variables inside their bounds (basic and non-basic ones), then release them again
    order = rng.permutation(n)[:6]
    for v in order:
        lo, hi = lp.mins[v], lp.maxs[v]
        val = float(np.clip(x[v] + 0.25 * rng.standard_normal(), lo if np.isfinite(lo) else -5.0, hi if np.isfinite(hi) else 5.0))
        try:
            r.fix_var(int(v), val)
            ok = True
        except oracle.Infeasible:
            ok = False
        if not ok:
            with pytest.raises(mb.Infeasible):
                g.fix_var(int(v), val)
            return
        g.fix_var(int(v), val)
        same(g, r)
        assert g[int(v)] == val
    for v in order[:3]:
        g, wg = g.unfix_var(int(v))
        _, wr = r.unfix_var(int(v))
        assert wg == wr
        same(g, r)
    # cut off the current optimum with random rows through it
    for t in range(4):
        x = r.values()
        idx = np.sort(rng.choice(n, size=min(n, 7), replace=False))
        co = rng.standard_normal(idx.size)
        act = float(co @ x[idx])
        e = [(int(j), float(c)) for j, c in zip(idx, co)]
        op, b = (Le, act - 0.05) if t % 2 == 0 else (Ge, act + 0.05)
        try:
            r.add_constraint(e, op, b)
            ok = True
        except oracle.Infeasible:
            ok = False
        if not ok:
            with pytest.raises(mb.Infeasible):
                g.add_constraint(e, op, b)
            return
        g.add_constraint(e, op, b)
        same(g, r)


@STORAGES
def test_gomory_cuts_match_oracle_objective(storage):
    """Gomory cuts carry slack coefficients; the engine eliminates them, which changes the dual
    steepest-edge weights of later pivots: end states are compared, not the sequence."""
    lp, g, r = both(0, 30, 40, 4, storage)
    for _ in range(3):
        x = r.values()
        frac = np.abs(x - np.round(x))
        basic = [v for v in np.argsort(-frac) if frac[v] > 1e-6]
        if not basic:
            break
        v = int(basic[0])
        r.add_gomory_cut(v)
        g.add_gomory_cut(v)
        assert close(g.objective(), r.objective(), 1e-7), (g.objective(), r.objective())


@STORAGES
def test_many_added_rows_grow_the_row_capacity(storage):
    """Solution::add_constraint has no limit on the number of rows (lib.rs:368-382); the engine's row arrays are allocated
    for m + max(64, m/8) rows and must grow beyond that (a cutting-plane loop adds hundreds)."""
    lp, g, r = both(0, 20, 30, 5, storage)
    same(g, r)
    rng = np.random.default_rng(11)
    n = 30
    added = 0
    for t in range(100):
        # a random row that cuts off the current optimum but keeps x = 0 feasible (dense_pos: A x <= b with b > 0, x >= 0),
        # so the problem stays feasible however many rows are added
        x = r.values()
        idx = np.sort(rng.choice(n, size=5, replace=False))
        co = np.round(rng.standard_normal(idx.size), 3)
        act = float(co @ x[idx])
        if abs(act) < 0.05:
            continue
        e = [(int(j), float(c)) for j, c in zip(idx, co)]
        op, b = (Le, 0.9 * act) if act > 0 else (Ge, 0.9 * act)
        r.add_constraint(e, op, b)
        g.add_constraint(e, op, b)
        added += 1
        assert close(g.objective(), r.objective()), (t, g.objective(), r.objective())
    assert added > 70, added  # past the initial reserve of 64 rows
    assert close(g.solver.values(), r.values())
    c = g.clone()  # a clone of a grown engine keeps working
    c.add_constraint([(0, 1.0)], Le, float(r.values()[0]) + 1.0)
    assert close(c.objective(), g.objective())


def test_unconstrained_problem_takes_incremental_constraints():
    """lib.rs:368: a Solution of a problem without constraints accepts add_constraint / fix_var / unfix_var."""
    p = mb.Problem(mb.OptimizationDirection.Minimize)
    x = p.add_var(1.0, (0.0, 10.0))
    y = p.add_var(2.0, (1.0, 10.0))
    sol = p.solve()
    assert (sol[x], sol[y], sol.objective()) == (0.0, 1.0, 2.0)
    sol = sol.fix_var(x, 3.0)
    assert sol.objective() == 5.0
    sol = sol.add_constraint([(x, 1.0), (y, 1.0)], Ge, 6.0)
    assert close(sol.objective(), 3.0 + 2.0 * 3.0) and close(sol[y], 3.0)
    sol, was = sol.unfix_var(x)
    assert was and close(sol.objective(), 5.0 + 2.0 * 1.0) and close(sol[x], 5.0)
