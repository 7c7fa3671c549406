"""Pins the oracle (oracle/minilp_oracle.hpp) against EVERY known-answer test the reference holds for the
simplex path.  Each test names the reference test it restates (file:line under /root/reference/src).
The reference is pure Rust and cannot be executed in this image; these vectors are the asserted values
of its own #[test]s and doctests."""
import os

import numpy as np
import pytest

import oracle
from oracle import (ComparisonOp, Infeasible, MpsFile, OptimizationDirection, Panic, Problem, SingularMatrix,
                    Unbounded, INF)


# ----------------------------------------------------------------------------- lib.rs
def test_readme_doctest_config1():
    """lib.rs:28-44 / README.md:28-45 — BASELINE config 1."""
    p = Problem(OptimizationDirection.Maximize)
    x = p.add_var(1.0, (0.0, INF))
    y = p.add_var(2.0, (0.0, 3.0))
    p.add_constraint([(x, 1.0), (y, 1.0)], ComparisonOp.Le, 4.0)
    p.add_constraint([(x, 2.0), (y, 1.0)], ComparisonOp.Ge, 2.0)
    s = p.solve()
    assert s.objective() == 7.0
    assert s[x] == 1.0
    assert s[y] == 3.0
    # hand-derived trace: one primal pivot, x enters at position 0, slack of row 0 leaves
    t = s.trace()
    assert t.shape[0] == 1
    assert (t[0, 0], t[0, 1], t[0, 3], t[0, 4]) == (1, 0, 0, 2)
    assert t[0, 5] == 1.0 and t[0, 6] == 1.0 and t[0, 7] == -7.0


def test_lib_optimize():
    """lib.rs:471-482 (note the unsorted constraint at 476)."""
    p = Problem(OptimizationDirection.Maximize)
    v1 = p.add_var(3.0, (12.0, INF))
    v2 = p.add_var(4.0, (5.0, INF))
    p.add_constraint([(v1, 1.0), (v2, 1.0)], ComparisonOp.Le, 20.0)
    p.add_constraint([(v2, -4.0), (v1, 1.0)], ComparisonOp.Ge, -20.0)
    s = p.solve()
    assert s[v1] == 12.0
    assert s[v2] == 8.0
    assert s.objective() == 68.0


def test_lib_empty_expr_constraints():
    """lib.rs:485-526."""
    trivial = [([], ComparisonOp.Eq, 0.0), ([], ComparisonOp.Ge, -1.0), ([], ComparisonOp.Le, 1.0)]
    infeasible = [([], ComparisonOp.Eq, 12.0), ([], ComparisonOp.Ge, 34.0), ([], ComparisonOp.Le, -56.0)]

    def base(extra=(), second_var=False):
        p = Problem(OptimizationDirection.Minimize)
        p.add_var(1.0, (0.0, INF))
        for e, op, b in trivial:
            p.add_constraint(e, op, b)
        for e, op, b in extra:
            p.add_constraint(e, op, b)
        if second_var:
            p.add_var(-1.0, (0.0, INF))
        return p

    assert base().solve().objective() == 0.0
    sol = base().solve()
    for e, op, b in trivial:
        sol = sol.add_constraint(e, op, b)
    assert sol.objective() == 0.0
    for c in infeasible:
        with pytest.raises(Infeasible):
            base(extra=[c]).solve()
    for e, op, b in infeasible:
        with pytest.raises(Infeasible):
            base().solve().add_constraint(e, op, b)
    with pytest.raises(Unbounded):
        base(second_var=True).solve()


def test_lib_free_variables():
    """lib.rs:529-541."""
    p = Problem(OptimizationDirection.Maximize)
    v1 = p.add_var(1.0, (0.0, INF))
    v2 = p.add_var(2.0, (-INF, INF))
    p.add_constraint([(v1, 1.0), (v2, 1.0)], ComparisonOp.Le, 4.0)
    p.add_constraint([(v1, 1.0), (v2, 1.0)], ComparisonOp.Ge, 2.0)
    p.add_constraint([(v1, 1.0), (v2, -1.0)], ComparisonOp.Ge, 0.0)
    s = p.solve()
    assert s[v1] == 2.0
    assert s[v2] == 2.0
    assert s.objective() == 6.0


def test_lib_fix_unfix_var():
    """lib.rs:544-576."""
    p = Problem(OptimizationDirection.Maximize)
    v1 = p.add_var(1.0, (0.0, 3.0))
    v2 = p.add_var(2.0, (0.0, 3.0))
    p.add_constraint([(v1, 1.0), (v2, 1.0)], ComparisonOp.Le, 4.0)
    p.add_constraint([(v1, 1.0), (v2, 1.0)], ComparisonOp.Ge, 1.0)
    orig = p.solve()

    s = orig.clone().fix_var(v1, 0.5)
    assert (s[v1], s[v2], s.objective()) == (0.5, 3.0, 6.5)
    s, was = s.unfix_var(v1)
    assert was
    assert (s[v1], s[v2], s.objective()) == (1.0, 3.0, 7.0)

    s = orig.clone().fix_var(v2, 2.5)
    assert (s[v1], s[v2], s.objective()) == (1.5, 2.5, 6.5)
    s, was = s.unfix_var(v2)
    assert (s[v1], s[v2], s.objective()) == (1.0, 3.0, 7.0)


def test_lib_add_constraint():
    """lib.rs:579-621."""
    p = Problem(OptimizationDirection.Minimize)
    v1 = p.add_var(2.0, (0.0, INF))
    v2 = p.add_var(1.0, (0.0, INF))
    p.add_constraint([(v1, 1.0), (v2, 1.0)], ComparisonOp.Le, 4.0)
    p.add_constraint([(v1, 1.0), (v2, 1.0)], ComparisonOp.Ge, 2.0)
    orig = p.solve()

    s = orig.clone().add_constraint([(v1, -1.0), (v2, 1.0)], ComparisonOp.Le, 0.0)
    assert (s[v1], s[v2], s.objective()) == (1.0, 1.0, 3.0)
    s = orig.clone().fix_var(v2, 1.5).add_constraint([(v1, -1.0), (v2, 1.0)], ComparisonOp.Le, 0.0)
    assert (s[v1], s[v2], s.objective()) == (1.5, 1.5, 4.5)
    s = orig.clone().add_constraint([(v1, -1.0), (v2, 1.0)], ComparisonOp.Ge, 3.0)
    assert (s[v1], s[v2], s.objective()) == (0.0, 3.0, 3.0)


def test_lib_gomory_cut():
    """lib.rs:624-645."""
    p = Problem(OptimizationDirection.Minimize)
    v1 = p.add_var(0.0, (0.0, INF))
    v2 = p.add_var(-1.0, (0.0, INF))
    p.add_constraint([(v1, 3.0), (v2, 2.0)], ComparisonOp.Le, 6.0)
    p.add_constraint([(v1, -3.0), (v2, 2.0)], ComparisonOp.Le, 0.0)
    s = p.solve()
    assert (s[v1], s[v2], s.objective()) == (1.0, 1.5, -1.5)
    s = s.add_gomory_cut(v2)
    assert abs(s[v1] - 2.0 / 3.0) < 1e-8
    assert s[v2] == 1.0
    assert s.objective() == -1.0
    s = s.add_gomory_cut(v1)
    assert abs(s[v1] - 1.0) < 1e-8
    assert s[v2] == 1.0
    assert s.objective() == -1.0


def test_lib_duplicate_variable_panics():
    """lib.rs:247-249: duplicate variable in one expression panics (sprs CsVec::new)."""
    p = Problem(OptimizationDirection.Minimize)
    x = p.add_var(1.0, (0.0, INF))
    with pytest.raises(Panic):
        p.add_constraint([(x, 1.0), (x, 2.0)], ComparisonOp.Le, 1.0)


# -------------------------------------------------------------------------- solver.rs
def test_solver_initialize():
    """solver.rs:1392-1441."""
    p = Problem(OptimizationDirection.Minimize)
    p.add_var(2.0, (-INF, 0.0))
    p.add_var(1.0, (5.0, INF))
    p.add_constraint([(0, 1.0), (1, 1.0)], ComparisonOp.Le, 6.0)
    p.add_constraint([(0, 1.0), (1, 2.0)], ComparisonOp.Le, 8.0)
    p.add_constraint([(0, 1.0), (1, 1.0)], ComparisonOp.Ge, 2.0)
    # to_sparse(&[0.0, 1.0]) drops the zero (helpers.rs:43-51)
    p.add_constraint([(1, 1.0)], ComparisonOp.Eq, 3.0)
    s = p.init_only()
    assert s.num_vars == 2
    assert not s.is_primal_feasible and not s.is_dual_feasible
    assert s.orig_obj_coeffs.tolist() == [2.0, 1.0, 0.0, 0.0, 0.0, 0.0]
    assert s.orig_var_mins.tolist() == [-INF, 5.0, 0.0, 0.0, -INF, 0.0]
    assert s.orig_var_maxs.tolist() == [0.0, INF, INF, INF, 0.0, 0.0]
    assert s.orig_constraints_dense().tolist() == [
        [1.0, 1.0, 1.0, 0.0, 0.0, 0.0],
        [1.0, 2.0, 0.0, 1.0, 0.0, 0.0],
        [1.0, 1.0, 0.0, 0.0, 1.0, 0.0],
        [0.0, 1.0, 0.0, 0.0, 0.0, 1.0],
    ]
    assert s.orig_rhs.tolist() == [6.0, 8.0, 2.0, 3.0]
    assert s.basic_vars.tolist() == [2, 3, 4, 5]
    assert s.basic_var_vals.tolist() == [1.0, -2.0, -3.0, -2.0]
    assert s.dual_edge_sq_norms.tolist() == [1.0, 1.0, 1.0, 1.0]
    assert s.nb_vars.tolist() == [0, 1]
    assert s.nb_var_obj_coeffs.tolist() == [-1.0, 1.0]
    assert s.nb_var_vals.tolist() == [0.0, 5.0]
    assert s.primal_edge_sq_norms.tolist() == [4.0, 8.0]
    assert s.cur_obj_val == 0.0


def test_solver_initial_solve():
    """solver.rs:1444-1479, plus the hand-derived three-pivot trace of SURVEY.md §8c."""
    p = Problem(OptimizationDirection.Minimize)
    p.add_var(-3.0, (-INF, 20.0))
    p.add_var(-4.0, (5.0, INF))
    p.add_constraint([(0, 1.0), (1, 1.0)], ComparisonOp.Le, 20.0)
    p.add_constraint([(0, -1.0), (1, 4.0)], ComparisonOp.Le, 20.0)
    s = p.solve()
    assert s.is_primal_feasible and s.is_dual_feasible
    assert s.basic_vars.tolist() == [0, 1]
    assert s.basic_var_vals.tolist() == [12.0, 8.0]
    assert s.nb_vars.tolist() == [2, 3]
    assert s.nb_var_vals.tolist() == [0.0, 0.0]
    assert s.nb_var_obj_coeffs.tolist() == [3.2, 0.2]
    assert s.cur_obj_val == -68.0
    t = s.trace()
    # dual pivot: row 0 leaves (s0: -5 -> 0), v0 enters with alpha = 1, delta = -5; then primal: v1 enters, row 1 leaves
    assert t.shape[0] == 2
    assert (t[0, 0], t[0, 1], t[0, 3], t[0, 5], t[0, 6]) == (0, 0, 0, 1.0, -5.0)
    assert (t[1, 0], t[1, 1], t[1, 3], t[1, 5], t[1, 6]) == (1, 1, 1, 5.0, 3.0)

    q = Problem(OptimizationDirection.Minimize)
    q.add_var(1.0, (0.0, INF))
    q.add_var(1.0, (0.0, INF))
    q.add_constraint([(0, 1.0), (1, 1.0)], ComparisonOp.Ge, 10.0)
    q.add_constraint([(0, 1.0), (1, 1.0)], ComparisonOp.Le, 5.0)
    with pytest.raises(Infeasible):
        q.solve()


# ------------------------------------------------------------------------------ lu.rs
def _csc_from_triplets(rows, cols, trips):
    """sprs TriMat::to_csc: entries sorted by (col, row)."""
    trips = sorted(trips, key=lambda t: (t[1], t[0]))
    ptr = [0] * (cols + 1)
    for _, c, _ in trips:
        ptr[c + 1] += 1
    for c in range(cols):
        ptr[c + 1] += ptr[c]
    return ptr, [t[0] for t in trips], [t[2] for t in trips]


def test_lu_simple():
    """lu.rs:480-552: exact factors, permutations and all four solve flavours."""
    ptr, idx, val = _csc_from_triplets(3, 4, [(0, 1, 2.0), (0, 0, 2.0), (0, 2, 123.0), (1, 2, 456.0), (1, 3, 1.0),
                                              (2, 1, 4.0), (2, 0, 3.0), (2, 2, 789.0), (2, 3, 1.0)])
    lu = oracle.LU(3, ptr, idx, val, [1, 0, 3], 0.9)
    assert lu.dense(0).tolist() == [[0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [0.0, 0.0, 0.0]]
    assert lu.dense(1).tolist() == [[0.0, 3.0, 1.0], [0.0, 0.0, -0.5], [0.0, 0.0, 0.0]]
    assert lu.dense(2).tolist() == [4.0, 0.5, 1.0]
    assert lu.perm(0).tolist() == [2, 0, 1]
    assert lu.perm(2).tolist() == [0, 1, 2]
    assert lu.solve_dense([6.0, 3.0, 13.0]).tolist() == [1.0, 2.0, 3.0]
    assert lu.solve_dense([14.0, 11.0, 5.0], transposed=True).tolist() == [1.0, 2.0, 3.0]
    out, _ = lu.solve_sparse([1], [-1.0])
    assert out.tolist() == [1.0, -1.0, -1.0]
    out, _ = lu.solve_sparse([1, 2], [-1.0, 1.0], transposed=True)
    assert out.tolist() == [-2.0, 0.0, 1.0]


def test_lu_singular():
    """lu.rs:555-609: symbolically and numerically singular."""
    ptr, idx, val = _csc_from_triplets(3, 3, [(0, 0, 1.0), (1, 0, 1.0), (1, 1, 2.0), (1, 2, 3.0)])
    with pytest.raises(SingularMatrix):
        oracle.LU(3, ptr, idx, val, [0, 1, 2], 0.9)
    ptr, idx, val = _csc_from_triplets(3, 3, [(0, 0, 1.0), (1, 0, 1.0), (1, 1, 2.0), (1, 2, 3.0), (2, 0, 2.0),
                                              (2, 1, 2.0), (2, 2, 3.0)])
    with pytest.raises(SingularMatrix):
        oracle.LU(3, ptr, idx, val, [0, 1, 2], 0.9)


@pytest.mark.parametrize("seed", [12345, 1, 2, 3])
def test_lu_rand_property(seed):
    """lu.rs:612-704 restated as a property test (the rand 0.7 / rand_pcg streams are not reproducible here):
    L*U equals the row/column-permuted matrix and all four solves have residual < 1e-5."""
    rng = np.random.default_rng(seed)
    size = 10
    while True:
        mask = rng.integers(0, 2, (size, size)) == 0
        a = np.where(mask, rng.random((size, size)), 0.0)
        if abs(np.linalg.det(a)) > 1e-4:
            break
    trips = [(r, c, a[r, c]) for r in range(size) for c in range(size) if mask[r, c]]
    ptr, idx, val = _csc_from_triplets(size, size, trips)
    lu = oracle.LU(size, ptr, idx, val, list(range(size)), 0.1)
    lmat = lu.dense(0) + np.eye(size)
    umat = lu.dense(1) + np.diag(lu.dense(2))
    o2n_row, n2o_col = lu.perm(1), lu.perm(2)
    prod = lmat @ umat
    for new_c in range(size):
        col = np.zeros(size)
        col[o2n_row] = a[:, n2o_col[new_c]]
        assert np.abs(prod[:, new_c] - col).sum() < 1e-5
    rhs = rng.random(size)
    assert np.linalg.norm(rhs - a @ lu.solve_dense(rhs)) < 1e-5
    assert np.linalg.norm(rhs - a.T @ lu.solve_dense(rhs, transposed=True)) < 1e-5
    sp_idx = [i for i in range(size) if rng.integers(0, 3) == 0] or [3]
    sp_val = rng.random(len(sp_idx))
    sp = np.zeros(size)
    sp[sp_idx] = sp_val
    out, order = lu.solve_sparse(sp_idx, sp_val)
    assert np.abs(sp - a @ out).sum() < 1e-5
    assert len(set(order.tolist())) == len(order)
    out, _ = lu.solve_sparse(sp_idx, sp_val, transposed=True)
    assert np.abs(sp - a.T @ out).sum() < 1e-5
    # transposed factors are the transposes (lu.rs:108-115)
    assert np.array_equal(lu.dense(10), lu.dense(1).T)
    assert np.array_equal(lu.dense(11), lu.dense(0).T)


# -------------------------------------------------------------------------- sparse.rs
def test_sparse_mat_transpose():
    """sparse.rs:345-359."""
    t_ptr, t_idx, t_dat = oracle.sparsemat_transpose(2, [0, 2, 3, 4], [0, 1, 1, 0], [1.1, 2.2, 3.3, 4.4])
    assert t_ptr.tolist() == [0, 2, 4]
    assert t_idx.tolist() == [2, 0, 1, 0]
    assert t_dat.tolist() == [4.4, 1.1, 3.3, 2.2]


# ----------------------------------------------------------------------------- mps.rs
MPS_TEST_FILE = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "testprob.mps")).read()


def test_mps_parse_and_solve():
    """mps.rs:437-476."""
    f = MpsFile.parse(MPS_TEST_FILE, OptimizationDirection.Minimize)
    assert f.problem_name == "TESTPROB"
    assert len(f.variables) == 3
    s = f.problem.solve()
    assert s[f.variables["XONE"]] == 4.0
    assert s[f.variables["YTWO"]] == -1.0
    assert s[f.variables["ZTHREE"]] == 6.0
    assert s.objective() == 54.0


def test_mps_ranges_and_negative_up():
    """mps.rs:294-322: negative UP without LO => (-inf, max]; RANGES expand to two rows."""
    text = """NAME R
ROWS
 N  COST
 G  R1
 E  R2
COLUMNS
    X  COST 1 R1 1
    X  R2 1
    Y  COST 1 R2 1
RHS
    RHS R1 1 R2 2
RANGES
    RNG R1 3 R2 -1
BOUNDS
 UP BND Y -1
ENDATA
"""
    f = MpsFile.parse(text, OptimizationDirection.Minimize)
    obj, mins, maxs, row_ptr, col_idx, vals, ops, rhs = f.problem.export()
    assert mins.tolist() == [0.0, -INF] and maxs.tolist() == [INF, -1.0]
    # R1: G, rhs 1, range 3 -> [1, 4];  R2: E, rhs 2, range -1 -> [1, 2]
    assert ops.tolist() == [ComparisonOp.Ge, ComparisonOp.Le, ComparisonOp.Ge, ComparisonOp.Le]
    assert rhs.tolist() == [1.0, 4.0, 1.0, 2.0]


# ------------------------------------------------------- oracle-internal consistency
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_dense_storage_matches_faithful(kind):
    """The memory-lean DenseMatrix storage must be bit-identical to the faithful CSR+CSC storage."""
    m, n = 24, 30
    d, a, obj, mins, maxs, ops, rhs = oracle.synth_dense(kind, m, n, seed=7)
    p = Problem(d)
    for j in range(n):
        p.add_var(obj[j], (mins[j], maxs[j]))
    for i in range(m):
        p.add_constraint([(j, a[i, j]) for j in range(n)], int(ops[i]), rhs[i])
    s1 = p.solve()
    s2 = oracle.DenseSolver(d, a, obj, mins, maxs, ops, rhs)
    assert s2.continue_solve()
    assert np.array_equal(s1.trace(), s2.trace())
    assert s1.objective() == s2.objective()
    assert np.array_equal(s1.values(), s2.values())
    assert s1.pivots_done > 3


def test_budgeted_solve_equals_unbudgeted():
    d, a, obj, mins, maxs, ops, rhs = oracle.synth_dense(3, 20, 25, seed=5)
    s1 = oracle.DenseSolver(d, a, obj, mins, maxs, ops, rhs)
    s1.continue_solve()
    s2 = oracle.DenseSolver(d, a, obj, mins, maxs, ops, rhs)
    while not s2.continue_solve(3):
        pass
    assert np.array_equal(s1.trace(), s2.trace())
    assert s1.objective() == s2.objective()
