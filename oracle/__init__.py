"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of the C++ restatement of ztlpn/minilp's revised-simplex path
(oracle/minilp_oracle.hpp).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package; the
product (minilp_b200) never does.

The class names mirror the reference's public API (lib.rs:61-464) so that the
known-answer tests read like the reference's own.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

INF = float("inf")


class Infeasible(Exception):
    """Error::Infeasible, lib.rs:175"""


class Unbounded(Exception):
    """Error::Unbounded, lib.rs:177"""


class SingularMatrix(Exception):
    """sparse.rs:335-338"""


class NonFinite(Exception):
    """assert!(…is_finite()) solver.rs:1149,1172"""


class Panic(Exception):
    """a reference panic (duplicate variable lib.rs:249, non-basic gomory var solver.rs:458, …)"""


class MpsError(Exception):
    """io::ErrorKind::InvalidData, mps.rs"""


_ERRORS = {1: Infeasible, 2: Unbounded, 3: SingularMatrix, 4: NonFinite, 5: Panic, 6: MpsError, 7: RuntimeError}


class OptimizationDirection:
    Minimize = 0
    Maximize = 1


class ComparisonOp:
    Eq = 0
    Le = 1
    Ge = 2


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++ -O2 -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "minilp_oracle.hpp", "synth_lp.hpp", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, i64, f64, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
    pd, pi64, pi32 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    sig = {
        "mlo_last_error": (C.c_char_p, []),
        "mlo_problem_new": (vp, [i32]),
        "mlo_problem_free": (None, [vp]),
        "mlo_problem_add_var": (i64, [vp, f64, f64, f64]),
        "mlo_problem_add_constraint": (i32, [vp, i64, pi64, pd, i32, f64]),
        "mlo_problem_num_vars": (i64, [vp]),
        "mlo_problem_num_constraints": (i64, [vp]),
        "mlo_problem_nnz": (i64, [vp]),
        "mlo_problem_export": (None, [vp, pd, pd, pd, pi64, pi64, pd, pi32, pd]),
        "mlo_problem_solve": (i32, [vp, i32, i64, C.POINTER(vp)]),
        "mlo_problem_init_only": (i32, [vp, C.POINTER(vp)]),
        "mlo_dense_new": (i32, [i32, i64, i64, pd, pd, pd, pd, pi32, pd, i32, i32, C.POINTER(vp)]),
        "mlo_dense_new_synth": (i32, [i32, i64, i64, C.c_uint64, i32, i32, C.POINTER(vp)]),
        "mlo_free": (None, [vp]),
        "mlo_clone": (vp, [vp]),
        "mlo_continue": (i32, [vp, i64, C.POINTER(i32)]),
        "mlo_continue_timed": (i32, [vp, i64, C.POINTER(i32), pd]),
        "mlo_objective": (f64, [vp]),
        "mlo_cur_obj_val": (f64, [vp]),
        "mlo_var_value": (f64, [vp, i64]),
        "mlo_set_record_trace": (None, [vp, i32]),
        "mlo_add_constraint": (i32, [vp, i64, pi64, pd, i32, f64]),
        "mlo_fix_var": (i32, [vp, i64, f64]),
        "mlo_unfix_var": (i32, [vp, i64, C.POINTER(i32)]),
        "mlo_add_gomory_cut": (i32, [vp, i64]),
        "mlo_get_i64": (i64, [vp, i32]),
        "mlo_get_f64_array": (i32, [vp, i32, pd, i64]),
        "mlo_get_i64_array": (i32, [vp, i32, pi64, i64]),
        "mlo_get_constraints_dense": (i32, [vp, pd]),
        "mlo_get_trace": (i64, [vp, i64, i64, pd]),
        "mlo_probe_ftran_col": (i32, [vp, i64, pd]),
        "mlo_probe_row_coeffs": (i32, [vp, i64, pd, pd]),
        "mlo_parse_mps": (i32, [C.c_char_p, i64, i32, C.POINTER(vp), C.POINTER(vp)]),
        "mlo_mps_free": (None, [vp]),
        "mlo_mps_name": (C.c_char_p, [vp]),
        "mlo_mps_num_vars": (i64, [vp]),
        "mlo_mps_var_name": (C.c_char_p, [vp, i64]),
        "mlo_mps_var_index": (i64, [vp, C.c_char_p]),
        "mlo_lu_new": (i32, [i64, i64, pi64, pi64, pd, pi64, f64, C.POINTER(vp)]),
        "mlo_lu_free": (None, [vp]),
        "mlo_lu_nnz": (i64, [vp]),
        "mlo_lu_get_dense": (None, [vp, i32, pd]),
        "mlo_lu_get_perm": (None, [vp, i32, pi64]),
        "mlo_lu_solve_dense": (None, [vp, i32, pd]),
        "mlo_lu_solve_sparse": (i64, [vp, i32, i64, pi64, pd, pd, pi64]),
        "mlo_sparsemat_transpose": (None, [i64, i64, pi64, pi64, pd, pi64, pi64, pd]),
        "mlo_synth_dense": (i32, [i32, i64, i64, C.c_uint64, i32, pd, pd, pd, pd, pi32, pd]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        msg = lib().mlo_last_error().decode()
        raise _ERRORS.get(rc, RuntimeError)(msg)


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _pi32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


TRACE_FIELDS = ("phase", "entering_var", "entering_col", "leaving_row", "leaving_var", "pivot_coeff",
                "entering_diff", "obj_after", "eta_count", "lu_nnz", "nnz_col", "nnz_rho", "refactored")


class _State:
    """Read-only view of Solver fields (solver.rs:15-58) of a live handle."""

    _h = None

    def _i64(self, what):
        return int(lib().mlo_get_i64(self._h, what))

    num_vars = property(lambda s: s._i64(0))
    num_constraints = property(lambda s: s._i64(1))
    pivots_done = property(lambda s: s._i64(2))
    refactor_count = property(lambda s: s._i64(3))
    tie_events = property(lambda s: s._i64(4))
    is_primal_feasible = property(lambda s: bool(s._i64(5)))
    is_dual_feasible = property(lambda s: bool(s._i64(6)))
    enable_primal_steepest_edge = property(lambda s: bool(s._i64(7)))
    enable_dual_steepest_edge = property(lambda s: bool(s._i64(8)))
    eta_count = property(lambda s: s._i64(10))
    lu_nnz = property(lambda s: s._i64(11))
    nnz = property(lambda s: s._i64(13))
    # pass-2 winners of the ratio tests that were contested: exactly (the reference then decides by list order) / within 1e-9
    tied_pivots = property(lambda s: s._i64(14))
    near_tie_pivots = property(lambda s: s._i64(15))
    first_tied_pivot = property(lambda s: s._i64(16))
    first_near_tie_pivot = property(lambda s: s._i64(17))
    # selections (pricing, dual row) whose runner-up scored within 1e-9 of the winner: same rule on both sides, rounding decides
    sel_near_tie_pivots = property(lambda s: s._i64(18))
    first_sel_near_tie_pivot = property(lambda s: s._i64(19))

    def _farr(self, what):
        cap = self.num_vars + 2 * self.num_constraints + 8
        out = np.empty(cap, dtype=np.float64)
        n = lib().mlo_get_f64_array(self._h, what, _pd(out), cap)
        assert n >= 0
        return out[:n].copy()

    def _iarr(self, what):
        cap = self.num_vars + 2 * self.num_constraints + 8
        out = np.empty(cap, dtype=np.int64)
        n = lib().mlo_get_i64_array(self._h, what, _pi64(out), cap)
        assert n >= 0
        return out[:n].copy()

    basic_var_vals = property(lambda s: s._farr(0))
    nb_var_vals = property(lambda s: s._farr(1))
    nb_var_obj_coeffs = property(lambda s: s._farr(2))
    primal_edge_sq_norms = property(lambda s: s._farr(3))
    dual_edge_sq_norms = property(lambda s: s._farr(4))
    orig_var_mins = property(lambda s: s._farr(5))
    orig_var_maxs = property(lambda s: s._farr(6))
    orig_obj_coeffs = property(lambda s: s._farr(7))
    basic_var_mins = property(lambda s: s._farr(8))
    basic_var_maxs = property(lambda s: s._farr(9))
    orig_rhs = property(lambda s: s._farr(10))
    basic_vars = property(lambda s: s._iarr(0))
    nb_vars = property(lambda s: s._iarr(1))
    nb_var_state_bits = property(lambda s: s._iarr(2))
    cur_obj_val = property(lambda s: float(lib().mlo_cur_obj_val(s._h)))

    def orig_constraints_dense(self):
        m, t = self.num_constraints, self.num_vars + self.num_constraints
        out = np.zeros((m, t), dtype=np.float64)
        lib().mlo_get_constraints_dense(self._h, _pd(out))
        return out

    def trace(self):
        n = self._i64(9)
        out = np.empty((max(n, 1), 13), dtype=np.float64)
        got = lib().mlo_get_trace(self._h, 0, n, _pd(out))
        return out[:got]

    def values(self):
        """all structural variable values (Solution::iter, lib.rs:350)"""
        return np.array([lib().mlo_var_value(self._h, v) for v in range(self.num_vars)])

    def probe_ftran_col(self, col):
        out = np.empty(self.num_constraints, dtype=np.float64)
        _check(lib().mlo_probe_ftran_col(self._h, col, _pd(out)))
        return out

    def probe_row_coeffs(self, row):
        rho = np.empty(self.num_constraints, dtype=np.float64)
        rc = np.empty(self._i64(12), dtype=np.float64)
        _check(lib().mlo_probe_row_coeffs(self._h, row, _pd(rho), _pd(rc)))
        return rho, rc


class Problem:
    """lib.rs:192-305"""

    def __init__(self, direction):
        self._p = lib().mlo_problem_new(direction)
        self.direction = direction

    def __del__(self):
        if getattr(self, "_p", None):
            lib().mlo_problem_free(self._p)
            self._p = None

    def add_var(self, obj_coeff, bounds):
        return int(lib().mlo_problem_add_var(self._p, obj_coeff, bounds[0], bounds[1]))

    def add_constraint(self, expr, cmp_op, rhs):
        expr = list(expr)
        vars_ = np.array([v for v, _ in expr], dtype=np.int64)
        coeffs = np.array([c for _, c in expr], dtype=np.float64)
        _check(lib().mlo_problem_add_constraint(self._p, len(expr), _pi64(vars_), _pd(coeffs), cmp_op, rhs))

    @property
    def num_vars(self):
        return int(lib().mlo_problem_num_vars(self._p))

    @property
    def num_constraints(self):
        return int(lib().mlo_problem_num_constraints(self._p))

    def export(self):
        """(obj_internal, mins, maxs, row_ptr, col_idx, vals, ops, rhs): obj already sign-flipped for Maximize."""
        n, m, z = self.num_vars, self.num_constraints, int(lib().mlo_problem_nnz(self._p))
        obj, mins, maxs = (np.empty(n) for _ in range(3))
        row_ptr = np.empty(m + 1, dtype=np.int64)
        col_idx = np.empty(max(z, 1), dtype=np.int64)
        vals = np.empty(max(z, 1))
        ops = np.empty(max(m, 1), dtype=np.int32)
        rhs = np.empty(max(m, 1))
        lib().mlo_problem_export(self._p, _pd(obj), _pd(mins), _pd(maxs), _pi64(row_ptr), _pi64(col_idx), _pd(vals),
                                 _pi32(ops), _pd(rhs))
        return obj, mins, maxs, row_ptr, col_idx[:z], vals[:z], ops[:m], rhs[:m]

    def solve(self, tie_lowest_index=False, max_pivots=-1):
        h = C.c_void_p()
        _check(lib().mlo_problem_solve(self._p, int(tie_lowest_index), max_pivots, C.byref(h)))
        return Solution(h, self.direction)

    def init_only(self):
        h = C.c_void_p()
        _check(lib().mlo_problem_init_only(self._p, C.byref(h)))
        return Solution(h, self.direction)


class Solution(_State):
    """lib.rs:313-424.  The reference's methods consume `self`; here they mutate in place and return self."""

    def __init__(self, h, direction):
        self._h = h
        self.direction = direction

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mlo_free(self._h)
            self._h = None

    def clone(self):
        return Solution(C.c_void_p(lib().mlo_clone(self._h)), self.direction)

    def objective(self):
        return float(lib().mlo_objective(self._h))

    def __getitem__(self, var):
        return float(lib().mlo_var_value(self._h, var))

    def var_value(self, var):
        return self[var]

    def continue_solve(self, max_pivots=-1):
        done = C.c_int(0)
        _check(lib().mlo_continue(self._h, max_pivots, C.byref(done)))
        return bool(done.value)

    def continue_timed(self, max_pivots):
        done, sec = C.c_int(0), C.c_double(0)
        _check(lib().mlo_continue_timed(self._h, max_pivots, C.byref(done), C.byref(sec)))
        return bool(done.value), sec.value

    def set_record_trace(self, on):
        lib().mlo_set_record_trace(self._h, int(on))

    def add_constraint(self, expr, cmp_op, rhs):
        expr = list(expr)
        vars_ = np.array([v for v, _ in expr], dtype=np.int64)
        coeffs = np.array([c for _, c in expr], dtype=np.float64)
        _check(lib().mlo_add_constraint(self._h, len(expr), _pi64(vars_), _pd(coeffs), cmp_op, rhs))
        return self

    def fix_var(self, var, val):
        _check(lib().mlo_fix_var(self._h, var, val))
        return self

    def unfix_var(self, var):
        was = C.c_int(0)
        _check(lib().mlo_unfix_var(self._h, var, C.byref(was)))
        return self, bool(was.value)

    def add_gomory_cut(self, var):
        _check(lib().mlo_add_gomory_cut(self._h, var))
        return self


class DenseSolver(Solution):
    """Memory-lean dense-storage Solver (bit-identical arithmetic; see DenseMatrix in the header)."""

    def __init__(self, direction, a, obj, mins, maxs, ops, rhs, tie_lowest_index=False):
        a = np.ascontiguousarray(a, dtype=np.float64)
        m, n = a.shape
        self._keep = (a,)
        obj, mins, maxs, rhs = (np.ascontiguousarray(x, dtype=np.float64) for x in (obj, mins, maxs, rhs))
        ops = np.ascontiguousarray(ops, dtype=np.int32)
        h = C.c_void_p()
        _check(lib().mlo_dense_new(direction, m, n, _pd(a), _pd(obj), _pd(mins), _pd(maxs), _pi32(ops), _pd(rhs),
                                   int(tie_lowest_index), 0, C.byref(h)))
        super().__init__(h, direction)

    @classmethod
    def synth(cls, kind, m, n, seed, threads=1, tie_lowest_index=False):
        h = C.c_void_p()
        _check(lib().mlo_dense_new_synth(kind, m, n, seed, threads, int(tie_lowest_index), C.byref(h)))
        self = cls.__new__(cls)
        self._keep = ()
        Solution.__init__(self, h, None)
        return self


class MpsFile:
    """mps.rs:7-16, 39-329"""

    def __init__(self, text, direction):
        if isinstance(text, str):
            text = text.encode()
        p, m = C.c_void_p(), C.c_void_p()
        _check(lib().mlo_parse_mps(text, len(text), direction, C.byref(p), C.byref(m)))
        self._m = m
        self.problem = Problem.__new__(Problem)
        self.problem._p = p
        self.problem.direction = direction
        self.problem_name = lib().mlo_mps_name(m).decode()
        n = int(lib().mlo_mps_num_vars(m))
        self.variables = {lib().mlo_mps_var_name(m, i).decode(): i for i in range(n)}

    @classmethod
    def parse(cls, text, direction):
        return cls(text, direction)

    def __del__(self):
        if getattr(self, "_m", None):
            lib().mlo_mps_free(self._m)
            self._m = None


class LU:
    """lu_factorize + LUFactors probes (lu.rs:51-304) over a CSC matrix given as scipy-like arrays."""

    def __init__(self, size, col_ptr, row_idx, vals, pick, stability):
        col_ptr = np.ascontiguousarray(col_ptr, dtype=np.int64)
        row_idx = np.ascontiguousarray(row_idx, dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        pick = np.ascontiguousarray(pick, dtype=np.int64)
        h = C.c_void_p()
        _check(lib().mlo_lu_new(size, len(col_ptr) - 1, _pi64(col_ptr), _pi64(row_idx), _pd(vals), _pi64(pick), stability,
                                C.byref(h)))
        self._h, self.size = h, size

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mlo_lu_free(self._h)
            self._h = None

    def nnz(self):
        return int(lib().mlo_lu_nnz(self._h))

    def dense(self, which):
        n = self.size
        out = np.zeros(n if which % 10 == 2 else n * n)
        lib().mlo_lu_get_dense(self._h, which, _pd(out))
        return out if which % 10 == 2 else out.reshape(n, n)

    def perm(self, which):
        out = np.zeros(self.size, dtype=np.int64)
        lib().mlo_lu_get_perm(self._h, which, _pi64(out))
        return out

    def solve_dense(self, rhs, transposed=False):
        r = np.array(rhs, dtype=np.float64)
        lib().mlo_lu_solve_dense(self._h, int(transposed), _pd(r))
        return r

    def solve_sparse(self, idx, val, transposed=False):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        val = np.ascontiguousarray(val, dtype=np.float64)
        out = np.zeros(self.size)
        order = np.zeros(self.size, dtype=np.int64)
        k = lib().mlo_lu_solve_sparse(self._h, int(transposed), len(idx), _pi64(idx), _pd(val), _pd(out), _pi64(order))
        return out, order[:k]


def sparsemat_transpose(n_rows, indptr, indices, data):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    data = np.ascontiguousarray(data, dtype=np.float64)
    n_cols = len(indptr) - 1
    t_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    t_idx = np.zeros(len(indices), dtype=np.int64)
    t_dat = np.zeros(len(indices))
    lib().mlo_sparsemat_transpose(n_rows, n_cols, _pi64(indptr), _pi64(indices), _pd(data), _pi64(t_ptr), _pi64(t_idx),
                                  _pd(t_dat))
    return t_ptr, t_idx, t_dat


def synth_dense(kind, m, n, seed, threads=1):
    """(direction, A, obj, mins, maxs, ops, rhs) from the oracle's own generator (synth_lp.hpp)."""
    a = np.empty((m, n))
    obj, mins, maxs = np.empty(n), np.empty(n), np.empty(n)
    ops, rhs = np.empty(m, dtype=np.int32), np.empty(m)
    d = lib().mlo_synth_dense(kind, m, n, seed, threads, _pd(a), _pd(obj), _pd(mins), _pd(maxs), _pi32(ops), _pd(rhs))
    return d, a, obj, mins, maxs, ops, rhs
