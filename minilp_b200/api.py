"""Host-side mirror of the reference's public API (lib.rs) and thin wrappers of the engine / solver handles."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (Counters, DualEntering, DualRow, Entering, InitState, Leaving, PivotInfo, PivotResult, Profile, pd, pi32,
                   pi64, pu8)

INF = float("inf")


class Error(Exception):
    """lib.rs:171-178"""


class Infeasible(Error):
    """Error::Infeasible"""


class Unbounded(Error):
    """Error::Unbounded"""


class SingularBasis(Error):
    """the reference panics here (solver.rs:316, 1301)"""


class NonFinite(Error):
    """the reference asserts here (solver.rs:1149, 1172)"""


class NoDevice(RuntimeError):
    pass


_STATUS = {1: Infeasible, 2: Unbounded, 3: SingularBasis, 4: NonFinite, 5: ValueError, 6: RuntimeError, 7: NoDevice,
           8: MemoryError}


def _check(rc):
    if rc != 0:
        raise _STATUS.get(rc, RuntimeError)(_lib.lib().mlp_last_error().decode() or f"mlp_status {rc}")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=pd):
    return a.ctypes.data_as(t)


class OptimizationDirection:
    """lib.rs:61-68"""
    Minimize = 0
    Maximize = 1


class ComparisonOp:
    """lib.rs:160-169"""
    Eq = 0
    Le = 1
    Ge = 2


def device_count():
    return int(_lib.lib().mlp_device_count())


def shard_range(n, world, rank):
    b, e = C.c_int64(), C.c_int64()
    _lib.lib().mlp_shard_range(n, world, rank, C.byref(b), C.byref(e))
    return b.value, e.value


def reduce_candidates(scores, pos, vars_):
    scores = _f64(scores)
    pos = np.ascontiguousarray(pos, dtype=np.int64)
    vars_ = np.ascontiguousarray(vars_, dtype=np.int64)
    return int(_lib.lib().mlp_reduce_candidates(_p(scores), _p(pos, pi64), _p(vars_, pi64), len(scores)))


# ------------------------------------------------------------------------------------------------ synthetic LPs
def synth_rows(kind, m, n, seed, row0, nrows, threads=1, out=None):
    if out is None:
        out = np.empty((nrows, n), dtype=np.float64)
    _lib.lib().mlp_synth_rows(kind, m, n, seed, row0, nrows, threads, _p(out))
    return out


def synth_block(kind, m, n, seed, row0, nrows, col0, ncols, threads=1, out=None):
    if out is None:
        out = np.empty((nrows, ncols), dtype=np.float64)
    _lib.lib().mlp_synth_block(kind, m, n, seed, row0, nrows, col0, ncols, threads, _p(out))
    return out


def synth_vectors(kind, m, n, seed):
    obj, mins, maxs = np.empty(n), np.empty(n), np.empty(n)
    ops, rhs = np.empty(m, dtype=np.int32), np.empty(m)
    d = _lib.lib().mlp_synth_vectors(kind, m, n, seed, _p(obj), _p(mins), _p(maxs), _p(ops, pi32), _p(rhs))
    return d, obj, mins, maxs, ops, rhs


class DenseLP:
    """A dense LP in host memory: direction, A (m x n), obj (user sign), bounds, row ops, rhs."""

    def __init__(self, direction, a, obj, mins, maxs, ops, rhs):
        self.direction = direction
        self.a = _f64(a)
        self.obj, self.mins, self.maxs, self.rhs = _f64(obj), _f64(mins), _f64(maxs), _f64(rhs)
        self.ops = np.ascontiguousarray(ops, dtype=np.int32)
        self.m, self.n = self.a.shape


def synth_dense(kind, m, n, seed, threads=1):
    d, obj, mins, maxs, ops, rhs = synth_vectors(kind, m, n, seed)
    return DenseLP(d, synth_rows(kind, m, n, seed, 0, m, threads), obj, mins, maxs, ops, rhs)


# ------------------------------------------------------------------------------------------------ engine
class LocalGroup:
    """In-process rendezvous for `world` logical shards driven by one host thread each (MLP_COMM_LOCAL)."""

    def __init__(self, world):
        h = C.c_void_p()
        _check(_lib.lib().mlp_local_group_create(world, C.byref(h)))
        self._g, self.world = h, world

    def __del__(self):
        if getattr(self, "_g", None) and _lib is not None and getattr(_lib, "lib", None) is not None:
            _lib.lib().mlp_local_group_destroy(self._g)
            self._g = None


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(_lib.lib().mlp_nccl_get_unique_id(buf))
    return buf.raw


class Engine:
    """Borrowed view of an mlp_engine* (owned by a Solver) exposing the per-operation ABI for tests and benches.
    `n` is the number of structural columns THIS shard owns; var-indexed downloads have n + m entries in local order
    (own structural columns, then the m slacks)."""

    def __init__(self, handle, m, n_global):
        self._e, self.m = handle, m
        b, e = C.c_int64(), C.c_int64()
        _check(_lib.lib().mlp_engine_local_range(handle, C.byref(b), C.byref(e)))
        self.col_begin, self.col_end, self.n_global = b.value, e.value, n_global
        self.n = e.value - b.value

    def exchange_kind(self):
        return {0: "none", 1: "nccl_allgather", 2: "in_process", 3: "nvlink_peer_memory"}[int(_lib.lib().mlp_engine_exchange_kind(self._e))]

    def global_ids(self):
        """GLOBAL variable index of every local variable slot."""
        return np.concatenate([np.arange(self.col_begin, self.col_end), self.n_global + np.arange(self.m)])

    def select_entering_primal(self):
        out = Entering()
        _check(_lib.lib().mlp_select_entering_primal(self._e, C.byref(out)))
        return out

    def ftran_col(self, var):
        _check(_lib.lib().mlp_ftran_col(self._e, var))

    def ratio_primal(self, sign, max_step0):
        out = Leaving()
        _check(_lib.lib().mlp_ratio_primal(self._e, int(sign), max_step0, C.byref(out)))
        return out

    def btran_unit(self, row):
        _check(_lib.lib().mlp_btran_unit(self._e, row))

    def price_row(self):
        _check(_lib.lib().mlp_price_row(self._e))

    def calc_row_coeffs(self, row):
        _check(_lib.lib().mlp_calc_row_coeffs(self._e, row))

    def select_row_dual(self):
        out = DualRow()
        _check(_lib.lib().mlp_select_row_dual(self._e, C.byref(out)))
        return out

    def ratio_dual(self, row, leaving_new_val):
        out = DualEntering()
        _check(_lib.lib().mlp_ratio_dual(self._e, row, leaving_new_val, C.byref(out)))
        return out

    def recalc_basic_vals(self):
        """recalc_basic_var_vals (solver.rs:1177-1197): x_B from scratch (row f4)."""
        _check(_lib.lib().mlp_recalc_basic_vals(self._e))

    def refactor(self):
        z = C.c_int64()
        _check(_lib.lib().mlp_refactor(self._e, C.byref(z)))
        return z.value

    def download(self, which):
        n = self.m if which in (3, 4, 5, 6, 8, 9) else self.n + self.m
        out = np.empty(n)
        _check(_lib.lib().mlp_download_f64(self._e, which, _p(out), n))
        return out

    def basic_vars(self):
        out = np.empty(self.m, dtype=np.int64)
        _check(_lib.lib().mlp_download_basic_vars(self._e, _p(out, pi64)))
        return out

    def download_csc(self, nnz):
        """The CSC copy of a sparse-storage engine as built on the device: (col_ptr, row_idx, vals)."""
        cp, ri, va = np.empty(self.n + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        _check(_lib.lib().mlp_engine_download_csc(self._e, _p(cp, pi64), _p(ri, pi32), _p(va)))
        return cp, ri, va

    def var_state(self):
        fl = np.empty(self.n + self.m, dtype=np.uint8)
        pos = np.empty(self.n + self.m, dtype=np.int32)
        _check(_lib.lib().mlp_download_var_state(self._e, _p(fl, pu8), _p(pos, pi32)))
        return fl, pos

    def counters(self):
        c = Counters()
        _check(_lib.lib().mlp_get_counters(self._e, C.byref(c)))
        return {k: getattr(c, k) for k, _ in Counters._fields_}

    def sync(self):
        _check(_lib.lib().mlp_engine_sync(self._e))

    def event_mark(self, slot):
        _check(_lib.lib().mlp_event_mark(self._e, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double()
        _check(_lib.lib().mlp_event_elapsed_ms(self._e, a, b, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        _check(_lib.lib().mlp_profile_enable(self._e, int(on)))

    def profile(self):
        p = Profile()
        _check(_lib.lib().mlp_profile_get(self._e, C.byref(p)))
        return {k: getattr(p, k) for k, _ in Profile._fields_}

    TUNE = {"price_tile": 0, "lane1_ldg": 1, "fused": 2, "fused_max": 3, "price_split": 4, "lu_every": 5}

    def set_tuning(self, knob, value):
        _check(_lib.lib().mlp_engine_set_tuning(self._e, self.TUNE[knob], int(value)))

    def get_tuning(self, knob=None):
        if knob is None:
            return {k: self.get_tuning(k) for k in self.TUNE}
        v = C.c_int32()
        _check(_lib.lib().mlp_engine_get_tuning(self._e, self.TUNE[knob], C.byref(v)))
        return v.value

    def bench_price_dense(self, iters):
        ms, by = C.c_double(), C.c_int64()
        _check(_lib.lib().mlp_bench_price_dense(self._e, iters, C.byref(ms), C.byref(by)))
        return ms.value, by.value


TRACE_FIELDS = ("phase", "entering_var", "entering_col", "leaving_row", "leaving_var", "pivot_coeff", "entering_diff",
                "obj_after", "eta_count", "lu_nnz", "nnz_col", "nnz_rho", "refactored")


class Solver:
    """solver.rs `Solver` after the swap: host control loop (csrc/host_solver.cpp) + device engine."""

    def __init__(self, m, n, device=0, rank=0, world=1, comm=None, csr=None):
        """comm: None (single shard), a LocalGroup, or the 128-byte NCCL unique id shared by all ranks.
        csr: (row_ptr int64[m+1], col_idx int32[nnz], vals f64[nnz]) selects the sparse-storage engine."""
        h = C.c_void_p()
        if csr is not None:
            rp = np.ascontiguousarray(csr[0], dtype=np.int64)
            ci = np.ascontiguousarray(csr[1], dtype=np.int32)
            va = _f64(csr[2])
            assert rp.shape[0] == m + 1 and ci.shape == va.shape
            if world == 1:
                kind, arg = 0, None
            elif isinstance(comm, LocalGroup):
                kind, arg = 2, comm._g
            else:
                buf = C.create_string_buffer(bytes(comm), 128)
                kind, arg = 1, C.cast(buf, C.c_void_p)
            _check(_lib.lib().mlp_solver_create_sparse_sharded(device, m, n, int(va.shape[0]), _p(rp, pi64), _p(ci, pi32), _p(va),
                                                               rank, world, kind, arg, C.byref(h)))
        elif world == 1:
            _check(_lib.lib().mlp_solver_create_dense(device, m, n, C.byref(h)))
        elif isinstance(comm, LocalGroup):
            _check(_lib.lib().mlp_solver_create_dense_sharded(device, m, n, rank, world, 2, comm._g, C.byref(h)))
        else:
            buf = C.create_string_buffer(bytes(comm), 128)
            _check(_lib.lib().mlp_solver_create_dense_sharded(device, m, n, rank, world, 1, C.cast(buf, C.c_void_p),
                                                              C.byref(h)))
        self._s, self.m, self.n, self.rank, self.world = h, m, n, rank, world
        self.engine = Engine(C.c_void_p(_lib.lib().mlp_solver_engine(h)), m, n)

    def close(self):
        if getattr(self, "_s", None) and _lib is not None and getattr(_lib, "lib", None) is not None:
            _lib.lib().mlp_solver_destroy(self._s)
            self._s = None

    def __del__(self):
        self.close()

    def upload_rows(self, row0, rows):
        rows = _f64(rows)
        _check(_lib.lib().mlp_solver_upload_rows(self._s, row0, rows.shape[0], _p(rows)))

    def upload_local_rows(self, row0, rows):
        rows = _f64(rows)
        assert rows.shape[1] == self.engine.n
        _check(_lib.lib().mlp_solver_upload_local_rows(self._s, row0, rows.shape[0], _p(rows)))

    def init(self, obj_internal, mins, maxs, ops, rhs):
        obj_internal, mins, maxs, rhs = _f64(obj_internal), _f64(mins), _f64(maxs), _f64(rhs)
        ops = np.ascontiguousarray(ops, dtype=np.int32)
        _check(_lib.lib().mlp_solver_init(self._s, _p(obj_internal), _p(mins), _p(maxs), _p(ops, pi32), _p(rhs)))

    @classmethod
    def from_dense(cls, lp, device=0, chunk_rows=None, rank=0, world=1, comm=None):
        """Problem::solve's set-up half: stream A from host memory, then Solver::try_new."""
        s = cls(lp.m, lp.n, device, rank, world, comm)
        step = chunk_rows or max(1, (64 << 20) // (8 * lp.n))
        for r0 in range(0, lp.m, step):
            s.upload_rows(r0, lp.a[r0:r0 + step])
        obj = -lp.obj if lp.direction == OptimizationDirection.Maximize else lp.obj  # lib.rs:235-238
        s.init(obj, lp.mins, lp.maxs, lp.ops, lp.rhs)
        s.direction = lp.direction
        return s

    @classmethod
    def from_csr(cls, direction, m, n, row_ptr, col_idx, vals, obj, mins, maxs, ops, rhs, device=0):
        """Sparse storage (BASELINE config 4).  obj in user sign."""
        s = cls(m, n, device, csr=(row_ptr, col_idx, vals))
        obj = _f64(obj)
        s.init(-obj if direction == OptimizationDirection.Maximize else obj, mins, maxs, ops, rhs)
        s.direction = direction
        return s

    def clone(self):
        """Solver: Clone (solver.rs:14) — device-to-device deep copy."""
        h = C.c_void_p()
        _check(_lib.lib().mlp_solver_clone(self._s, C.byref(h)))
        c = object.__new__(Solver)
        c._s, c.m, c.n, c.rank, c.world = h, self.m, self.n, 0, 1
        c.engine = Engine(C.c_void_p(_lib.lib().mlp_solver_engine(h)), self.m, self.n)
        c.direction = getattr(self, "direction", None)
        return c

    def run(self, max_pivots=-1):
        done = C.c_int32(0)
        _check(_lib.lib().mlp_solver_run(self._s, max_pivots, C.byref(done)))
        return bool(done.value)

    cur_obj_val = property(lambda s: float(_lib.lib().mlp_solver_cur_obj_val(s._s)))
    pivots_done = property(lambda s: int(_lib.lib().mlp_solver_pivots_done(s._s)))

    def values(self):
        out = np.zeros(self.n)
        _check(_lib.lib().mlp_solver_values(self._s, _p(out)))
        return out

    def trace(self):
        k = int(_lib.lib().mlp_solver_trace_len(self._s))
        out = np.empty((max(k, 1), 13))
        got = _lib.lib().mlp_solver_get_trace(self._s, 0, k, _p(out))
        return out[:got]

    def set_record_trace(self, on):
        _lib.lib().mlp_solver_set_record_trace(self._s, int(on))

    def set_refactor_factor(self, factor):
        """Refactorize when eta nnz >= factor * lu nnz (solver.rs:1096-1097 is factor = 1, the default)."""
        _lib.lib().mlp_solver_set_refactor_factor(self._s, float(factor))

    def set_recalc_period(self, period):
        """Row f4: recompute x_B and d from scratch every `period` pivots (0 = never, the reference's behaviour)."""
        _lib.lib().mlp_solver_set_recalc_period(self._s, int(period))

    @property
    def recalcs_done(self):
        return int(_lib.lib().mlp_solver_recalcs_done(self._s))

    def nb_vars(self):
        out = np.empty(self.n, dtype=np.int64)
        _check(_lib.lib().mlp_solver_get_nb_vars(self._s, _p(out, pi64)))
        return out

    def basic_vars(self):
        out = np.empty(self.m, dtype=np.int64)
        _check(_lib.lib().mlp_solver_get_basic_vars(self._s, _p(out, pi64)))
        return out

    def tie_stats(self):
        """Pivots whose ratio-test winner was contested: dict(tied_pivots, near_tie_pivots, first_tied_pivot, first_near_tie_pivot)."""
        out = np.zeros(4, dtype=np.int64)
        _lib.lib().mlp_solver_tie_stats(self._s, _p(out, pi64))
        return dict(zip(("tied_pivots", "near_tie_pivots", "first_tied_pivot", "first_near_tie_pivot"), (int(x) for x in out)))

    def timers(self):
        a, b = C.c_double(), C.c_double()
        _lib.lib().mlp_solver_timers(self._s, C.byref(a), C.byref(b))
        return a.value, b.value

    # position-indexed views matching the reference's Solver fields (for parity tests; single-shard engines)
    def nb_var_obj_coeffs(self):
        return self.engine.download(0)[self.nb_vars()]

    def primal_edge_sq_norms(self):
        return self.engine.download(1)[self.nb_vars()]

    def nb_var_vals(self):
        return self.engine.download(2)[self.nb_vars()]

    def basic_var_vals(self):
        return self.engine.download(3)

    def dual_edge_sq_norms(self):
        return self.engine.download(4)


# ------------------------------------------------------------------------------------------------ public API mirror
class Problem:
    """lib.rs:192-305.  Constraints are collected on the host; solve() hands the matrix to the device engine.  A problem that
    comes out of the native MPS reader holds its constraints as flat CSR arrays (`from_arrays`) — the list-of-tuples view
    (`constraints`) is only materialised if somebody asks for it."""

    def __init__(self, direction):
        self.direction = direction
        self.obj_coeffs, self.var_mins, self.var_maxs = [], [], []
        self._constraints = []
        self._bulk = None  # (row_ptr, col_idx, vals, ops, rhs) over ALL constraints, empty ones included

    @classmethod
    def from_arrays(cls, direction, obj_user, mins, maxs, row_ptr, col_idx, vals, ops, rhs):
        """obj_user: objective in the user's sign (add_var applies lib.rs:235-238); rows with ascending variable indices."""
        p = cls(direction)
        sign = 1.0 if direction == OptimizationDirection.Minimize else -1.0
        p.obj_coeffs = (sign * np.asarray(obj_user, dtype=np.float64)).tolist()
        p.var_mins, p.var_maxs = np.asarray(mins, dtype=np.float64).tolist(), np.asarray(maxs, dtype=np.float64).tolist()
        p._bulk = (np.ascontiguousarray(row_ptr, dtype=np.int64), np.ascontiguousarray(col_idx, dtype=np.int32),
                   np.ascontiguousarray(vals, dtype=np.float64), np.ascontiguousarray(ops, dtype=np.int32),
                   np.ascontiguousarray(rhs, dtype=np.float64))
        p._constraints = None
        return p

    @property
    def constraints(self):
        if self._constraints is None:
            rp, ci, va, ops, rhs = self._bulk
            ci_l, va_l = ci.tolist(), va.tolist()
            self._constraints = [(list(zip(ci_l[rp[i]:rp[i + 1]], va_l[rp[i]:rp[i + 1]])), int(ops[i]), float(rhs[i]))
                                 for i in range(len(ops))]
            self._bulk = None
        return self._constraints

    @constraints.setter
    def constraints(self, value):
        self._constraints, self._bulk = value, None

    def add_var(self, obj_coeff, bounds):
        v = len(self.obj_coeffs)
        self.obj_coeffs.append(obj_coeff if self.direction == OptimizationDirection.Minimize else -obj_coeff)  # 235-238
        self.var_mins.append(bounds[0])
        self.var_maxs.append(bounds[1])
        return v

    def add_constraint(self, expr, cmp_op, rhs):
        expr = [(int(v), float(c)) for v, c in expr]
        vs = [v for v, _ in expr]
        if len(set(vs)) != len(vs):
            raise ValueError("variable added more than once to a constraint")  # lib.rs:247-249 panics
        if any(v < 0 or v >= len(self.obj_coeffs) for v in vs):
            raise ValueError("unknown variable")
        self.constraints.append((sorted(expr), cmp_op, float(rhs)))  # CsVec::new sorts by index (lib.rs:279)

    def _all_rows(self):
        """Every constraint as CSR (row_ptr, col_idx, vals, ops, rhs), empty rows included."""
        if self._constraints is None:
            return self._bulk
        cons = self._constraints
        row_ptr = np.zeros(len(cons) + 1, dtype=np.int64)
        for i, (e, _, _) in enumerate(cons):
            row_ptr[i + 1] = row_ptr[i] + len(e)
        col_idx = np.fromiter((v for e, _, _ in cons for v, _ in e), dtype=np.int32, count=int(row_ptr[-1]))
        vals = np.fromiter((c for e, _, _ in cons for _, c in e), dtype=np.float64, count=int(row_ptr[-1]))
        ops = np.array([op for _, op, _ in cons], dtype=np.int32)
        rhs = np.array([r for _, _, r in cons], dtype=np.float64)
        return row_ptr, col_idx, vals, ops, rhs

    def to_csr(self):
        """The constraint rows that survive Solver::try_new's empty-row filter (solver.rs:201-213) as CSR, plus ops / rhs."""
        row_ptr, col_idx, vals, ops, rhs = self._all_rows()
        lens = np.diff(row_ptr)
        keep = lens > 0
        if keep.all():
            return row_ptr, col_idx, vals, ops, rhs
        new_ptr = np.zeros(int(keep.sum()) + 1, dtype=np.int64)
        np.cumsum(lens[keep], out=new_ptr[1:])
        return new_ptr, col_idx, vals, ops[keep], rhs[keep]  # empty rows own no entries: the entry arrays stay as they are

    def solve(self, device=0, max_pivots=-1, storage="auto"):
        """storage: "dense" (row-major f64 A in HBM), "sparse" (CSR + CSC), or "auto" (sparse below 10 % density once A
        has more than 2^20 entries)."""
        n = len(self.obj_coeffs)
        if np.any(np.asarray(self.var_mins) > np.asarray(self.var_maxs)):
            raise Infeasible("min > max")  # solver.rs:138-140
        row_ptr, col_idx, vals, ops, rhs = self._all_rows()
        lens = np.diff(row_ptr)
        empty = lens == 0
        if empty.any():  # solver.rs:201-213
            o, r = ops[empty], rhs[empty]
            ok = np.where(o == ComparisonOp.Eq, 0.0 == r, np.where(o == ComparisonOp.Le, 0.0 <= r, 0.0 >= r))
            if not ok.all():
                raise Infeasible("empty constraint cannot hold")
        m = int((~empty).sum())
        if n == 0 or m == 0:
            return _TrivialSolution(self, n)
        row_ptr, col_idx, vals, ops, rhs = self.to_csr()
        nnz = int(row_ptr[-1])
        if storage == "auto":
            storage = "sparse" if (m * n > (1 << 20) and nnz < 0.1 * m * n) else "dense"
        if storage == "sparse":
            s = Solver(m, n, device, csr=(row_ptr, col_idx, vals))
        else:
            a = np.zeros((m, n))
            a[np.repeat(np.arange(m), np.diff(row_ptr)), col_idx] = vals
            s = Solver(m, n, device)
            s.upload_rows(0, a)
        s.init(np.array(self.obj_coeffs), np.array(self.var_mins), np.array(self.var_maxs), ops, rhs)
        s.direction = self.direction
        s.run(max_pivots)
        return Solution(s, self.direction, n)


class Solution:
    """lib.rs:313-423: objective, var_value, iteration and the incremental methods.  The reference's
    incremental methods consume `self` and return the new solution; here they update in place and return self;
    `clone()` is `Solution: Clone` (lib.rs:313)."""

    def __init__(self, solver, direction, num_vars):
        self.solver, self.direction, self.num_vars = solver, direction, num_vars
        self._vals = solver.values()

    def clone(self):
        return Solution(self.solver.clone(), self.direction, self.num_vars)

    def _refresh(self):
        s = self.solver
        s.m = int(_lib.lib().mlp_solver_num_constraints(s._s))
        s.engine.m = s.m
        self._vals = s.values()
        return self

    def add_constraint(self, expr, cmp_op, rhs):
        """lib.rs:368-382 -> Solver::add_constraint (solver.rs:549-634)."""
        expr = sorted((int(v), float(c)) for v, c in expr)
        vs = [v for v, _ in expr]
        if len(set(vs)) != len(vs):
            raise ValueError("variable added more than once to a constraint")  # CsVec::new panics (lib.rs:376)
        if any(v < 0 or v >= self.num_vars for v in vs):
            raise ValueError("unknown variable")
        va = np.array(vs, dtype=np.int64)
        co = np.array([c for _, c in expr], dtype=np.float64)
        _check(_lib.lib().mlp_solver_add_constraint(self.solver._s, len(vs), _p(va, pi64), _p(co), int(cmp_op), float(rhs)))
        return self._refresh()

    def fix_var(self, var, val):
        """lib.rs:391-395 -> Solver::fix_var (solver.rs:378-415)."""
        assert 0 <= var < self.num_vars
        _check(_lib.lib().mlp_solver_fix_var(self.solver._s, int(var), float(val)))
        return self._refresh()

    def unfix_var(self, var):
        """lib.rs:400-404 -> Solver::unfix_var (solver.rs:418-438).  Returns (solution, was_fixed)."""
        assert 0 <= var < self.num_vars
        was = C.c_int32(0)
        _check(_lib.lib().mlp_solver_unfix_var(self.solver._s, int(var), C.byref(was)))
        return self._refresh(), bool(was.value)

    def add_gomory_cut(self, var):
        """lib.rs:419-423 -> Solver::add_gomory_cut (solver.rs:440-460); the variable must be basic."""
        assert 0 <= var < self.num_vars
        _check(_lib.lib().mlp_solver_add_gomory_cut(self.solver._s, int(var)))
        return self._refresh()

    def objective(self):
        v = self.solver.cur_obj_val
        return v if self.direction == OptimizationDirection.Minimize else -v  # lib.rs:334-339

    def var_value(self, var):
        assert var < self.num_vars
        return float(self._vals[var])

    __getitem__ = var_value

    def __iter__(self):
        return iter(enumerate(self._vals.tolist()))


class _TrivialSolution:
    """No constraints survive try_new (all tautological): the reference's loops make no pivot unless a variable
    can improve without bound (solver.rs:841-844).  The incremental methods of lib.rs:368-423 work here too, as they do on
    the reference's Solution of an unconstrained problem: the first real constraint builds the device engine."""

    def __init__(self, p, n, fixed=None):
        self.direction, self.num_vars = p.direction, n
        self._p = p
        self._fixed = dict(fixed or {})  # var -> value (Solver::fix_var on a non-basic variable, solver.rs:394-411)
        vals, obj = [], 0.0
        for j, (c, mn, mx) in enumerate(zip(p.obj_coeffs, p.var_mins, p.var_maxs)):
            if j in self._fixed:
                x = self._fixed[j]
            elif mn == mx:
                x = mn
            elif c > 0:
                x = mn
            elif c < 0:
                x = mx
            else:
                x = mn if np.isfinite(mn) else (mx if np.isfinite(mx) else 0.0)
            if not np.isfinite(x):
                raise Unbounded("problem is unbounded")
            vals.append(x)
            obj += c * x
        self._vals, self._obj = np.array(vals), obj

    def objective(self):
        return self._obj if self.direction == OptimizationDirection.Minimize else -self._obj

    def var_value(self, var):
        return float(self._vals[var])

    __getitem__ = var_value

    def __iter__(self):
        return iter(enumerate(self._vals.tolist()))

    def clone(self):
        return _TrivialSolution(self._copy_problem(), self.num_vars, self._fixed)

    def _copy_problem(self):
        q = Problem(self._p.direction)
        q.obj_coeffs, q.var_mins, q.var_maxs = list(self._p.obj_coeffs), list(self._p.var_mins), list(self._p.var_maxs)
        q.constraints = list(self._p.constraints)
        return q

    def add_constraint(self, expr, cmp_op, rhs):
        """lib.rs:368-382.  An empty expression is a tautology or infeasible (solver.rs:558-570); anything else makes this a
        constrained problem: it is solved on the device and the variables fixed so far are fixed there again."""
        q = self._copy_problem()
        q.add_constraint(expr, cmp_op, rhs)
        sol = q.solve()
        for v, val in self._fixed.items():
            sol = sol.fix_var(v, val)
        return sol

    def fix_var(self, var, val):
        """lib.rs:391-395 / solver.rs:378-415 for a non-basic variable (every variable is non-basic here)."""
        assert 0 <= var < self.num_vars
        if val < self._p.var_mins[var] or val > self._p.var_maxs[var]:
            raise Infeasible("value outside the variable's bounds")  # solver.rs:379-381
        f = dict(self._fixed)
        f[var] = float(val)
        return _TrivialSolution(self._p, self.num_vars, f)

    def unfix_var(self, var):
        """lib.rs:400-404 / solver.rs:418-438: (solution, was_fixed)."""
        assert 0 <= var < self.num_vars
        if var not in self._fixed:
            return self, False
        f = dict(self._fixed)
        del f[var]
        return _TrivialSolution(self._p, self.num_vars, f), True

    def add_gomory_cut(self, var):
        raise ValueError("var is not basic")  # solver.rs:458 panics: without constraints no variable is basic
