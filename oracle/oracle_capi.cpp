// ORACLE — TEST INFRASTRUCTURE ONLY (see minilp_oracle.hpp header).
// extern "C" surface over the C++ restatement so that tests/, smoke() and the
// CPU-baseline legs of bench.py can drive it through ctypes.
#include "minilp_oracle.hpp"
#include "synth_lp.hpp"

#include <chrono>
#include <cstring>
#include <memory>

using namespace mlo;

namespace {
thread_local std::string g_last_error;

struct Handle {
  // exactly one of these is live
  std::unique_ptr<Solution> sparse;             // faithful storage (Problem/Solution API)
  std::unique_ptr<Solver<DenseMatrix>> dense;   // memory-lean dense storage
  std::vector<double> dense_a;                  // owned copy of A when generated here
  Direction direction = Direction::Minimize;
  bool done = false;
};

template <class F> int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const SolveError& e) {
    g_last_error = e.what();
    return (int)e.code;  // 1 infeasible, 2 unbounded, 4 nonfinite
  } catch (const SingularMatrix& e) {
    g_last_error = e.what();
    return 3;
  } catch (const MpsError& e) {
    g_last_error = e.what();
    return 6;
  } catch (const Panic& e) {
    g_last_error = e.what();
    return 5;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 7;
  }
}

template <class S> void copy_usize(const std::vector<usize>& v, int64_t* out) {
  for (usize i = 0; i < v.size(); ++i) out[i] = (int64_t)v[i];
}

template <class Sv> int64_t get_i64(const Sv& s, int what) {
  switch (what) {
    case 0: return (int64_t)s.num_vars;
    case 1: return (int64_t)s.num_constraints();
    case 2: return s.pivots_done;
    case 3: return s.refactor_count;
    case 4: return s.tie_events;
    case 5: return s.is_primal_feasible;
    case 6: return s.is_dual_feasible;
    case 7: return s.enable_primal_steepest_edge;
    case 8: return s.enable_dual_steepest_edge;
    case 9: return (int64_t)s.trace.size();
    case 10: return (int64_t)s.basis_solver.eta_matrices.len();
    case 11: return (int64_t)s.basis_solver.lu_factors.nnz();
    case 12: return (int64_t)s.nb_vars.size();
    case 13: return (int64_t)s.mat.nnz();
    case 14: return s.tied_pivots;
    case 15: return s.near_tie_pivots;
    case 16: return s.first_tied_pivot;
    case 17: return s.first_near_tie_pivot;
    case 18: return s.sel_near_tie_pivots;
    case 19: return s.first_sel_near_tie_pivot;
    default: return -1;
  }
}

template <class Sv> int get_f64_array(const Sv& s, int what, double* out, int64_t cap) {
  const std::vector<double>* v = nullptr;
  switch (what) {
    case 0: v = &s.basic_var_vals; break;
    case 1: v = &s.nb_var_vals; break;
    case 2: v = &s.nb_var_obj_coeffs; break;
    case 3: v = &s.primal_edge_sq_norms; break;
    case 4: v = &s.dual_edge_sq_norms; break;
    case 5: v = &s.orig_var_mins; break;
    case 6: v = &s.orig_var_maxs; break;
    case 7: v = &s.orig_obj_coeffs; break;
    case 8: v = &s.basic_var_mins; break;
    case 9: v = &s.basic_var_maxs; break;
    case 10: v = &s.orig_rhs; break;
    default: return -1;
  }
  if ((int64_t)v->size() > cap) return -2;
  std::memcpy(out, v->data(), v->size() * sizeof(double));
  return (int)v->size();
}

template <class Sv> int get_i64_array(const Sv& s, int what, int64_t* out, int64_t cap) {
  switch (what) {
    case 0:
      if ((int64_t)s.basic_vars.size() > cap) return -2;
      for (usize i = 0; i < s.basic_vars.size(); ++i) out[i] = (int64_t)s.basic_vars[i];
      return (int)s.basic_vars.size();
    case 1:
      if ((int64_t)s.nb_vars.size() > cap) return -2;
      for (usize i = 0; i < s.nb_vars.size(); ++i) out[i] = (int64_t)s.nb_vars[i];
      return (int)s.nb_vars.size();
    case 2:  // nb_var_states: bit0 at_min, bit1 at_max, bit2 fixed
      if ((int64_t)s.nb_var_states.size() > cap) return -2;
      for (usize i = 0; i < s.nb_var_states.size(); ++i)
        out[i] = (s.nb_var_states[i].at_min ? 1 : 0) | (s.nb_var_states[i].at_max ? 2 : 0) | (s.nb_var_is_fixed[i] ? 4 : 0);
      return (int)s.nb_var_states.size();
    default: return -1;
  }
}
}  // namespace

#define WITH_SOLVER(h, expr_sparse_or_dense)                 \
  do {                                                       \
    if ((h)->sparse) { auto& S = (h)->sparse->solver; expr_sparse_or_dense; } \
    else { auto& S = *(h)->dense; expr_sparse_or_dense; }    \
  } while (0)

extern "C" {

const char* mlo_last_error() { return g_last_error.c_str(); }

// ------------------------------------------------------------ Problem (lib.rs:192-305)
void* mlo_problem_new(int direction) {
  Problem* p = new Problem();
  p->direction = (Direction)direction;
  return p;
}
void mlo_problem_free(void* p) { delete (Problem*)p; }
int64_t mlo_problem_add_var(void* p, double obj, double mn, double mx) { return (int64_t)((Problem*)p)->add_var(obj, mn, mx); }
int mlo_problem_add_constraint(void* p, int64_t nnz, const int64_t* vars, const double* coeffs, int op, double rhs) {
  return guarded([&] {
    std::vector<usize> v(vars, vars + nnz);
    std::vector<double> c(coeffs, coeffs + nnz);
    ((Problem*)p)->add_constraint(v, c, (ComparisonOp)op, rhs);
  });
}
int64_t mlo_problem_num_vars(void* p) { return (int64_t)((Problem*)p)->obj_coeffs.size(); }
int64_t mlo_problem_num_constraints(void* p) { return (int64_t)((Problem*)p)->constraints.size(); }
// Flattened export of a Problem (used to hand MPS-parsed problems to the engine in tests).
int64_t mlo_problem_nnz(void* p) {
  int64_t z = 0;
  for (auto& c : ((Problem*)p)->constraints) z += (int64_t)c.coeffs.indices.size();
  return z;
}
void mlo_problem_export(void* pp, double* obj_internal, double* mins, double* maxs, int64_t* row_ptr, int64_t* col_idx,
                        double* vals, int32_t* ops, double* rhs) {
  Problem* p = (Problem*)pp;
  usize n = p->obj_coeffs.size();
  std::memcpy(obj_internal, p->obj_coeffs.data(), n * 8);
  std::memcpy(mins, p->var_mins.data(), n * 8);
  std::memcpy(maxs, p->var_maxs.data(), n * 8);
  int64_t z = 0;
  row_ptr[0] = 0;
  for (usize i = 0; i < p->constraints.size(); ++i) {
    auto& c = p->constraints[i];
    for (usize k = 0; k < c.coeffs.indices.size(); ++k) { col_idx[z] = (int64_t)c.coeffs.indices[k]; vals[z] = c.coeffs.data[k]; ++z; }
    row_ptr[i + 1] = z;
    ops[i] = (int32_t)c.op;
    rhs[i] = c.rhs;
  }
}

// Problem::solve with an optional pivot budget (max_pivots < 0: run to completion).
// *out receives a handle even when the budget ran out (continue with mlo_continue).
int mlo_problem_solve(void* p, int tie_lowest_index, int64_t max_pivots, void** out) {
  *out = nullptr;
  auto h = std::make_unique<Handle>();
  h->sparse = std::make_unique<Solution>();
  h->direction = ((Problem*)p)->direction;
  SolverOptions o;
  o.tie_lowest_index = tie_lowest_index != 0;
  int rc = guarded([&] {
    problem_begin_solve(*(Problem*)p, *h->sparse, o);
    h->done = h->sparse->solver.initial_solve_budget(max_pivots);
  });
  if (rc == 0) *out = h.release();
  return rc;
}
// Only try_new (solver.rs:108-369), no iterations: for the `initialize` known-answer test.
int mlo_problem_init_only(void* p, void** out) {
  *out = nullptr;
  auto h = std::make_unique<Handle>();
  h->sparse = std::make_unique<Solution>();
  h->direction = ((Problem*)p)->direction;
  int rc = guarded([&] { problem_begin_solve(*(Problem*)p, *h->sparse); });
  if (rc == 0) *out = h.release();
  return rc;
}

// Dense-storage solver over caller-owned row-major A (must outlive the handle).
// obj is the USER objective; direction flips it as Problem::add_var does (lib.rs:235-238).
int mlo_dense_new(int direction, int64_t m, int64_t n, const double* a, const double* obj, const double* mins,
                  const double* maxs, const int32_t* ops, const double* rhs, int tie_lowest_index, int copy_a, void** out) {
  *out = nullptr;
  auto h = std::make_unique<Handle>();
  h->dense = std::make_unique<Solver<DenseMatrix>>();
  h->direction = (Direction)direction;
  h->dense->opts.tie_lowest_index = tie_lowest_index != 0;
  const double* ap = a;
  if (copy_a) { h->dense_a.assign(a, a + (usize)m * (usize)n); ap = h->dense_a.data(); }
  int rc = guarded([&] {
    std::vector<double> o(obj, obj + n), mn(mins, mins + n), mx(maxs, maxs + n), r(rhs, rhs + m);
    if (direction == (int)Direction::Maximize) for (double& x : o) x = -x;
    std::vector<int> op(ops, ops + m);
    solver_init_dense(*h->dense, o, mn, mx, (usize)m, ap, op, r);
  });
  if (rc == 0) *out = h.release();
  return rc;
}
// Same, with the synthetic LP generated in here (for the 50k x 50k CPU baseline: no numpy copy).
int mlo_dense_new_synth(int kind, int64_t m, int64_t n, uint64_t seed, int threads, int tie_lowest_index, void** out) {
  *out = nullptr;
  auto h = std::make_unique<Handle>();
  h->dense = std::make_unique<Solver<DenseMatrix>>();
  h->dense->opts.tie_lowest_index = tie_lowest_index != 0;
  int rc = guarded([&] {
    synth::DenseLP lp;
    synth::generate(kind, (usize)m, (usize)n, seed, threads, lp, h->dense_a);
    h->direction = (Direction)lp.direction;
    std::vector<double> o = lp.obj;
    if (lp.direction == 1) for (double& x : o) x = -x;
    solver_init_dense(*h->dense, o, lp.mins, lp.maxs, (usize)m, h->dense_a.data(), lp.ops, lp.rhs);
  });
  if (rc == 0) *out = h.release();
  return rc;
}

void mlo_free(void* hh) { delete (Handle*)hh; }
void* mlo_clone(void* hh) {  // Solution: Clone, lib.rs:313
  Handle* h = (Handle*)hh;
  if (!h->sparse) return nullptr;
  Handle* c = new Handle();
  c->sparse = std::make_unique<Solution>(*h->sparse);
  c->direction = h->direction;
  c->done = h->done;
  return c;
}

// Continue initial_solve for up to max_pivots more pivots. *done = 1 when finished.
int mlo_continue(void* hh, int64_t max_pivots, int* done) {
  Handle* h = (Handle*)hh;
  int rc = guarded([&] { WITH_SOLVER(h, h->done = S.initial_solve_budget(max_pivots)); });
  *done = h->done ? 1 : 0;
  return rc;
}
// Timed variant for the CPU baseline: seconds spent inside the iteration loop only.
int mlo_continue_timed(void* hh, int64_t max_pivots, int* done, double* seconds) {
  auto t0 = std::chrono::steady_clock::now();
  int rc = mlo_continue(hh, max_pivots, done);
  *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

double mlo_objective(void* hh) {  // lib.rs:334-339
  Handle* h = (Handle*)hh;
  double v = 0;
  WITH_SOLVER(h, v = S.cur_obj_val);
  return h->direction == Direction::Minimize ? v : -v;
}
double mlo_cur_obj_val(void* hh) { double v = 0; WITH_SOLVER((Handle*)hh, v = S.cur_obj_val); return v; }
double mlo_var_value(void* hh, int64_t var) { double v = 0; WITH_SOLVER((Handle*)hh, v = S.get_value((usize)var)); return v; }
void mlo_set_record_trace(void* hh, int on) { WITH_SOLVER((Handle*)hh, S.opts.record_trace = on != 0); }

int mlo_add_constraint(void* hh, int64_t nnz, const int64_t* vars, const double* coeffs, int op, double rhs) {
  Handle* h = (Handle*)hh;
  if (!h->sparse) return 5;
  return guarded([&] {
    std::vector<usize> v(vars, vars + nnz);
    std::vector<double> c(coeffs, coeffs + nnz);
    h->sparse->add_constraint(v, c, (ComparisonOp)op, rhs);
  });
}
int mlo_fix_var(void* hh, int64_t var, double val) {
  Handle* h = (Handle*)hh;
  if (!h->sparse) return 5;
  return guarded([&] { h->sparse->fix_var((usize)var, val); });
}
int mlo_unfix_var(void* hh, int64_t var, int* was_fixed) {
  Handle* h = (Handle*)hh;
  if (!h->sparse) return 5;
  return guarded([&] { *was_fixed = h->sparse->unfix_var((usize)var) ? 1 : 0; });
}
int mlo_add_gomory_cut(void* hh, int64_t var) {
  Handle* h = (Handle*)hh;
  if (!h->sparse) return 5;
  return guarded([&] { h->sparse->add_gomory_cut((usize)var); });
}

int64_t mlo_get_i64(void* hh, int what) { int64_t v = -1; WITH_SOLVER((Handle*)hh, v = get_i64(S, what)); return v; }
int mlo_get_f64_array(void* hh, int what, double* out, int64_t cap) { int v = -1; WITH_SOLVER((Handle*)hh, v = get_f64_array(S, what, out, cap)); return v; }
int mlo_get_i64_array(void* hh, int what, int64_t* out, int64_t cap) { int v = -1; WITH_SOLVER((Handle*)hh, v = get_i64_array(S, what, out, cap)); return v; }

// Dense copy of [A|I] row by row (small problems only), for solver.rs:1421-1427.
int mlo_get_constraints_dense(void* hh, double* out /* rows x total */) {
  Handle* h = (Handle*)hh;
  WITH_SOLVER(h, {
    usize total = S.num_total_vars();
    for (usize r = 0; r < S.num_constraints(); ++r) S.mat.for_row(r, [&](usize v, double val) { out[r * total + v] = val; });
  });
  return 0;
}

// trace: 13 values per record as doubles (integers are exact below 2^53)
int64_t mlo_get_trace(void* hh, int64_t first, int64_t count, double* out) {
  int64_t n = 0;
  WITH_SOLVER((Handle*)hh, {
    for (int64_t i = first; i < first + count && i < (int64_t)S.trace.size(); ++i, ++n) {
      const PivotRecord& r = S.trace[(usize)i];
      double* o = out + n * 13;
      o[0] = r.phase; o[1] = (double)r.entering_var; o[2] = (double)r.entering_col; o[3] = (double)r.leaving_row;
      o[4] = (double)r.leaving_var; o[5] = r.pivot_coeff; o[6] = r.entering_diff; o[7] = r.obj_after;
      o[8] = (double)r.eta_count; o[9] = (double)r.lu_nnz; o[10] = (double)r.nnz_col; o[11] = (double)r.nnz_rho; o[12] = r.refactored;
    }
  });
  return n;
}

// Per-operation probes for kernel-level parity tests (state is left as the call leaves it).
// FTRAN of non-basic position `col` (calc_col_coeffs, solver.rs:671): dense m-vector out.
int mlo_probe_ftran_col(void* hh, int64_t col, double* out_dense) {
  return guarded([&] {
    WITH_SOLVER((Handle*)hh, {
      S.calc_col_coeffs((usize)col);
      std::fill(out_dense, out_dense + S.num_constraints(), 0.0);
      for (usize k = 0; k < S.col_coeffs.len(); ++k) out_dense[S.col_coeffs.indices[k]] = S.col_coeffs.values[k];
    });
  });
}
// BTRAN of e_row + price-out (calc_row_coeffs, solver.rs:680): rho (m) and row_coeffs by non-basic position.
int mlo_probe_row_coeffs(void* hh, int64_t row, double* rho_dense, double* row_coeffs_by_pos) {
  return guarded([&] {
    WITH_SOLVER((Handle*)hh, {
      S.calc_row_coeffs((usize)row);
      std::fill(rho_dense, rho_dense + S.num_constraints(), 0.0);
      for (usize k = 0; k < S.inv_basis_row_coeffs.len(); ++k) rho_dense[S.inv_basis_row_coeffs.indices[k]] = S.inv_basis_row_coeffs.values[k];
      std::fill(row_coeffs_by_pos, row_coeffs_by_pos + S.nb_vars.size(), 0.0);
      for (usize c : S.row_coeffs.nonzero) row_coeffs_by_pos[c] = S.row_coeffs.values[c];
    });
  });
}

// ------------------------------------------------------------ MPS (mps.rs:39-329)
int mlo_parse_mps(const char* text, int64_t len, int direction, void** problem_out, void** mps_out) {
  *problem_out = nullptr;
  *mps_out = nullptr;
  auto mf = std::make_unique<MpsFile>();
  int rc = guarded([&] { *mf = parse_mps(std::string(text, (usize)len), (Direction)direction); });
  if (rc != 0) return rc;
  *problem_out = new Problem(mf->problem);
  *mps_out = mf.release();
  return 0;
}
void mlo_mps_free(void* m) { delete (MpsFile*)m; }
const char* mlo_mps_name(void* m) { return ((MpsFile*)m)->problem_name.c_str(); }
int64_t mlo_mps_num_vars(void* m) { return (int64_t)((MpsFile*)m)->var_names.size(); }
const char* mlo_mps_var_name(void* m, int64_t i) { return ((MpsFile*)m)->var_names[(usize)i].c_str(); }
int64_t mlo_mps_var_index(void* m, const char* name) {
  auto& mp = ((MpsFile*)m)->variables;
  auto it = mp.find(name);
  return it == mp.end() ? -1 : (int64_t)it->second;
}

// ------------------------------------------------------------ LU probes (lu.rs tests 480-704)
struct LuHandle { LUFactors lu, lut; ScratchSpace scratch; usize size; };
int mlo_lu_new(int64_t size, int64_t ncols_mat, const int64_t* col_ptr, const int64_t* row_idx, const double* vals,
               const int64_t* pick, double stability, void** out) {
  *out = nullptr;
  auto h = std::make_unique<LuHandle>();
  h->size = (usize)size;
  h->scratch = ScratchSpace((usize)size);
  std::vector<usize> ptr(col_ptr, col_ptr + ncols_mat + 1), idx(row_idx, row_idx + col_ptr[ncols_mat]);
  std::vector<double> val(vals, vals + col_ptr[ncols_mat]);
  CscCols cols{&ptr, &idx, &val, std::vector<usize>(pick, pick + size)};
  int rc = guarded([&] {
    h->lu = lu_factorize((usize)size, cols, stability, h->scratch);
    h->lut = h->lu.transpose();
  });
  if (rc == 0) *out = h.release();
  return rc;
}
void mlo_lu_free(void* h) { delete (LuHandle*)h; }
int64_t mlo_lu_nnz(void* h) { return (int64_t)((LuHandle*)h)->lu.nnz(); }
// which: 0 L nondiag, 1 U nondiag (dense row-major size x size), 2 U diag, transposed factors: +10
void mlo_lu_get_dense(void* hh, int which, double* out) {
  LuHandle* h = (LuHandle*)hh;
  const LUFactors& f = which >= 10 ? h->lut : h->lu;
  int w = which % 10;
  usize n = h->size;
  if (w == 2) { const TriangleMat& t = f.upper.has_diag ? f.upper : f.lower; for (usize i = 0; i < n; ++i) out[i] = t.diag[i]; return; }
  const SparseMat& m = (w == 0) ? f.lower.nondiag : f.upper.nondiag;
  std::fill(out, out + n * n, 0.0);
  for (usize c = 0; c < m.cols(); ++c)
    for (usize p = m.indptr[c]; p < m.indptr[c + 1]; ++p) out[m.indices[p] * n + c] = m.data[p];
}
// which: 0 row_perm.new2orig, 1 row_perm.orig2new, 2 col_perm.new2orig, 3 col_perm.orig2new
void mlo_lu_get_perm(void* hh, int which, int64_t* out) {
  LuHandle* h = (LuHandle*)hh;
  const std::vector<usize>& v = which == 0 ? h->lu.row_perm.new2orig : which == 1 ? h->lu.row_perm.orig2new
                                : which == 2 ? h->lu.col_perm.new2orig : h->lu.col_perm.orig2new;
  for (usize i = 0; i < v.size(); ++i) out[i] = (int64_t)v[i];
}
void mlo_lu_solve_dense(void* hh, int transposed, double* rhs_inout) {
  LuHandle* h = (LuHandle*)hh;
  std::vector<double> r(rhs_inout, rhs_inout + h->size);
  (transposed ? h->lut : h->lu).solve_dense(r, h->scratch);
  std::memcpy(rhs_inout, r.data(), h->size * 8);
}
// sparse solve: input (idx,val) pairs in order; output dense values + the `nonzero` order list.
int64_t mlo_lu_solve_sparse(void* hh, int transposed, int64_t nnz, const int64_t* idx, const double* val, double* out_dense,
                            int64_t* out_order) {
  LuHandle* h = (LuHandle*)hh;
  ScatteredVec rhs(h->size);
  rhs.begin_set();
  for (int64_t k = 0; k < nnz; ++k) rhs.put((usize)idx[k], val[k]);
  (transposed ? h->lut : h->lu).solve(rhs, h->scratch);
  std::memcpy(out_dense, rhs.values.data(), h->size * 8);
  for (usize k = 0; k < rhs.nonzero.size(); ++k) out_order[k] = (int64_t)rhs.nonzero[k];
  return (int64_t)rhs.nonzero.size();
}

// SparseMat::transpose probe (sparse.rs:345-359)
void mlo_sparsemat_transpose(int64_t n_rows, int64_t n_cols, const int64_t* indptr, const int64_t* indices, const double* data,
                             int64_t* t_indptr, int64_t* t_indices, double* t_data) {
  SparseMat m((usize)n_rows);
  for (int64_t c = 0; c < n_cols; ++c) {
    for (int64_t p = indptr[c]; p < indptr[c + 1]; ++p) m.push((usize)indices[p], data[p]);
    m.seal_column();
  }
  SparseMat t = m.transpose();
  for (usize i = 0; i < t.indptr.size(); ++i) t_indptr[i] = (int64_t)t.indptr[i];
  for (usize i = 0; i < t.indices.size(); ++i) { t_indices[i] = (int64_t)t.indices[i]; t_data[i] = t.data[i]; }
}

// ------------------------------------------------------------ synthetic LPs (oracle's own generator)
// Fills caller arrays: a (m*n), obj (n), mins (n), maxs (n), ops (m), rhs (m); returns direction.
int mlo_synth_dense(int kind, int64_t m, int64_t n, uint64_t seed, int threads, double* a, double* obj, double* mins,
                    double* maxs, int32_t* ops, double* rhs) {
  synth::DenseLP lp;
  std::vector<double> A;
  synth::generate(kind, (usize)m, (usize)n, seed, threads, lp, A);
  std::memcpy(a, A.data(), A.size() * 8);
  std::memcpy(obj, lp.obj.data(), n * 8);
  std::memcpy(mins, lp.mins.data(), n * 8);
  std::memcpy(maxs, lp.maxs.data(), n * 8);
  for (int64_t i = 0; i < m; ++i) { ops[i] = lp.ops[i]; rhs[i] = lp.rhs[i]; }
  return lp.direction;
}

}  // extern "C"
