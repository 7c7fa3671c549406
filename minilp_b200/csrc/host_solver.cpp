// Host control loop: the part of ztlpn/minilp's `Solver` that stays on the CPU after the swap —
// problem set-up (Solver::try_new, solver.rs:108-369), phase logic (initial_solve 470-485), the two
// iteration loops (optimize 487-511, restore_feasibility 513-547), the scalar glue of choose_pivot
// (741-748, 825-852) and of pivot (1027, 1096-1103).  All bulk work goes through the engine's C ABI
// (include/minilp_b200.h) and nothing else: this file is what the Rust `Solver` looks like once its
// vector loops are replaced by `extern "C"` calls.
#include "minilp_b200.h"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {
constexpr double kInf = std::numeric_limits<double>::infinity();
using Clock = std::chrono::steady_clock;

struct PivotRecord {
  int32_t phase;
  int64_t entering_var, entering_col, leaving_row, leaving_var;
  double pivot_coeff, entering_diff, obj_after;
  int64_t eta_count, lu_nnz, nnz_col, nnz_rho;
  int32_t refactored;
};
}  // namespace

struct mlp_solver {
  mlp_engine* eng = nullptr;
  int64_t m = 0, n = 0;
  // host mirrors of the O(1)-per-pivot fields of Solver (solver.rs:15-58)
  std::vector<double> orig_var_mins, orig_var_maxs, orig_obj_coeffs;
  std::vector<int64_t> nb_vars, basic_vars;
  std::vector<double> nb_var_vals;
  bool is_primal_feasible = false, is_dual_feasible = false;
  bool enable_primal_steepest_edge = false, enable_dual_steepest_edge = false;
  double cur_obj_val = 0.0;
  int64_t eta_nnz = 0, lu_nnz = 0;  // eta_matrices.coeff_cols.nnz(), lu_factors.nnz() (solver.rs:1096-1097)
  int stage = 0;                    // initial_solve state machine: 0 start, 1 dual, 2 recalc, 3 primal, 4 done
  int64_t pivots_done = 0;
  bool record_trace = true;
  std::vector<PivotRecord> trace;
  double run_seconds = 0.0, refactor_seconds = 0.0;
  // pivots whose ratio-test winner (pass 2 of 804-823 / 982-1002) was tied exactly / within 1e-9: where the reference's
  // list-order rule and the engine's lowest-index rule can part ways
  int64_t tied_pivots = 0, near_tie_pivots = 0, first_tied_pivot = -1, first_near_tie_pivot = -1;
  // f4: every recalc_period pivots (0 = never, the reference's behaviour) recompute x_B and d from scratch — the TODO at
  // solver.rs:1024-1025 — with recalc_basic_var_vals (1177-1197) and recalc_obj_coeffs (1199-1231)
  int64_t recalc_period = 0, recalcs_done = 0;
  // Refactor rule.  The reference refactorizes when the eta file holds as many entries as the LU factors (solver.rs:1096-1097:
  // the point where ITS solves cost twice a fresh factorization's).  refactor_factor scales the right-hand side: 1 (default) is
  // the reference's rule; a larger value lets the eta file grow longer — on the device an eta column is one more column
  // of a coalesced m x K pass, while a refactorization is O(k) sequential pivot steps plus O(k^3) work, so the balance
  // point lies elsewhere.  Only the rounding of later pivots depends on it.
  double refactor_factor = 1.0;
  // the first half of a dual iteration as one engine call / one host round trip (mlp_dual_select_ratio); MLP_FUSED_DUAL=0: the
  // three calls of the reference's control flow
  bool fused_dual = [] { const char* v = getenv("MLP_FUSED_DUAL"); return !(v && atoi(v) == 0); }();
  bool artificial_obj = false;  // solver.rs:261: while the artificial objective is in place d must not be recomputed from c
  bool initialized = false;
};

static void note_ties(mlp_solver* s, int64_t ties, int64_t near_ties) {
  if (ties > 0) { s->tied_pivots += 1; if (s->first_tied_pivot < 0) s->first_tied_pivot = s->pivots_done; }
  if (near_ties > 0) { s->near_tie_pivots += 1; if (s->first_near_tie_pivot < 0) s->first_near_tie_pivot = s->pivots_done; }
}

#define ST(x)                        \
  do {                               \
    mlp_status st__ = (x);           \
    if (st__ != MLP_OK) return st__; \
  } while (0)

// Solver::pivot, host half (solver.rs:1023-1104): objective, mirrors, refactor rule, trace.
static mlp_status do_pivot(mlp_solver* s, int phase, int64_t entering_var, int64_t col, double obj_coeff, double entering_new_val,
                           double entering_diff, bool has_elem, int64_t row, double coeff, double leaving_new_val) {
  s->cur_obj_val += obj_coeff * entering_diff;  // 1027
  mlp_pivot_info pi;
  std::memset(&pi, 0, sizeof(pi));
  pi.entering_var = entering_var;
  pi.col = col;
  pi.entering_obj_coeff = obj_coeff;
  pi.entering_new_val = entering_new_val;
  pi.entering_diff = entering_diff;
  pi.has_elem = has_elem ? 1 : 0;
  pi.row = row;
  pi.coeff = coeff;
  pi.leaving_new_val = leaving_new_val;
  const bool keep_etas = s->refactor_factor == 1.0 ? s->eta_nnz < s->lu_nnz  // 1096-1103
                                                   : (double)s->eta_nnz < s->refactor_factor * (double)s->lu_nnz;
  pi.refactor = (has_elem && !keep_etas) ? 1 : 0;
  mlp_pivot_result pr;
  auto t0 = Clock::now();
  ST(mlp_pivot(s->eng, &pi, &pr));
  if (!has_elem) {
    s->nb_var_vals[col] = entering_new_val;  // 1034
  } else {
    s->nb_var_vals[col] = leaving_new_val;  // 1068
    s->basic_vars[row] = entering_var;      // 1088-1091
    s->nb_vars[col] = pr.leaving_var;
    if (pr.refactored) {
      s->eta_nnz = 0;
      s->lu_nnz = pr.lu_nnz;
      s->refactor_seconds += std::chrono::duration<double>(Clock::now() - t0).count();
    } else {
      s->eta_nnz += pr.col_nnz;
    }
  }
  s->pivots_done += 1;
  if (s->record_trace) {
    PivotRecord r;
    r.phase = phase;
    r.entering_var = entering_var;
    r.entering_col = col;
    r.leaving_row = has_elem ? row : -1;
    r.leaving_var = has_elem ? pr.leaving_var : -1;
    r.pivot_coeff = has_elem ? coeff : 0.0;
    r.entering_diff = entering_diff;
    r.obj_after = s->cur_obj_val;
    r.eta_count = pr.eta_count;
    r.lu_nnz = s->lu_nnz;
    r.nnz_col = pr.col_nnz;
    r.nnz_rho = 0;
    r.refactored = pr.refactored;
    s->trace.push_back(r);
  }
  return MLP_OK;
}

// f4, between two iterations of either loop
static mlp_status maybe_recalc(mlp_solver* s, bool dual_loop) {
  if (s->recalc_period <= 0 || s->pivots_done == 0 || s->pivots_done % s->recalc_period != 0) return MLP_OK;
  ST(mlp_recalc_basic_vals(s->eng));
  if (!(dual_loop && s->artificial_obj)) ST(mlp_recalc_obj_coeffs(s->eng, &s->cur_obj_val));
  mlp_counters c;
  mlp_get_counters(s->eng, &c);
  s->eta_nnz = 0;  // both refactorize whenever etas exist (1188-1191, 1200-1203)
  s->lu_nnz = c.lu_nnz;
  s->recalcs_done += 1;
  return MLP_OK;
}

// One iteration of optimize() (solver.rs:497-498): choose_pivot (695-853) + pivot. *moved = 0 at the optimum.
static mlp_status primal_iteration(mlp_solver* s, int* moved) {
  mlp_entering en;
  ST(mlp_select_entering_primal(s->eng, &en));
  if (en.var < 0) { *moved = 0; return MLP_OK; }  // 737
  const double entering_cur_val = en.cur_val;                                          // 741
  const bool entering_diff_sign = en.obj_coeff < 0.0;                                  // 743
  const double entering_other_val = entering_diff_sign ? s->orig_var_maxs[en.var] : s->orig_var_mins[en.var];  // 744-748
  ST(mlp_ftran_col(s->eng, en.var));                                                   // 750
  mlp_leaving lv;
  ST(mlp_ratio_primal(s->eng, entering_diff_sign ? 1 : 0, std::fabs(entering_other_val - entering_cur_val), &lv));  // 782-823
  if (lv.row >= 0) {
    note_ties(s, lv.ties, lv.near_ties);
    ST(mlp_calc_row_coeffs(s->eng, lv.row));                                           // 826
    const double entering_diff = (lv.basic_val - lv.leaving_new_val) / lv.coeff;       // 828
    const double entering_new_val = entering_cur_val + entering_diff;                  // 829
    ST(do_pivot(s, 1, en.var, en.pos, en.obj_coeff, entering_new_val, entering_diff, true, lv.row, lv.coeff, lv.leaving_new_val));
  } else {
    if (std::isinf(entering_other_val)) return MLP_UNBOUNDED;                          // 842-844
    ST(do_pivot(s, 1, en.var, en.pos, en.obj_coeff, entering_other_val, entering_other_val - entering_cur_val, false, -1, 0.0, 0.0));
  }
  *moved = 1;
  return MLP_OK;
}

// One iteration of restore_feasibility() (solver.rs:529-533).
static mlp_status dual_iteration(mlp_solver* s, int* moved) {
  mlp_dual_row dr;
  mlp_dual_entering de;
  double leaving_new_val;  // 908-915
  if (s->fused_dual) {
    // choose_pivot_row_dual -> calc_row_coeffs -> choose_entering_col_dual queued back to back on the device: one round trip
    ST(mlp_dual_select_ratio(s->eng, &dr, &de));
    if (dr.row < 0) { *moved = 0; return MLP_OK; }
    if (dr.val < dr.min) leaving_new_val = dr.min;
    else if (dr.val > dr.max) leaving_new_val = dr.max;
    else return MLP_INVALID;  // unreachable!() in the reference
  } else {
    ST(mlp_select_row_dual(s->eng, &dr));
    if (dr.row < 0) { *moved = 0; return MLP_OK; }
    if (dr.val < dr.min) leaving_new_val = dr.min;
    else if (dr.val > dr.max) leaving_new_val = dr.max;
    else return MLP_INVALID;  // unreachable!() in the reference
    ST(mlp_calc_row_coeffs(s->eng, dr.row));  // 530
    ST(mlp_ratio_dual(s->eng, dr.row, leaving_new_val, &de));  // 531
  }
  if (de.var < 0) return MLP_INFEASIBLE;                      // 1019
  note_ties(s, de.ties, de.near_ties);
  const double entering_diff = (dr.val - leaving_new_val) / de.coeff;  // 1005
  const double entering_new_val = de.cur_val + entering_diff;          // 1006
  ST(mlp_ftran_col(s->eng, de.var));                                   // 532
  ST(do_pivot(s, 0, de.var, de.pos, de.obj_coeff, entering_new_val, entering_diff, true, dr.row, de.coeff, leaving_new_val));
  *moved = 1;
  return MLP_OK;
}

extern "C" {

mlp_status mlp_solver_create_dense(int device, int64_t m, int64_t n, mlp_solver** out) {
  *out = nullptr;
  mlp_engine* e = nullptr;
  ST(mlp_engine_create_dense(device, m, n, &e));
  mlp_solver* s = new mlp_solver();
  s->eng = e;
  s->m = m;
  s->n = n;
  *out = s;
  return MLP_OK;
}
mlp_status mlp_solver_create_dense_sharded(int device, int64_t m, int64_t n_global, int32_t rank, int32_t world,
                                           int32_t comm_kind, const void* comm_arg, mlp_solver** out) {
  *out = nullptr;
  mlp_engine* e = nullptr;
  ST(mlp_engine_create_dense_sharded(device, m, n_global, rank, world, comm_kind, comm_arg, &e));
  mlp_solver* s = new mlp_solver();
  s->eng = e;
  s->m = m;
  s->n = n_global;  // the control loop works on GLOBAL indices; sharding is an engine-internal property
  *out = s;
  return MLP_OK;
}
mlp_status mlp_solver_create_sparse(int device, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr, const int32_t* col_idx,
                                    const double* vals, mlp_solver** out) {
  *out = nullptr;
  mlp_engine* e = nullptr;
  ST(mlp_engine_create_sparse(device, m, n, nnz, row_ptr, col_idx, vals, &e));
  mlp_solver* s = new mlp_solver();
  s->eng = e;
  s->m = m;
  s->n = n;
  *out = s;
  return MLP_OK;
}
mlp_status mlp_solver_create_sparse_sharded(int device, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
                                            const int32_t* col_idx, const double* vals, int32_t rank, int32_t world,
                                            int32_t comm_kind, const void* comm_arg, mlp_solver** out) {
  *out = nullptr;
  mlp_engine* e = nullptr;
  ST(mlp_engine_create_sparse_sharded(device, m, n, nnz, row_ptr, col_idx, vals, rank, world, comm_kind, comm_arg, &e));
  mlp_solver* s = new mlp_solver();
  s->eng = e;
  s->m = m;
  s->n = n;  // GLOBAL indices in the control loop
  *out = s;
  return MLP_OK;
}
mlp_status mlp_solver_upload_local_rows(mlp_solver* s, int64_t row0, int64_t nrows, const double* rows_local) {
  return mlp_engine_upload_local_rows(s->eng, row0, nrows, rows_local);
}
void mlp_solver_destroy(mlp_solver* s) {
  if (!s) return;
  mlp_engine_destroy(s->eng);
  delete s;
}
mlp_engine* mlp_solver_engine(mlp_solver* s) { return s ? s->eng : nullptr; }
mlp_status mlp_solver_upload_rows(mlp_solver* s, int64_t row0, int64_t nrows, const double* rows_host) {
  return mlp_engine_upload_rows(s->eng, row0, nrows, rows_host);
}

// Solver::try_new, solver.rs:108-369.  (Dense rows are never empty, so the tautology branch 201-213 does not arise.)
mlp_status mlp_solver_init(mlp_solver* s, const double* obj, const double* mins, const double* maxs, const int32_t* ops,
                           const double* rhs) {
  if (!s) return MLP_INVALID;
  const int64_t n = s->n, m = s->m, nt = n + m;
  s->orig_var_mins.assign(mins, mins + n);
  s->orig_var_maxs.assign(maxs, maxs + n);
  s->orig_obj_coeffs.assign(obj, obj + n);
  s->orig_obj_coeffs.resize(nt, 0.0);  // 244-245
  s->nb_vars.resize(n);
  s->nb_var_vals.resize(n);
  std::vector<uint8_t> nb_states(n);
  double obj_val = 0.0;
  bool is_dual_feasible = true;
  for (int64_t v = 0; v < n; ++v) {  // 133-187
    const double mn = mins[v], mx = maxs[v];
    if (mn > mx) return MLP_INFEASIBLE;  // 138-140
    s->nb_vars[v] = v;
    double init_val;
    if (mn == mx) init_val = mn;
    else if (std::isinf(mn) && std::isinf(mx)) { if (obj[v] != 0.0) is_dual_feasible = false; init_val = 0.0; }
    else if (obj[v] > 0.0) { if (std::isfinite(mn)) init_val = mn; else { is_dual_feasible = false; init_val = mx; } }
    else if (obj[v] < 0.0) { if (std::isfinite(mx)) init_val = mx; else { is_dual_feasible = false; init_val = mn; } }
    else if (std::isfinite(mn)) init_val = mn;
    else init_val = mx;
    s->nb_var_vals[v] = init_val;
    obj_val += init_val * obj[v];
    nb_states[v] = (uint8_t)((init_val == mn ? MLP_AT_MIN : 0u) | (init_val == mx ? MLP_AT_MAX : 0u));
  }
  std::vector<double> bmin(m), bmax(m);
  s->basic_vars.resize(m);
  for (int64_t i = 0; i < m; ++i) {  // 218-232
    double smin, smax;
    if (ops[i] == 1) { smin = 0.0; smax = kInf; }
    else if (ops[i] == 2) { smin = -kInf; smax = 0.0; }
    else { smin = 0.0; smax = 0.0; }
    s->orig_var_mins.push_back(smin);
    s->orig_var_maxs.push_back(smax);
    bmin[i] = smin;
    bmax[i] = smax;
    s->basic_vars[i] = n + i;
  }
  // basic_var_vals = rhs - A x_N (234-238) is a pass over A: done on the device, then read back once
  // to decide primal feasibility (255-259).
  mlp_init_state st;
  std::memset(&st, 0, sizeof(st));
  std::vector<double> nb_obj(n, 0.0);
  st.orig_var_mins = s->orig_var_mins.data();
  st.orig_var_maxs = s->orig_var_maxs.data();
  st.orig_obj_coeffs = s->orig_obj_coeffs.data();
  st.orig_rhs = rhs;
  st.nb_vars = s->nb_vars.data();
  st.nb_var_vals = s->nb_var_vals.data();
  st.nb_var_obj_coeffs = nb_obj.data();
  st.nb_var_states = nb_states.data();
  st.primal_edge_sq_norms = nullptr;
  st.basic_vars = s->basic_vars.data();
  st.basic_var_vals = nullptr;
  st.basic_var_mins = bmin.data();
  st.basic_var_maxs = bmax.data();
  st.dual_edge_sq_norms = nullptr;
  const bool enable_steepest_edge = true;  // 114
  s->enable_dual_steepest_edge = enable_steepest_edge;
  s->enable_primal_steepest_edge = enable_steepest_edge && !is_dual_feasible;  // 272
  st.enable_primal_steepest_edge = s->enable_primal_steepest_edge;
  st.enable_dual_steepest_edge = s->enable_dual_steepest_edge;
  // First pass: upload with provisional reduced costs; need basic values to know whether the artificial
  // objective is required (261), which only changes nb_var_obj_coeffs.
  for (int64_t v = 0; v < n; ++v) nb_obj[v] = obj[v];
  ST(mlp_engine_init_state(s->eng, &st));
  std::vector<double> xb(m);
  ST(mlp_download_f64(s->eng, MLP_ARR_BASIC_VALS, xb.data(), m));
  bool is_primal_feasible = true;
  for (int64_t i = 0; i < m; ++i)
    if (!(xb[i] >= bmin[i] && xb[i] <= bmax[i])) is_primal_feasible = false;  // 255-259
  const bool need_artificial_obj = !is_primal_feasible && !is_dual_feasible;   // 261
  if (need_artificial_obj) {  // 284-292
    for (int64_t v = 0; v < n; ++v) {
      const bool at_min = nb_states[v] & MLP_AT_MIN, at_max = nb_states[v] & MLP_AT_MAX;
      nb_obj[v] = (at_min && !at_max) ? 1.0 : (at_max && !at_min) ? -1.0 : 0.0;
    }
    st.basic_var_vals = xb.data();
    ST(mlp_engine_init_state(s->eng, &st));
  }
  s->cur_obj_val = need_artificial_obj ? 0.0 : obj_val;  // 302
  s->artificial_obj = need_artificial_obj;
  s->is_primal_feasible = is_primal_feasible;
  s->is_dual_feasible = is_dual_feasible;
  s->eta_nnz = 0;
  ST(mlp_refactor(s->eng, &s->lu_nnz));  // value of lu_factors.nnz() for the slack basis
  s->stage = 0;
  s->pivots_done = 0;
  s->tied_pivots = s->near_tie_pivots = 0;
  s->first_tied_pivot = s->first_near_tie_pivot = -1;
  s->trace.clear();
  s->initialized = true;
  return MLP_OK;
}

// Solver::initial_solve (solver.rs:470-485) as a resumable state machine with a pivot budget.
mlp_status mlp_solver_run(mlp_solver* s, int64_t max_pivots, int32_t* done) {
  if (!s || !s->initialized) return MLP_INVALID;
  *done = 0;
  const auto t0 = Clock::now();
  const int64_t target = max_pivots < 0 ? -1 : s->pivots_done + max_pivots;
  auto budget_left = [&] { return target < 0 || s->pivots_done < target; };
  mlp_status rc = MLP_OK;
  for (bool go = true; go && rc == MLP_OK;) {
    switch (s->stage) {
      case 0: s->stage = s->is_primal_feasible ? 2 : 1; break;
      case 1: {
        while (budget_left()) {
          int moved = 0;
          rc = dual_iteration(s, &moved);
          if (rc != MLP_OK) break;
          if (!moved) { s->is_primal_feasible = true; s->stage = 2; break; }
          rc = maybe_recalc(s, true);
          if (rc != MLP_OK) break;
        }
        if (rc == MLP_OK && s->stage == 1) go = false;
        break;
      }
      case 2:
        if (!s->is_dual_feasible) {
          rc = mlp_recalc_obj_coeffs(s->eng, &s->cur_obj_val);  // 476
          if (rc == MLP_OK) {
            mlp_counters c;
            mlp_get_counters(s->eng, &c);
            s->eta_nnz = 0;  // recalc_obj_coeffs refactors whenever etas exist (1200-1203)
            s->lu_nnz = c.lu_nnz;
            s->stage = 3;
          }
        } else s->stage = 4;
        break;
      case 3: {
        while (budget_left()) {
          int moved = 0;
          rc = primal_iteration(s, &moved);
          if (rc != MLP_OK) break;
          if (!moved) { s->is_dual_feasible = true; s->stage = 4; break; }
          rc = maybe_recalc(s, false);
          if (rc != MLP_OK) break;
        }
        if (rc == MLP_OK && s->stage == 3) go = false;
        break;
      }
      default:
        s->enable_primal_steepest_edge = false;  // 482
        mlp_engine_set_primal_steepest_edge(s->eng, 0);
        *done = 1;
        go = false;
        break;
    }
  }
  s->run_seconds += std::chrono::duration<double>(Clock::now() - t0).count();
  return rc;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ incremental API (row f2)
// restore_feasibility (solver.rs:513-547) / optimize (487-511) run to completion, as the incremental entry points do.
static mlp_status restore_feasibility(mlp_solver* s) {
  for (;;) {
    int moved = 0;
    ST(dual_iteration(s, &moved));
    if (!moved) break;
  }
  s->is_primal_feasible = true;
  return MLP_OK;
}
static mlp_status optimize(mlp_solver* s) {
  for (;;) {
    int moved = 0;
    ST(primal_iteration(s, &moved));
    if (!moved) break;
  }
  s->is_dual_feasible = true;
  return MLP_OK;
}
static bool finished(const mlp_solver* s) { return s && s->initialized && s->stage == 4; }

// Solver::add_constraint (solver.rs:549-634); slack_coeffs (m doubles or null) only from add_gomory_cut.
static mlp_status add_constraint_impl(mlp_solver* s, const std::vector<double>& coeffs, const double* slack_coeffs, bool empty,
                                      int32_t cmp_op, double rhs) {
  if (!finished(s) || !s->is_primal_feasible || !s->is_dual_feasible) return MLP_INVALID;  // assert! 555-556
  if (empty) {  // 558-570
    const bool taut = cmp_op == 0 ? 0.0 == rhs : cmp_op == 1 ? 0.0 <= rhs : 0.0 >= rhs;
    return taut ? MLP_OK : MLP_INFEASIBLE;
  }
  double smin, smax;  // 573-577
  if (cmp_op == 1) { smin = 0.0; smax = kInf; }
  else if (cmp_op == 2) { smin = -kInf; smax = 0.0; }
  else { smin = 0.0; smax = 0.0; }
  mlp_add_row_result ar;
  ST(mlp_engine_add_row(s->eng, coeffs.data(), slack_coeffs, smin, smax, rhs, &ar));
  s->orig_obj_coeffs.push_back(0.0);  // 579-585
  s->orig_var_mins.push_back(smin);
  s->orig_var_maxs.push_back(smax);
  s->basic_vars.push_back(ar.slack_var);
  s->m += 1;
  s->eta_nnz = 0;
  s->lu_nnz = ar.lu_nnz;
  s->is_primal_feasible = false;  // 632-633
  return restore_feasibility(s);
}

extern "C" {
mlp_status mlp_solver_add_constraint(mlp_solver* s, int64_t count, const int64_t* vars, const double* coeffs, int32_t cmp_op,
                                     double rhs) {
  if (!s || count < 0 || (count > 0 && (!vars || !coeffs))) return MLP_INVALID;
  std::vector<double> row((size_t)s->n, 0.0);
  for (int64_t t = 0; t < count; ++t) {
    if (vars[t] < 0 || vars[t] >= s->n) return MLP_INVALID;
    row[(size_t)vars[t]] = coeffs[t];  // CsVec::new rejects duplicates (lib.rs:376); the Python mirror checks before calling
  }
  return add_constraint_impl(s, row, nullptr, count == 0, cmp_op, rhs);
}

// Solution: Clone (lib.rs:313-318)
mlp_status mlp_solver_clone(mlp_solver* s, mlp_solver** out) {
  if (!s || !out) return MLP_INVALID;
  *out = nullptr;
  mlp_engine* e = nullptr;
  ST(mlp_engine_clone(s->eng, &e));
  mlp_solver* c = new mlp_solver(*s);
  c->eng = e;
  *out = c;
  return MLP_OK;
}

// Solver::fix_var, solver.rs:378-415
mlp_status mlp_solver_fix_var(mlp_solver* s, int64_t var, double val) {
  if (!finished(s) || var < 0 || var >= s->n) return MLP_INVALID;
  if (val < s->orig_var_mins[var] || val > s->orig_var_maxs[var]) return MLP_INFEASIBLE;  // 379-381
  mlp_var_info vi;
  ST(mlp_get_var(s->eng, var, &vi));
  if (vi.flags & MLP_BASIC) {  // 384-392: pivot it out of the basis at the value `val`
    const int64_t row = vi.pos_or_row;
    ST(mlp_calc_row_coeffs(s->eng, row));
    mlp_dual_entering de;
    ST(mlp_ratio_dual(s->eng, row, val, &de));
    if (de.var < 0) return MLP_INFEASIBLE;  // 1019
    note_ties(s, de.ties, de.near_ties);
    const double entering_diff = (vi.value - val) / de.coeff;  // 1005
    const double entering_new_val = de.cur_val + entering_diff;
    ST(mlp_ftran_col(s->eng, de.var));
    ST(do_pivot(s, 0, de.var, de.pos, de.obj_coeff, entering_new_val, entering_diff, true, row, de.coeff, val));
  } else {  // 394-404: move the non-basic variable, no basis change (not a pivot: no trace record)
    const int64_t col = vi.pos_or_row;
    ST(mlp_ftran_col(s->eng, var));
    const double diff = val - vi.value;
    mlp_pivot_info pi;
    std::memset(&pi, 0, sizeof(pi));
    pi.entering_var = var;
    pi.col = col;
    pi.entering_new_val = val;
    pi.entering_diff = diff;
    pi.has_elem = 0;
    mlp_pivot_result pr;
    ST(mlp_pivot(s->eng, &pi, &pr));
    s->cur_obj_val += diff * vi.obj_coeff;  // 401
    s->nb_var_vals[col] = val;
  }
  ST(mlp_set_nb_state(s->eng, var, MLP_AT_MIN | MLP_AT_MAX | MLP_FIXED));  // 407-411
  s->is_primal_feasible = false;
  return restore_feasibility(s);
}

// Solver::unfix_var, solver.rs:418-438
mlp_status mlp_solver_unfix_var(mlp_solver* s, int64_t var, int32_t* was_fixed) {
  if (was_fixed) *was_fixed = 0;
  if (!finished(s) || var < 0 || var >= s->n) return MLP_INVALID;
  mlp_var_info vi;
  ST(mlp_get_var(s->eng, var, &vi));
  if ((vi.flags & MLP_BASIC) || !(vi.flags & MLP_FIXED)) return MLP_OK;
  const uint32_t f = (vi.value == s->orig_var_mins[var] ? MLP_AT_MIN : 0u) | (vi.value == s->orig_var_maxs[var] ? MLP_AT_MAX : 0u);
  ST(mlp_set_nb_state(s->eng, var, f));  // 424-428
  s->is_dual_feasible = false;
  ST(optimize(s));  // 432: .unwrap() in the reference — an error status here is that panic
  if (was_fixed) *was_fixed = 1;
  return MLP_OK;
}

// Solver::add_gomory_cut, solver.rs:440-460
mlp_status mlp_solver_add_gomory_cut(mlp_solver* s, int64_t var) {
  if (!finished(s) || var < 0 || var >= s->n) return MLP_INVALID;
  mlp_var_info vi;
  ST(mlp_get_var(s->eng, var, &vi));
  if (!(vi.flags & MLP_BASIC)) return MLP_INVALID;  // panic!("var is not basic"), 458
  const int64_t row = vi.pos_or_row, nt = s->n + s->m;
  ST(mlp_calc_row_coeffs(s->eng, row));
  std::vector<double> rc((size_t)nt), flags_dummy;
  ST(mlp_download_f64(s->eng, MLP_ARR_ROW_COEFFS, rc.data(), nt));
  std::vector<uint8_t> fl((size_t)nt);
  std::vector<int32_t> pos((size_t)nt);
  ST(mlp_download_var_state(s->eng, fl.data(), pos.data()));
  std::vector<double> cut((size_t)s->n, 0.0), cut_slack((size_t)s->m, 0.0);
  bool any_slack = false, any = false;
  for (int64_t v = 0; v < nt; ++v) {  // 444-448 over the non-zeros of row_coeffs
    if ((fl[v] & MLP_BASIC) || rc[v] == 0.0) continue;
    const double c = std::floor(rc[v]) - rc[v];
    any = true;
    if (v < s->n) cut[(size_t)v] = c;
    else { cut_slack[(size_t)(v - s->n)] = c; any_slack = any_slack || c != 0.0; }
  }
  const double cut_bound = std::floor(vi.value) - vi.value;  // 450
  return add_constraint_impl(s, cut, any_slack ? cut_slack.data() : nullptr, !any, 1 /* Le */, cut_bound);
}
}  // extern "C"

extern "C" {

double mlp_solver_cur_obj_val(mlp_solver* s) { return s->cur_obj_val; }
int64_t mlp_solver_pivots_done(mlp_solver* s) { return s->pivots_done; }
int64_t mlp_solver_num_vars(mlp_solver* s) { return s->n; }
int64_t mlp_solver_num_constraints(mlp_solver* s) { return s->m; }

mlp_status mlp_solver_values(mlp_solver* s, double* out) {  // Solver::get_value, 371-376
  std::vector<double> xb(s->m);
  ST(mlp_download_f64(s->eng, MLP_ARR_BASIC_VALS, xb.data(), s->m));
  std::vector<int64_t> where(s->n, -1);
  for (int64_t r = 0; r < s->m; ++r)
    if (s->basic_vars[r] < s->n) out[s->basic_vars[r]] = xb[r];
  for (int64_t c = 0; c < s->n; ++c)
    if (s->nb_vars[c] < s->n) out[s->nb_vars[c]] = s->nb_var_vals[c];
  return MLP_OK;
}

int64_t mlp_solver_trace_len(mlp_solver* s) { return (int64_t)s->trace.size(); }
int64_t mlp_solver_get_trace(mlp_solver* s, int64_t first, int64_t count, double* out) {
  int64_t k = 0;
  if (!s || !out || first < 0 || count < 0) return 0;
  for (int64_t i = first; i < first + count && i < (int64_t)s->trace.size(); ++i, ++k) {
    const PivotRecord& r = s->trace[(size_t)i];
    double* o = out + k * 13;
    o[0] = r.phase; o[1] = (double)r.entering_var; o[2] = (double)r.entering_col; o[3] = (double)r.leaving_row;
    o[4] = (double)r.leaving_var; o[5] = r.pivot_coeff; o[6] = r.entering_diff; o[7] = r.obj_after;
    o[8] = (double)r.eta_count; o[9] = (double)r.lu_nnz; o[10] = (double)r.nnz_col; o[11] = (double)r.nnz_rho; o[12] = r.refactored;
  }
  return k;
}
void mlp_solver_set_record_trace(mlp_solver* s, int32_t on) { s->record_trace = on != 0; }
mlp_status mlp_solver_get_nb_vars(mlp_solver* s, int64_t* out) {
  std::memcpy(out, s->nb_vars.data(), s->n * sizeof(int64_t));
  return MLP_OK;
}
mlp_status mlp_solver_get_basic_vars(mlp_solver* s, int64_t* out) {
  std::memcpy(out, s->basic_vars.data(), s->m * sizeof(int64_t));
  return MLP_OK;
}
void mlp_solver_set_refactor_factor(mlp_solver* s, double factor) { s->refactor_factor = factor > 0.0 ? factor : 1.0; }
void mlp_solver_set_recalc_period(mlp_solver* s, int64_t period) { s->recalc_period = period > 0 ? period : 0; }
int64_t mlp_solver_recalcs_done(mlp_solver* s) { return s->recalcs_done; }
void mlp_solver_tie_stats(mlp_solver* s, int64_t out4[4]) {
  out4[0] = s->tied_pivots;
  out4[1] = s->near_tie_pivots;
  out4[2] = s->first_tied_pivot;
  out4[3] = s->first_near_tie_pivot;
}
void mlp_solver_timers(mlp_solver* s, double* run_seconds, double* refactor_seconds) {
  *run_seconds = s->run_seconds;
  *refactor_seconds = s->refactor_seconds;
}

}  // extern "C"
