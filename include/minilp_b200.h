/*
 * minilp_b200 — C ABI of the B200-native revised-simplex pivot engine.
 *
 * This is the drop-in boundary for the hot path of ztlpn/minilp (pure Rust, commit b99146b).
 * The reference has no FFI of its own: everything below `Problem`/`Solution` is private
 * (lib.rs:52-59).  The seam is `BasisSolver` (solver.rs:1265-1339) widened to the O(m)/O(n)
 * loops of `Solver` that consume its results, so that no m- or n-vector crosses the bus per
 * pivot.  Each entry point names the reference item it replaces (file:line under
 * /root/reference/src).  INTEGRATION.md shows the Rust `extern "C"` block a maintainer would add.
 *
 * Conventions
 *  - plain C: opaque handles, pointers and sizes; no C++/torch types.
 *  - every function returns an mlp_status; outputs go through out-pointers.
 *  - one host thread per handle; calls are synchronous with respect to their outputs.
 *  - "var" is a variable index in [0, n+m): structural j < n, slack of row i is n+i
 *    (solver.rs:230-232).  "pos" is a non-basic position (index into nb_vars, solver.rs:44),
 *    "row" a basis position / constraint row (index into basic_vars, solver.rs:37).
 *  - the engine never falls back to the CPU: without a CUDA device every compute entry point
 *    returns MLP_NO_DEVICE.
 */
#ifndef MINILP_B200_H
#define MINILP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mlp_status {
  MLP_OK = 0,
  MLP_INFEASIBLE = 1, /* Error::Infeasible, lib.rs:175 */
  MLP_UNBOUNDED = 2,  /* Error::Unbounded,  lib.rs:177 */
  MLP_SINGULAR = 3,   /* Error::SingularMatrix (sparse.rs:335); the reference unwrap()s it: solver.rs:316,1301 */
  MLP_NONFINITE = 4,  /* assert!(is_finite) solver.rs:1149,1172 */
  MLP_INVALID = 5,    /* bad argument / call order */
  MLP_CUDA_ERROR = 6,
  MLP_NO_DEVICE = 7,
  MLP_NOMEM = 8
} mlp_status;

const char* mlp_last_error(void);
const char* mlp_version(void);
/* number of visible CUDA devices (0 when there is none); never fails */
int mlp_device_count(void);

/* ===================================================================== engine (device state) */
typedef struct mlp_engine mlp_engine;

/* non-basic variable state bits: NonBasicVarState + nb_var_is_fixed (solver.rs:47-48, 66-70) */
#define MLP_AT_MIN 1u
#define MLP_AT_MAX 2u
#define MLP_BASIC 4u
#define MLP_FIXED 8u

/* Allocate the device-resident state for an m x n dense constraint matrix A (row-major f64 in HBM).
 * Replaces the storage half of Solver (solver.rs:15-58: orig_constraints / orig_constraints_csc). */
mlp_status mlp_engine_create_dense(int device, int64_t m, int64_t n, mlp_engine** out);
/* Sparse storage (BASELINE config 4): A as CSR — row_ptr (m+1), col_idx / vals (nnz), columns ascending within a row, the
 * order CsVec::new gives every constraint (lib.rs:279) — copied to the device together with the CSC copy the engine
 * derives (CsMat::to_csc, solver.rs:253).  Replaces orig_constraints / orig_constraints_csc (solver.rs:21-22) with 32-bit
 * indices.  Every other entry point is storage-agnostic; mlp_engine_upload_rows is rejected. */
mlp_status mlp_engine_create_sparse(int device, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
                                    const int32_t* col_idx, const double* vals, mlp_engine** out);
/* The CSC copy the engine derived on the device (SparseMat::transpose, sparse.rs:230-269: rows ascending within a column);
 * col_ptr: n+1, row_idx / vals: nnz; NULL pointers are skipped. */
mlp_status mlp_engine_download_csc(mlp_engine* e, int64_t* col_ptr, int32_t* row_idx, double* vals);
/* Column-sharded engine: rank `rank` of `world` owns the structural columns
 * mlp_shard_range(n_global, world, rank) of A and all per-variable arrays of those columns; m-sized state and the
 * basis factors are replicated and every rank runs the identical host control loop (SPMD).  All variable indices
 * in this ABI stay GLOBAL.  The one exchange step per pivot — arg-reduce of the per-shard pricing candidates plus
 * the winner's column — is a single all-gather inside mlp_select_entering_primal / mlp_ratio_dual.
 *   comm_kind MLP_COMM_NCCL : comm_arg = 128-byte ncclUniqueId from mlp_nccl_get_unique_id (one process per GPU;
 *                             an id serves one engine: fetch a fresh one per create call)
 *   comm_kind MLP_COMM_LOCAL: comm_arg = handle from mlp_local_group_create (one host thread per shard, one process) */
#define MLP_COMM_NONE 0
#define MLP_COMM_NCCL 1
#define MLP_COMM_LOCAL 2
mlp_status mlp_engine_create_dense_sharded(int device, int64_t m, int64_t n_global, int32_t rank, int32_t world,
                                           int32_t comm_kind, const void* comm_arg, mlp_engine** out);
/* Column-sharded SPARSE engine (BASELINE north_star: Netlib-shaped LPs at 1/2/4/8 GPUs).  Every rank passes the WHOLE CSR
 * matrix and keeps it (12 nnz bytes — small next to HBM; FTRAN / BTRAN / refactorization need the basic columns wherever
 * they are priced, and with it no column ever crosses NVLink); what is sharded is the work that scales with the matrix and
 * the variables: the price-out runs over the rank's own column block of the CSC copy only, and the per-variable arrays
 * (reduced costs, steepest-edge norms, states) hold that block plus the slacks.  Exchange step as in the dense case. */
mlp_status mlp_engine_create_sparse_sharded(int device, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
                                            const int32_t* col_idx, const double* vals, int32_t rank, int32_t world,
                                            int32_t comm_kind, const void* comm_arg, mlp_engine** out);
mlp_status mlp_nccl_get_unique_id(void* out128);
mlp_status mlp_local_group_create(int32_t world, void** out);
void mlp_local_group_destroy(void* group);
mlp_status mlp_engine_local_range(mlp_engine* e, int64_t* col_begin, int64_t* col_end);
/* How the per-pivot candidate exchange runs: 0 single shard, 1 NCCL all-gather, 2 in-process group, 3 one kernel over
 * NVLink peer memory (CUDA IPC; chosen automatically with MLP_COMM_NCCL when every rank can map every peer, MLP_P2P=0
 * disables it). */
int32_t mlp_engine_exchange_kind(mlp_engine* e);
void mlp_engine_destroy(mlp_engine* e);
/* Stream `nrows` consecutive rows of A from HOST memory, starting at row0: full rows of n_global doubles (a sharded
 * engine takes its own column slice) ... */
mlp_status mlp_engine_upload_rows(mlp_engine* e, int64_t row0, int64_t nrows, const double* rows_host);
/* ... or rows that hold only this shard's columns (col_end - col_begin doubles each). */
mlp_status mlp_engine_upload_local_rows(mlp_engine* e, int64_t row0, int64_t nrows, const double* rows_local);

/* State at the end of Solver::try_new (solver.rs:108-369).  Arrays are HOST pointers.
 * Position-indexed arrays (nb_*) have n entries, row-indexed ones m, var-indexed ones n+m. */
typedef struct mlp_init_state {
  const double* orig_var_mins;        /* n+m  solver.rs:19,224 */
  const double* orig_var_maxs;        /* n+m  solver.rs:20,225 */
  const double* orig_obj_coeffs;      /* n+m  solver.rs:18,244-245 (internal sign: Maximize already negated) */
  const double* orig_rhs;             /* m    solver.rs:23 */
  const int64_t* nb_vars;             /* n    solver.rs:44 (n = n_global on a sharded engine: every rank gets the full arrays) */
  const double* nb_var_vals;          /* n    solver.rs:46 */
  const double* nb_var_obj_coeffs;    /* n    solver.rs:45,279-295 */
  const uint8_t* nb_var_states;       /* n    MLP_AT_MIN | MLP_AT_MAX | MLP_FIXED */
  const double* primal_edge_sq_norms; /* n or NULL: computed on device as |a_j|^2 + 1 (solver.rs:297-299) */
  const int64_t* basic_vars;          /* m    solver.rs:37 */
  const double* basic_var_vals;       /* m or NULL: computed on device as rhs - A x_N (solver.rs:234-238) */
  const double* basic_var_mins;       /* m    solver.rs:39 */
  const double* basic_var_maxs;       /* m    solver.rs:40 */
  const double* dual_edge_sq_norms;   /* m or NULL: all 1.0 (solver.rs:264-268) */
  int32_t enable_primal_steepest_edge; /* solver.rs:25,272 */
  int32_t enable_dual_steepest_edge;   /* solver.rs:26,263 */
} mlp_init_state;
/* Uploads the state and factorizes the initial basis (lu_factorize at solver.rs:305-317). */
mlp_status mlp_engine_init_state(mlp_engine* e, const mlp_init_state* st);
/* solver.rs:482 clears the flag after the initial solve */
mlp_status mlp_engine_set_primal_steepest_edge(mlp_engine* e, int32_t enable);

/* BasisSolver::reset (solver.rs:1286-1303): refactorize the current basis, drop the eta file.
 * Dense engine: the LU of the basis is computed on the device.  lu_nnz receives
 * the value LUFactors::nnz (lu.rs:52-54) would report for the refactor rule at solver.rs:1096-1097. */
mlp_status mlp_refactor(mlp_engine* e, int64_t* lu_nnz);

/* choose_pivot's pricing scan (solver.rs:696-739). var = -1: no eligible column (optimal). */
typedef struct mlp_entering {
  int64_t var, pos;
  double obj_coeff; /* nb_var_obj_coeffs[pos] */
  double score;
  double cur_val;   /* nb_var_vals[pos] */
} mlp_entering;
mlp_status mlp_select_entering_primal(mlp_engine* e, mlp_entering* out);

/* calc_col_coeffs (solver.rs:671-677) = BasisSolver::solve (FTRAN, 1305-1319) of the column of `var`.
 * Result stays on the device as col_coeffs. */
mlp_status mlp_ftran_col(mlp_engine* e, int64_t var);

/* The two-pass Harris ratio test of choose_pivot (solver.rs:752-823) over col_coeffs.
 * entering_diff_sign: solver.rs:743; max_step0 = |entering_other_val - entering_cur_val| (782).
 * row = -1: no blocking row (bound flip or unbounded, 841-852). */
typedef struct mlp_leaving {
  int64_t row;
  double coeff;           /* pivot_coeff */
  double leaving_new_val; /* 813-819 */
  double basic_val;       /* basic_var_vals[row] (828) */
  /* How contested the pass-2 winner was (804-823 keep the FIRST maximal |coeff| in col_coeffs list order, sparse.rs:75-80;
   * the engine keeps the lowest row): other eligible rows whose |coeff| equals the winner's exactly / within 1e-9 relative.
   * ties == 0 means the choice does not depend on the scan order. */
  int64_t ties, near_ties;
} mlp_leaving;
mlp_status mlp_ratio_primal(mlp_engine* e, int32_t entering_diff_sign, double max_step0, mlp_leaving* out);

/* calc_row_coeffs (solver.rs:680-693): BTRAN of e_row (BasisSolver::solve_transp, 1322-1338) followed
 * by the price-out of the tableau row.  Results stay on the device (inv_basis_row_coeffs, row_coeffs). */
mlp_status mlp_btran_unit(mlp_engine* e, int64_t row);
mlp_status mlp_price_row(mlp_engine* e);
mlp_status mlp_calc_row_coeffs(mlp_engine* e, int64_t row);

/* choose_pivot_row_dual (solver.rs:855-917). row = -1: primal feasible. */
typedef struct mlp_dual_row {
  int64_t row;
  double val, min, max;
} mlp_dual_row;
mlp_status mlp_select_row_dual(mlp_engine* e, mlp_dual_row* out);

/* choose_entering_col_dual (solver.rs:919-1021) over row_coeffs. var = -1: Err(Infeasible) (1019). */
typedef struct mlp_dual_entering {
  int64_t var, pos;
  double coeff;     /* pivot_coeff */
  double obj_coeff; /* nb_var_obj_coeffs[pos] */
  double cur_val;   /* nb_var_vals[pos] */
  int64_t ties, near_ties; /* as in mlp_leaving, for pass 2 at 982-1002 (the engine keeps the lowest variable index) */
} mlp_dual_entering;
mlp_status mlp_ratio_dual(mlp_engine* e, int64_t row, double leaving_new_val, mlp_dual_entering* out);
/* The first half of one iteration of restore_feasibility (solver.rs:529-531) with ONE host round trip:
 * choose_pivot_row_dual -> calc_row_coeffs(row) -> choose_entering_col_dual queued back to back, the chosen row handed from
 * kernel to kernel in device memory; *row as mlp_select_row_dual, *out as mlp_ratio_dual (with leaving_new_val = the violated
 * bound, 908-915).  row->row < 0: no infeasible row, *out is meaningless.  Same kernels and results as the three calls. */
mlp_status mlp_dual_select_ratio(mlp_engine* e, mlp_dual_row* row, mlp_dual_entering* out);

/* PivotInfo / PivotElem (solver.rs:1245-1261) plus the refactor decision of solver.rs:1096-1103,
 * which the host takes from the running eta nnz and lu_nnz. */
typedef struct mlp_pivot_info {
  int64_t entering_var, col;
  double entering_obj_coeff; /* nb_var_obj_coeffs[col] as returned by the selection (the column may live on another shard) */
  double entering_new_val, entering_diff;
  int32_t has_elem;
  int64_t row;
  double coeff, leaving_new_val;
  int32_t refactor; /* 1: BasisSolver::reset instead of push_eta_matrix */
} mlp_pivot_info;
typedef struct mlp_pivot_result {
  int64_t leaving_var;   /* -1 on a bound flip */
  int64_t col_nnz;       /* structural size of the pushed eta column (nnz of col_coeffs) */
  int64_t eta_count;     /* eta_matrices.len() after the call */
  int64_t lu_nnz;        /* valid after a refactor, else unchanged value */
  int32_t refactored;    /* the engine may also force a refactor when its eta arena is full */
} mlp_pivot_result;
/* Solver::pivot (solver.rs:1023-1104) incl. update_dual_sq_norms (1153-1174), update_primal_sq_norms
 * (1106-1151) and push_eta_matrix (1274-1284).  Returns MLP_NONFINITE where the reference asserts. */
mlp_status mlp_pivot(mlp_engine* e, const mlp_pivot_info* info, mlp_pivot_result* out);

/* recalc_obj_coeffs (solver.rs:1199-1231): y = B^-T c_B, d_N = c_N - N^T y, objective from scratch. */
mlp_status mlp_recalc_obj_coeffs(mlp_engine* e, double* cur_obj_val);

/* recalc_basic_var_vals (solver.rs:1177-1197; dead code in the reference, asked for by the TODO at 1024-1025): x_B =
 * B^-1 (rhs - N x_N) from scratch, refactorizing first when etas exist.  SURVEY.md §8 row f4: never called unless the
 * host asks (mlp_solver_set_recalc_period), so the default path is the reference's. */
mlp_status mlp_recalc_basic_vals(mlp_engine* e);

/* ---- incremental API: the engine half of Solver::fix_var (solver.rs:378-415), unfix_var (418-438),
 * add_constraint (549-634) and add_gomory_cut (440-460).  Single-shard engines, dense or sparse storage. */
typedef struct mlp_var_info {
  uint32_t flags;      /* MLP_BASIC or MLP_AT_MIN | MLP_AT_MAX | MLP_FIXED */
  int64_t pos_or_row;  /* VarState::Basic(row) / NonBasic(col), solver.rs:60-64 */
  double obj_coeff;    /* nb_var_obj_coeffs[col] (0 for a basic variable) */
  double value;        /* Solver::get_value, 371-376 */
} mlp_var_info;
mlp_status mlp_get_var(mlp_engine* e, int64_t var, mlp_var_info* out);
/* nb_var_states[col] / nb_var_is_fixed[col] of a non-basic variable (405-411, 420-428) */
mlp_status mlp_set_nb_state(mlp_engine* e, int64_t var, uint32_t flags);
typedef struct mlp_add_row_result {
  int64_t row, slack_var;
  double basic_val;  /* basic_var_vals.push(rhs - lhs_val), 591 */
  double rhs;        /* stored right-hand side (differs from the argument when slack coefficients were substituted) */
  int64_t lu_nnz;    /* after basis_solver.reset, 612 */
} mlp_add_row_result;
/* Appends the row  coeffs . x + slack_coeffs . s + s_new = rhs  with s_new in [slack_min, slack_max] (563-571), makes
 * s_new basic, refactorizes (612) and extends the steepest-edge norms by the new tableau row (615-630).
 * coeffs: n doubles (structural variables); slack_coeffs: m doubles or NULL — coefficients on EXISTING slack variables, as
 * a Gomory cut has; they are eliminated through s_i = rhs_i - a_i x, so slack columns stay unit columns. */
mlp_status mlp_engine_add_row(mlp_engine* e, const double* coeffs, const double* slack_coeffs, double slack_min,
                              double slack_max, double rhs, mlp_add_row_result* out);

/* Solver: Clone (solver.rs:14; Solution: Clone, lib.rs:313): device-to-device deep copy, factors and eta file included. */
mlp_status mlp_engine_clone(mlp_engine* src, mlp_engine** out);

/* Downloads (device -> caller buffer).  Var-indexed arrays have n+m entries, row-indexed m. */
typedef enum mlp_array {
  MLP_ARR_OBJ_COEFFS = 0,   /* d, by var (valid where non-basic) */
  MLP_ARR_PRIMAL_NORMS = 1, /* by var */
  MLP_ARR_NB_VALS = 2,      /* by var */
  MLP_ARR_BASIC_VALS = 3,   /* by row */
  MLP_ARR_DUAL_NORMS = 4,   /* by row */
  MLP_ARR_COL_COEFFS = 5,   /* by row: last FTRAN result */
  MLP_ARR_INV_BASIS_ROW = 6,/* by row: last BTRAN result */
  MLP_ARR_ROW_COEFFS = 7,   /* by var: last price-out */
  MLP_ARR_BASIC_MINS = 8,
  MLP_ARR_BASIC_MAXS = 9,
  MLP_ARR_SE_HELPER = 10    /* by var: N^T v of the last primal steepest-edge update */
} mlp_array;
mlp_status mlp_download_f64(mlp_engine* e, int32_t which, double* out, int64_t count);
mlp_status mlp_download_basic_vars(mlp_engine* e, int64_t* out /* m */);
mlp_status mlp_download_var_state(mlp_engine* e, uint8_t* flags /* n+m */, int32_t* pos_or_row /* n+m */);

/* Instrumentation. */
typedef struct mlp_counters {
  int64_t kernel_launches;
  int64_t h2d_bytes, d2h_bytes;
  int64_t refactors, etas_pushed;
  int64_t k_structural; /* structural columns in the last factorized basis */
  int64_t lu_nnz;       /* LUFactors::nnz of the current factors (lu.rs:52-54) */
  int64_t eta_count;    /* eta_matrices.len() */
  int64_t ratio_ties;      /* ratio tests (primal or dual) whose pass-2 winner was tied exactly */
  int64_t ratio_near_ties; /* ... or within 1e-9 relative (includes the exact ones) */
  int64_t refreshes;       /* of `refactors`: those done as a product-form refresh of the core inverse (MLP_TUNE_LU_EVERY) */
  int64_t refresh_rejects; /* refreshes whose accuracy probe (max |C C^-1 - I| over sampled columns) exceeded the tolerance: the
                              next refactorization was a true one */
} mlp_counters;
mlp_status mlp_get_counters(mlp_engine* e, mlp_counters* out);
/* cudaStream_t of the engine (as void*), for CUDA-event timing by the caller. */
void* mlp_engine_stream(mlp_engine* e);
mlp_status mlp_engine_sync(mlp_engine* e);

/* CUDA-event marks on the engine's stream: mark(slot) records event `slot` (0..3); elapsed gives milliseconds
 * between two recorded marks (synchronizes on the later one). */
mlp_status mlp_event_mark(mlp_engine* e, int32_t slot);
mlp_status mlp_event_elapsed_ms(mlp_engine* e, int32_t slot_a, int32_t slot_b, double* ms);

/* Live per-launch timing of the price-out kernel pair (k_price_partial + k_price_finish) with CUDA events on
 * the engine's stream, accumulated over the pivots made while enabled.  Slot "v" is the N^T v product of
 * update_primal_sq_norms (solver.rs:1117-1132; v is dense), slot "rho" the tableau-row price-out (685-692).
 * bytes are the ALGORITHMIC bytes 8 n s + 8 s + 8 n with s the actual support size of each launch. */
typedef struct mlp_profile {
  double price_v_ms, price_rho_ms;
  int64_t price_v_launches, price_rho_launches;
  int64_t price_v_bytes, price_rho_bytes;
} mlp_profile;
mlp_status mlp_profile_enable(mlp_engine* e, int32_t on);
mlp_status mlp_profile_get(mlp_engine* e, mlp_profile* out);

/* Kernel-isolated bench hooks (bench.py roofline): run the price-out kernel `iters` times with a dense
 * multiplier vector over all m rows; returns the mean device time of one launch pair in milliseconds. */
mlp_status mlp_bench_price_dense(mlp_engine* e, int32_t iters, double* ms_per_launch, int64_t* bytes_per_launch);

/* Tuning knobs of the device path (experiments and A/B measurements; none of them changes a result beyond the rounding
 * of re-ordered reductions).  Takes effect from the next call on.  The same knobs are read from the environment when an
 * engine is created (MLP_PRICE_TILE, MLP_PRICE_SPLIT, MLP_LANE1_LDG, MLP_FUSED, MLP_FUSED_MAX). */
enum {
  MLP_TUNE_PRICE_TILE = 0, /* columns per tile of the bulk-copy price-out: a multiple of 64 in [128, 4096]; 0 = automatic.
                              Resets the split to 1. */
  MLP_TUNE_LANE1_LDG = 1,  /* 1: the tableau-row price-out runs as the LDG kernel beside lane 0's bulk-copy kernel */
  MLP_TUNE_FUSED = 2,      /* 1: FTRAN -> BTRAN chain of a primal pivot as one cooperative kernel */
  MLP_TUNE_FUSED_MAX = 3,  /* largest k / K that takes the fused chain (<= 512) */
  MLP_TUNE_PRICE_SPLIT = 4, /* column slices (1, 2, 4) per work item of the price-out's last, partial round */
  MLP_TUNE_LU_EVERY = 5     /* sparse storage: pivots between two TRUE factorizations (lu.rs:118-304).  The refactorizations the
                               refactor rule asks for in between (solver.rs:1096-1103) fold the eta file into the explicit inverse
                               of the core instead (product form, O(k^2 K); csrc/refresh_inverse.cuh).  0: every refactorization
                               is a true factorization (BasisSolver::reset as the reference does it). */
};
mlp_status mlp_engine_set_tuning(mlp_engine* e, int32_t knob, int32_t value);
mlp_status mlp_engine_get_tuning(mlp_engine* e, int32_t knob, int32_t* value);

/* ===================================================================== host control loop */
/* C++ mirror of the reference's Solver control flow (try_new 108-369, initial_solve 470-485,
 * optimize 487-511, restore_feasibility 513-547, choose_pivot 695-853, pivot's host half) written
 * ONLY against the engine ABI above — it is what the Rust `Solver` would look like after the swap. */
typedef struct mlp_solver mlp_solver;

mlp_status mlp_solver_create_dense(int device, int64_t m, int64_t n, mlp_solver** out);
mlp_status mlp_solver_create_dense_sharded(int device, int64_t m, int64_t n_global, int32_t rank, int32_t world,
                                           int32_t comm_kind, const void* comm_arg, mlp_solver** out);
mlp_status mlp_solver_create_sparse(int device, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
                                    const int32_t* col_idx, const double* vals, mlp_solver** out);
mlp_status mlp_solver_create_sparse_sharded(int device, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
                                            const int32_t* col_idx, const double* vals, int32_t rank, int32_t world,
                                            int32_t comm_kind, const void* comm_arg, mlp_solver** out);
mlp_status mlp_solver_upload_local_rows(mlp_solver* s, int64_t row0, int64_t nrows, const double* rows_local);
void mlp_solver_destroy(mlp_solver* s);
mlp_engine* mlp_solver_engine(mlp_solver* s);
mlp_status mlp_solver_upload_rows(mlp_solver* s, int64_t row0, int64_t nrows, const double* rows_host);
/* Solver::try_new.  obj_coeffs are internal-sign (lib.rs:235-238 already applied); cmp_ops: 0 Eq, 1 Le, 2 Ge. */
mlp_status mlp_solver_init(mlp_solver* s, const double* obj_coeffs, const double* var_mins, const double* var_maxs,
                           const int32_t* cmp_ops, const double* rhs);
/* Solver::initial_solve with a pivot budget (max_pivots < 0: to completion). *done = 1 when finished. */
mlp_status mlp_solver_run(mlp_solver* s, int64_t max_pivots, int32_t* done);
double mlp_solver_cur_obj_val(mlp_solver* s);
int64_t mlp_solver_pivots_done(mlp_solver* s);
int64_t mlp_solver_num_vars(mlp_solver* s);
int64_t mlp_solver_num_constraints(mlp_solver* s);
/* Solver::get_value for all structural variables (solver.rs:371-376). out: n doubles. */
mlp_status mlp_solver_values(mlp_solver* s, double* out);
/* Per-pivot trace, 13 doubles per record, same layout as the oracle's. */
int64_t mlp_solver_trace_len(mlp_solver* s);
int64_t mlp_solver_get_trace(mlp_solver* s, int64_t first, int64_t count, double* out);
void mlp_solver_set_record_trace(mlp_solver* s, int32_t on);
/* Solution::add_constraint / fix_var / unfix_var / add_gomory_cut (lib.rs:368-423) on a solved problem.
 * add_constraint: `count` (variable, coefficient) pairs over structural variables, cmp_op 0 Eq / 1 Le / 2 Ge.
 * Status MLP_INFEASIBLE where the reference returns Err(Infeasible); *was_fixed mirrors unfix_var's bool. */
mlp_status mlp_solver_add_constraint(mlp_solver* s, int64_t count, const int64_t* vars, const double* coeffs, int32_t cmp_op,
                                     double rhs);
mlp_status mlp_solver_fix_var(mlp_solver* s, int64_t var, double val);
mlp_status mlp_solver_unfix_var(mlp_solver* s, int64_t var, int32_t* was_fixed);
mlp_status mlp_solver_add_gomory_cut(mlp_solver* s, int64_t var);
mlp_status mlp_solver_clone(mlp_solver* s, mlp_solver** out);
/* host mirrors of nb_vars (n) and basic_vars (m) */
mlp_status mlp_solver_get_nb_vars(mlp_solver* s, int64_t* out);
mlp_status mlp_solver_get_basic_vars(mlp_solver* s, int64_t* out);
/* The refactor rule of solver.rs:1096-1097 is `eta nnz >= lu nnz`; here `eta nnz >= factor * lu nnz`.  factor = 1 (default):
 * the reference's rule.  factor > 1 keeps the eta file longer (fewer refactorizations); results change in rounding only. */
void mlp_solver_set_refactor_factor(mlp_solver* s, double factor);
/* Row f4: every `period` pivots (0, the default = never = the reference's behaviour) recompute x_B and — unless the
 * artificial objective of solver.rs:261 is in place — d and the objective from scratch. */
void mlp_solver_set_recalc_period(mlp_solver* s, int64_t period);
int64_t mlp_solver_recalcs_done(mlp_solver* s);
/* Pivots whose ratio-test winner was contested (see mlp_leaving.ties): out4 = { pivots tied exactly, pivots tied within
 * 1e-9, index of the first exactly tied pivot or -1, index of the first near-tied pivot or -1 }.  All zero / -1 means the
 * pivot sequence does not depend on the reference's list-order rule (sparse.rs:75-80 with solver.rs:811, 996). */
void mlp_solver_tie_stats(mlp_solver* s, int64_t out4[4]);
/* seconds of wall clock spent inside mlp_solver_run so far, and of that inside refactorizations */
void mlp_solver_timers(mlp_solver* s, double* run_seconds, double* refactor_seconds);

/* ===================================================================== MPS ingest (host, no GPU) */
/* MpsFile::parse (mps.rs:39-329) over one in-memory buffer of free-format MPS text: same sections, same first-vector
 * rules, same bound defaults, ranged rows doubled (306-321).  Syntax errors: MLP_INVALID, mlp_last_error() = the reference's
 * "line N: ..." message.  The result is held as flat arrays — variables (objective in the user's sign, bounds as
 * Problem::add_var gets them, mps.rs:294-303) and every constraint as a CSR row with ascending variable indices
 * (CsVec::new, lib.rs:279) — ready for Solver::try_new / mlp_solver_create_sparse after the empty-row filter
 * (solver.rs:201-213). */
typedef struct mlp_mps mlp_mps;
mlp_status mlp_mps_parse(const char* text, int64_t len, mlp_mps** out);
void mlp_mps_free(mlp_mps* f);
const char* mlp_mps_name(mlp_mps* f); /* MpsFile::problem_name */
void mlp_mps_sizes(mlp_mps* f, int64_t* num_vars, int64_t* num_constraints, int64_t* nnz, int64_t* names_bytes);
/* Copies out whatever is non-NULL.  obj/mins/maxs: num_vars; row_ptr: num_constraints + 1; col_idx/vals: nnz; ops (0 Eq,
 * 1 Le, 2 Ge) / rhs: num_constraints; names_blob: names_bytes (variable names concatenated, MpsFile::variables),
 * names_off: num_vars + 1. */
mlp_status mlp_mps_export(mlp_mps* f, double* obj, double* mins, double* maxs, int64_t* row_ptr, int32_t* col_idx, double* vals,
                          int32_t* ops, double* rhs, char* names_blob, int64_t* names_off);

/* ===================================================================== sharding helpers (host, no GPU) */
/* Column block [begin, end) of rank `rank` of `world` over n structural columns. */
void mlp_shard_range(int64_t n, int32_t world, int32_t rank, int64_t* begin, int64_t* end);
/* Deterministic arg-reduce of per-rank pricing candidates (score, pos, var): highest score wins, ties go
 * to the lowest position (solver.rs:719).  A candidate with var < 0 is "none".  Returns winner rank or -1. */
int32_t mlp_reduce_candidates(const double* scores, const int64_t* pos, const int64_t* vars, int32_t world);

/* ===================================================================== synthetic dense LPs (host) */
/* Independent implementation of the generator the oracle defines in oracle/synth_lp.hpp. */
void mlp_synth_rows(int32_t kind, int64_t m, int64_t n, uint64_t seed, int64_t row0, int64_t nrows, int32_t threads,
                    double* out_rows);
/* the column block [col0, col0+ncols) of those rows (nrows x ncols row-major): what one shard uploads */
void mlp_synth_block(int32_t kind, int64_t m, int64_t n, uint64_t seed, int64_t row0, int64_t nrows, int64_t col0,
                     int64_t ncols, int32_t threads, double* out_block);
/* obj (user sign), mins, maxs: n; ops, rhs: m. Returns the optimization direction (0 min, 1 max).
 * Kind 3 needs A·x0 and therefore generates rows internally. */
int32_t mlp_synth_vectors(int32_t kind, int64_t m, int64_t n, uint64_t seed, double* obj, double* mins, double* maxs,
                          int32_t* ops, double* rhs);

#ifdef __cplusplus
}
#endif
#endif
