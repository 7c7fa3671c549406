// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement (C++17) of the revised-simplex path of ztlpn/minilp @ b99146b
// (pure Rust; no Rust toolchain exists in this image, so the reference itself
// cannot be executed here).  Every function cites the reference file:line it
// follows.  Iteration orders, strict/non-strict comparisons, constants and the
// order of floating-point operations are kept as in the reference so that the
// pivot sequence it produces is the reference's.  Build with
// `-O2 -ffp-contract=off` (rustc never fuses a*b+c).
//
// Parity pinning: the oracle is checked against every known-answer test the
// reference holds for this path (tests/test_oracle_golden.py): lib.rs:471-645,
// solver.rs:1392-1479, lu.rs:480-609, sparse.rs:345-359, mps.rs:437-476 and the
// README doctest lib.rs:28-44.  The reference pins END STATES only, never a
// pivot sequence, so "sequence parity" rests on this restatement being faithful.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may use anything in this directory.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace mlo {

using usize = std::size_t;
constexpr double EPS = 1e-8;  // solver.rs:12
constexpr double INF = std::numeric_limits<double>::infinity();

enum class Error { None = 0, Infeasible = 1, Unbounded = 2, Singular = 3, NonFinite = 4 };
struct SolveError : std::runtime_error {
  Error code;
  SolveError(Error c, const char* what) : std::runtime_error(what), code(c) {}
};
struct SingularMatrix : std::runtime_error {  // sparse.rs:335-338
  SingularMatrix() : std::runtime_error("singular matrix") {}
};
struct Panic : std::runtime_error {  // the reference panics here
  explicit Panic(const std::string& s) : std::runtime_error(s) {}
};

// ---------------------------------------------------------------- sparse.rs
// SparseVec, sparse.rs:5-39
struct SparseVec {
  std::vector<usize> indices;
  std::vector<double> values;
  void clear() { indices.clear(); values.clear(); }
  void push(usize i, double v) { indices.push_back(i); values.push_back(v); }
  usize len() const { return indices.size(); }
  double sq_norm() const {  // sparse.rs:32-34 (sum in storage order)
    double s = 0.0;
    for (double v : values) s += v * v;
    return s;
  }
};

// ScatteredVec, sparse.rs:41-136.  `nonzero` is insertion-ordered and is the
// iteration order of every ratio test downstream.
struct ScatteredVec {
  std::vector<double> values;
  std::vector<uint8_t> is_nonzero;
  std::vector<usize> nonzero;

  ScatteredVec() = default;
  explicit ScatteredVec(usize n) : values(n, 0.0), is_nonzero(n, 0) {}
  usize len() const { return values.size(); }
  double get(usize i) const { return values[i]; }
  double& get_mut(usize i) {  // sparse.rs:75-80
    if (!is_nonzero[i]) { is_nonzero[i] = 1; nonzero.push_back(i); }
    return values[i];
  }
  double sq_norm() const {  // sparse.rs:82-87
    double s = 0.0;
    for (usize i : nonzero) s += values[i] * values[i];
    return s;
  }
  void clear() {  // sparse.rs:89-95
    for (usize i : nonzero) { values[i] = 0.0; is_nonzero[i] = 0; }
    nonzero.clear();
  }
  void clear_and_resize(usize n) {  // sparse.rs:97-101
    clear();
    values.resize(n, 0.0);
    is_nonzero.resize(n, 0);
  }
  // sparse.rs:103-113; caller feeds (i, val) pairs through `put`.
  void begin_set() { clear(); }
  void put(usize i, double v) { is_nonzero[i] = 1; nonzero.push_back(i); values[i] = v; }
  void to_sparse_vec(SparseVec& out) const {  // sparse.rs:115-121
    out.clear();
    for (usize i : nonzero) out.push(i, values[i]);
  }
};

// SparseMat (append-only CSC), sparse.rs:138-270
struct SparseMat {
  usize n_rows = 0;
  std::vector<usize> indptr{0};
  std::vector<usize> indices;
  std::vector<double> data;

  SparseMat() = default;
  explicit SparseMat(usize rows) : n_rows(rows) {}
  usize rows() const { return n_rows; }
  usize cols() const { return indptr.size() - 1; }
  usize nnz() const { return data.size(); }
  void clear_and_resize(usize rows) {
    data.clear(); indices.clear(); indptr.assign(1, 0); n_rows = rows;
  }
  void push(usize row, double v) { indices.push_back(row); data.push_back(v); }
  void seal_column() { indptr.push_back(indices.size()); }
  usize col_begin(usize c) const { return indptr[c]; }
  usize col_end(usize c) const { return indptr[c + 1]; }

  // counting transpose, sparse.rs:230-269.  Note the placement loop walks
  // columns forward while filling each output row from its END, so entries of
  // an output column come out in DESCENDING source-column order.
  SparseMat transpose() const {
    SparseMat out;
    out.n_rows = cols();
    out.indptr.assign(rows() + 1, 0);
    for (usize c = 0; c < cols(); ++c)
      for (usize p = indptr[c]; p < indptr[c + 1]; ++p) out.indptr[indices[p]] += 1;
    for (usize r = 1; r < out.indptr.size(); ++r) out.indptr[r] += out.indptr[r - 1];
    out.indices.assign(nnz(), 0);
    out.data.assign(nnz(), 0.0);
    for (usize c = 0; c < cols(); ++c)
      for (usize p = indptr[c]; p < indptr[c + 1]; ++p) {
        usize r = indices[p];
        out.indptr[r] -= 1;
        out.indices[out.indptr[r]] = c;
        out.data[out.indptr[r]] = data[p];
      }
    out.indptr.back() = nnz();
    return out;
  }
};

// TriangleMat, sparse.rs:272-316.  has_diag == false means unit diagonal.
struct TriangleMat {
  SparseMat nondiag;
  bool has_diag = false;
  std::vector<double> diag;
  usize rows() const { return nondiag.rows(); }
  usize cols() const { return nondiag.cols(); }
  TriangleMat transpose() const {
    TriangleMat t;
    t.nondiag = nondiag.transpose();
    t.has_diag = has_diag;
    t.diag = diag;
    return t;
  }
};

struct Perm {  // sparse.rs:329-333
  std::vector<usize> orig2new, new2orig;
};

// -------------------------------------------------------------- ordering.rs
// order_simple, ordering.rs:4-21, with ColsQueue ordering.rs:394-460: a bucket
// queue keyed by (column length - 1), FIFO inside a bucket (add() links the new
// column just before the head of a circular list, pop_min() takes the head).
// A stable counting sort by score produces the identical sequence.
// `len - 1` on an empty column underflows usize in the reference (debug panic /
// out-of-bounds in release); restated as a singular-matrix error.
inline Perm order_simple(usize size, const std::vector<usize>& col_len) {
  std::vector<usize> count(size + 1, 0);
  for (usize c = 0; c < size; ++c) {
    if (col_len[c] == 0) throw SingularMatrix();
    usize score = col_len[c] - 1;
    if (score >= size) throw Panic("order_simple: score out of range");
    count[score + 1] += 1;
  }
  for (usize s = 1; s <= size; ++s) count[s] += count[s - 1];
  Perm p;
  p.new2orig.assign(size, 0);
  p.orig2new.assign(size, 0);
  for (usize c = 0; c < size; ++c) {
    usize score = col_len[c] - 1;
    p.new2orig[count[score]++] = c;
  }
  for (usize nw = 0; nw < size; ++nw) p.orig2new[p.new2orig[nw]] = nw;
  return p;
}

// -------------------------------------------------------------------- lu.rs
// MarkNonzero, lu.rs:306-407: iterative DFS giving reverse-topological order.
struct MarkNonzero {
  struct DfsStep { usize orig_i; usize cur_child; };
  std::vector<DfsStep> dfs_stack;
  std::vector<uint8_t> is_visited;
  std::vector<usize> visited;

  void clear() {
    for (usize i : visited) is_visited[i] = 0;
    visited.clear();
  }
  void clear_and_resize(usize n) { clear(); is_visited.resize(n, 0); }

  // children(new_i) -> (ptr,len) of row indices; filter(new_i); o2n(orig_i)
  template <class Children, class Filter, class O2N>
  void run(ScatteredVec& rhs, Children children, Filter filter, O2N o2n) {  // lu.rs:343-406
    clear();
    for (usize k = 0; k < rhs.nonzero.size(); ++k) {
      usize orig_r = rhs.nonzero[k];
      usize new_r = o2n(orig_r);
      if (!filter(new_r)) continue;
      if (is_visited[orig_r]) continue;
      dfs_stack.push_back({orig_r, 0});
      while (!dfs_stack.empty()) {
        DfsStep& cur = dfs_stack.back();
        usize new_i = o2n(cur.orig_i);
        const usize* ch = nullptr;
        usize nch = 0;
        if (filter(new_i)) {
          auto pr = children(new_i);
          ch = pr.first;
          nch = pr.second;
        }
        if (!is_visited[cur.orig_i]) is_visited[cur.orig_i] = 1;
        else cur.cur_child += 1;
        while (cur.cur_child < nch) {
          if (!is_visited[ch[cur.cur_child]]) break;
          cur.cur_child += 1;
        }
        if (cur.cur_child < nch) {
          usize child = ch[cur.cur_child];
          dfs_stack.push_back({child, 0});  // may invalidate `cur`; not used after
        } else {
          visited.push_back(cur.orig_i);
          dfs_stack.pop_back();
        }
      }
    }
    for (usize i : visited)
      if (!rhs.is_nonzero[i]) { rhs.is_nonzero[i] = 1; rhs.nonzero.push_back(i); }
  }
};

struct ScratchSpace {  // lu.rs:11-31
  ScatteredVec rhs;
  std::vector<double> dense_rhs;
  MarkNonzero mark_nonzero;
  explicit ScratchSpace(usize n = 0) : rhs(n), dense_rhs(n, 0.0) { mark_nonzero.is_visited.assign(n, 0); }
  void clear_sparse(usize size) {
    rhs.clear_and_resize(size);
    mark_nonzero.clear_and_resize(size);
  }
};

// tri_solve_process_col, lu.rs:450-463
inline void tri_solve_process_col(const TriangleMat& t, usize col, double* rhs) {
  double x = t.has_diag ? rhs[col] / t.diag[col] : rhs[col];
  rhs[col] = x;
  const SparseMat& nd = t.nondiag;
  for (usize p = nd.indptr[col]; p < nd.indptr[col + 1]; ++p) rhs[nd.indices[p]] -= x * nd.data[p];
}
// tri_solve_dense, lu.rs:414-429
inline void tri_solve_dense(const TriangleMat& t, bool lower, double* rhs) {
  if (lower) for (usize c = 0; c < t.cols(); ++c) tri_solve_process_col(t, c, rhs);
  else for (usize c = t.cols(); c-- > 0;) tri_solve_process_col(t, c, rhs);
}
// tri_solve_sparse, lu.rs:432-448
inline void tri_solve_sparse(const TriangleMat& t, ScratchSpace& s) {
  const SparseMat& nd = t.nondiag;
  s.mark_nonzero.run(
      s.rhs,
      [&](usize col) { return std::make_pair(nd.indices.data() + nd.indptr[col], nd.indptr[col + 1] - nd.indptr[col]); },
      [](usize) { return true; }, [](usize i) { return i; });
  for (usize k = s.mark_nonzero.visited.size(); k-- > 0;)
    tri_solve_process_col(t, s.mark_nonzero.visited[k], s.rhs.values.data());
}

struct LUFactors {  // lu.rs:3-116
  TriangleMat lower, upper;
  Perm row_perm, col_perm;  // always Some() when produced by lu_factorize

  usize nnz() const { return lower.nondiag.nnz() + upper.nondiag.nnz() + lower.cols(); }  // lu.rs:52-54

  void solve_dense(std::vector<double>& rhs, ScratchSpace& s) const {  // lu.rs:56-77
    s.dense_rhs.resize(rhs.size(), 0.0);
    for (usize i = 0; i < rhs.size(); ++i) s.dense_rhs[row_perm.orig2new[i]] = rhs[i];
    tri_solve_dense(lower, true, s.dense_rhs.data());
    tri_solve_dense(upper, false, s.dense_rhs.data());
    for (usize i = 0; i < rhs.size(); ++i) rhs[col_perm.new2orig[i]] = s.dense_rhs[i];
  }
  void solve(ScatteredVec& rhs, ScratchSpace& s) const {  // lu.rs:79-106
    s.rhs.clear();
    for (usize i : rhs.nonzero) {
      usize ni = row_perm.orig2new[i];
      s.rhs.nonzero.push_back(ni);
      s.rhs.is_nonzero[ni] = 1;
      s.rhs.values[ni] = rhs.values[i];
    }
    tri_solve_sparse(lower, s);
    tri_solve_sparse(upper, s);
    rhs.clear();
    for (usize i : s.rhs.nonzero) {
      usize ni = col_perm.new2orig[i];
      rhs.nonzero.push_back(ni);
      rhs.is_nonzero[ni] = 1;
      rhs.values[ni] = s.rhs.values[i];
    }
  }
  LUFactors transpose() const {  // lu.rs:108-115
    LUFactors t;
    t.lower = upper.transpose();
    t.upper = lower.transpose();
    t.row_perm = col_perm;
    t.col_perm = row_perm;
    return t;
  }
};

// lu_factorize, lu.rs:118-304 (left-looking Gilbert–Peierls, threshold pivoting).
// `Cols` provides: usize len(c); template<F> void for_each(c, F f /*(row,val)*/).
template <class Cols>
LUFactors lu_factorize(usize size, const Cols& cols, double stability_coeff, ScratchSpace& scratch) {
  std::vector<usize> col_len(size);
  for (usize c = 0; c < size; ++c) col_len[c] = cols.len(c);
  Perm col_perm = order_simple(size, col_len);  // lu.rs:140

  std::vector<usize> orig_row2elt_count(size, 0);  // lu.rs:142-147
  for (usize c = 0; c < size; ++c) cols.for_each(c, [&](usize r, double) { orig_row2elt_count[r] += 1; });

  scratch.clear_sparse(size);
  SparseMat lower(size), upper(size);
  std::vector<double> upper_diag;
  upper_diag.reserve(size);
  std::vector<usize> new2orig_row(size), orig2new_row(size);
  for (usize i = 0; i < size; ++i) new2orig_row[i] = orig2new_row[i] = i;

  for (usize i_col = 0; i_col < size; ++i_col) {
    scratch.rhs.begin_set();  // lu.rs:167
    cols.for_each(col_perm.new2orig[i_col], [&](usize r, double v) { scratch.rhs.put(r, v); });

    scratch.mark_nonzero.run(  // lu.rs:169-174
        scratch.rhs,
        [&](usize new_i) { return std::make_pair(lower.indices.data() + lower.indptr[new_i], lower.indptr[new_i + 1] - lower.indptr[new_i]); },
        [&](usize new_i) { return new_i < i_col; }, [&](usize orig_r) { return orig2new_row[orig_r]; });

    for (usize k = scratch.mark_nonzero.visited.size(); k-- > 0;) {  // lu.rs:179-188
      usize orig_i = scratch.mark_nonzero.visited[k];
      usize new_i = orig2new_row[orig_i];
      if (new_i < i_col) {
        double x_val = scratch.rhs.values[orig_i];
        for (usize p = lower.indptr[new_i]; p < lower.indptr[new_i + 1]; ++p)
          scratch.rhs.values[lower.indices[p]] -= x_val * lower.data[p];
      }
    }

    // threshold pivot choice, lu.rs:194-233
    double max_abs = 0.0;
    for (usize orig_r : scratch.rhs.nonzero) {
      if (orig2new_row[orig_r] < i_col) continue;
      double a = std::fabs(scratch.rhs.values[orig_r]);
      if (a > max_abs) max_abs = a;
    }
    if (max_abs < 1e-8) throw SingularMatrix();  // lu.rs:207
    if (!std::isnormal(max_abs)) throw Panic("lu_factorize: max_abs not normal");  // lu.rs:211
    bool have_best = false;
    usize best_orig_r = 0, best_elt_count = 0;
    for (usize orig_r : scratch.rhs.nonzero) {
      if (orig2new_row[orig_r] < i_col) continue;
      if (std::fabs(scratch.rhs.values[orig_r]) >= stability_coeff * max_abs) {
        usize ec = orig_row2elt_count[orig_r];
        if (!have_best || best_elt_count > ec) { have_best = true; best_orig_r = orig_r; best_elt_count = ec; }
      }
    }
    if (!have_best) throw Panic("lu_factorize: no pivot");  // lu.rs:232 unwrap
    usize pivot_orig_r = best_orig_r;
    double pivot_val = scratch.rhs.values[pivot_orig_r];

    {  // lu.rs:237-244
      usize row = i_col;
      usize orig_row = new2orig_row[row];
      usize pivot_row = orig2new_row[pivot_orig_r];
      std::swap(new2orig_row[row], new2orig_row[pivot_row]);
      std::swap(orig2new_row[orig_row], orig2new_row[pivot_orig_r]);
    }

    for (usize orig_r : scratch.rhs.nonzero) {  // lu.rs:248-263
      double val = scratch.rhs.values[orig_r];
      if (val == 0.0) continue;
      usize new_r = orig2new_row[orig_r];
      if (new_r < i_col) upper.push(new_r, val);
      else if (new_r == i_col) upper_diag.push_back(pivot_val);
      else lower.push(orig_r, val / pivot_val);
    }
    upper.seal_column();
    lower.seal_column();
  }

  for (usize& r : lower.indices) r = orig2new_row[r];  // lu.rs:270-274

  LUFactors f;
  f.lower.nondiag = std::move(lower);
  f.lower.has_diag = false;
  f.upper.nondiag = std::move(upper);
  f.upper.has_diag = true;
  f.upper.diag = std::move(upper_diag);
  f.row_perm.orig2new = std::move(orig2new_row);
  f.row_perm.new2orig = std::move(new2orig_row);
  f.col_perm = std::move(col_perm);
  return f;
}

// Column provider over explicit CSC arrays (used by the LU unit tests).
struct CscCols {
  const std::vector<usize>* ptr;
  const std::vector<usize>* idx;
  const std::vector<double>* val;
  std::vector<usize> pick;  // column c of the basis = matrix column pick[c]
  usize len(usize c) const { usize j = pick[c]; return (*ptr)[j + 1] - (*ptr)[j]; }
  template <class F> void for_each(usize c, F f) const {
    usize j = pick[c];
    for (usize p = (*ptr)[j]; p < (*ptr)[j + 1]; ++p) f((*idx)[p], (*val)[p]);
  }
};

// ------------------------------------------------------------ constraint matrix
// The reference keeps `[A | I]` twice: CSR (`orig_constraints`) and CSC
// (`orig_constraints_csc`, produced by sprs `to_csc`: rows ascending inside a
// column), both with usize indices (solver.rs:21-22, 247-253).
struct CsMatrix {  // faithful storage
  usize n_rows = 0, n_cols = 0;  // n_cols = num_total_vars
  std::vector<usize> r_ptr{0}, r_idx;
  std::vector<double> r_val;
  std::vector<usize> c_ptr, c_idx;
  std::vector<double> c_val;

  usize rows() const { return n_rows; }
  usize nnz() const { return r_val.size(); }
  void append_row(const std::vector<usize>& idx, const std::vector<double>& val) {
    r_idx.insert(r_idx.end(), idx.begin(), idx.end());
    r_val.insert(r_val.end(), val.begin(), val.end());
    r_ptr.push_back(r_idx.size());
    n_rows += 1;
  }
  void build_csc() {  // sprs to_csc: stable counting transpose, rows ascending per column
    c_ptr.assign(n_cols + 1, 0);
    for (usize j : r_idx) c_ptr[j + 1] += 1;
    for (usize j = 0; j < n_cols; ++j) c_ptr[j + 1] += c_ptr[j];
    c_idx.assign(nnz(), 0);
    c_val.assign(nnz(), 0.0);
    std::vector<usize> next(c_ptr.begin(), c_ptr.end() - 1);
    for (usize r = 0; r < n_rows; ++r)
      for (usize p = r_ptr[r]; p < r_ptr[r + 1]; ++p) {
        usize q = next[r_idx[p]]++;
        c_idx[q] = r;
        c_val[q] = r_val[p];
      }
  }
  template <class F> void for_row(usize r, F f) const {
    for (usize p = r_ptr[r]; p < r_ptr[r + 1]; ++p) f(r_idx[p], r_val[p]);
  }
  template <class F> void for_col(usize v, F f) const {
    for (usize p = c_ptr[v]; p < c_ptr[v + 1]; ++p) f(c_idx[p], c_val[p]);
  }
  usize col_len(usize v) const { return c_ptr[v + 1] - c_ptr[v]; }
  // sprs squared_l2_norm of every column: sum of squares in storage (ascending row) order
  void col_sq_norms(usize nvars, std::vector<double>& out) const {
    out.assign(nvars, 0.0);
    for (usize v = 0; v < nvars; ++v) for_col(v, [&](usize, double x) { out[v] += x * x; });
  }
};

// Memory-lean storage for fully dense A (every a_ij is a stored entry, exactly
// as if the caller had listed all n terms of every row): one row-major array,
// the slack identity implicit.  Iteration orders are identical to CsMatrix, so
// the arithmetic — and therefore every result — is bit-identical
// (tests/test_oracle_golden.py::test_dense_storage_matches_faithful).  Exists
// because the faithful layout of a 50k x 50k LP needs ~120 GB.
struct DenseMatrix {
  usize n_rows = 0, n_struct = 0;
  const double* a = nullptr;  // row-major n_rows x n_struct (not owned)
  usize rows() const { return n_rows; }
  usize nnz() const { return n_rows * n_struct + n_rows; }
  template <class F> void for_row(usize r, F f) const {
    const double* row = a + r * n_struct;
    for (usize v = 0; v < n_struct; ++v) f(v, row[v]);
    f(n_struct + r, 1.0);
  }
  template <class F> void for_col(usize v, F f) const {
    if (v < n_struct) for (usize r = 0; r < n_rows; ++r) f(r, a[r * n_struct + v]);
    else f(v - n_struct, 1.0);
  }
  usize col_len(usize v) const { return v < n_struct ? n_rows : 1; }
  // same per-column accumulation order (rows ascending) as CsMatrix::col_sq_norms, swept row-major for locality
  void col_sq_norms(usize nvars, std::vector<double>& out) const {
    out.assign(nvars, 0.0);
    for (usize r = 0; r < n_rows; ++r) {
      const double* row = a + r * n_struct;
      for (usize v = 0; v < n_struct && v < nvars; ++v) out[v] += row[v] * row[v];
    }
    for (usize v = n_struct; v < nvars; ++v) out[v] = 1.0;
  }
};

enum class ComparisonOp { Eq = 0, Le = 1, Ge = 2 };  // lib.rs:160-169
enum class Direction { Minimize = 0, Maximize = 1 };  // lib.rs:61-68

// sprs CsVec::new (lib.rs:279, 376): sorts by index, panics on duplicates / out of range.
struct CsVec {
  usize dim = 0;
  std::vector<usize> indices;
  std::vector<double> data;
  static CsVec make(usize dim, std::vector<usize> idx, std::vector<double> val) {
    std::vector<usize> perm(idx.size());
    for (usize i = 0; i < perm.size(); ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](usize a, usize b) { return idx[a] < idx[b]; });
    CsVec v;
    v.dim = dim;
    for (usize p : perm) { v.indices.push_back(idx[p]); v.data.push_back(val[p]); }
    for (usize i = 0; i < v.indices.size(); ++i) {
      if (v.indices[i] >= dim) throw Panic("CsVec: index out of range");
      if (i > 0 && v.indices[i] == v.indices[i - 1]) throw Panic("CsVec: duplicate index");  // lib.rs:249
    }
    return v;
  }
};
struct Constraint { CsVec coeffs; ComparisonOp op; double rhs; };

// One record per successful pivot() call (solver.rs:1023), for sequence parity.
struct PivotRecord {
  int32_t phase;          // 0 = dual loop (restore_feasibility), 1 = primal loop (optimize)
  int64_t entering_var, entering_col;
  int64_t leaving_row;    // -1: bound flip (no basis change)
  int64_t leaving_var;    // -1 on bound flip
  double pivot_coeff, entering_diff, obj_after;
  int64_t eta_count, lu_nnz, nnz_col, nnz_rho;
  int32_t refactored;
};

struct VarState { bool basic; usize idx; };           // solver.rs:60-64
struct NonBasicVarState { bool at_min, at_max; };     // solver.rs:66-70

struct PivotElem { usize row; double coeff; double leaving_new_val; };  // solver.rs:1256-1261
struct PivotInfo {                                                        // solver.rs:1245-1254
  usize col; double entering_new_val; double entering_diff; bool has_elem; PivotElem elem;
};

struct EtaMatrices {  // solver.rs:1341-1368
  std::vector<usize> leaving_rows;
  SparseMat coeff_cols;
  usize len() const { return leaving_rows.size(); }
  void clear_and_resize(usize n) { leaving_rows.clear(); coeff_cols.clear_and_resize(n); }
};

template <class Mat>
struct BasisCols {  // get_col closure of solver.rs:307-312 / 1292-1297
  const Mat* mat;
  const std::vector<usize>* basic_vars;
  usize len(usize c) const { return mat->col_len((*basic_vars)[c]); }
  template <class F> void for_each(usize c, F f) const { mat->for_col((*basic_vars)[c], f); }
};

template <class Mat>
struct BasisSolver {  // solver.rs:1263-1339
  LUFactors lu_factors, lu_factors_transp;
  ScratchSpace scratch;
  EtaMatrices eta_matrices;
  ScatteredVec rhs;

  void push_eta_matrix(const SparseVec& col_coeffs, usize r_leaving, double pivot_coeff) {  // 1274-1284
    eta_matrices.leaving_rows.push_back(r_leaving);
    for (usize k = 0; k < col_coeffs.len(); ++k) {
      usize r = col_coeffs.indices[k];
      double coeff = col_coeffs.values[k];
      double val = (r == r_leaving) ? 1.0 - 1.0 / pivot_coeff : coeff / pivot_coeff;
      eta_matrices.coeff_cols.push(r, val);
    }
    eta_matrices.coeff_cols.seal_column();
  }
  void reset(const Mat& csc, const std::vector<usize>& basic_vars) {  // 1286-1303
    scratch.clear_sparse(basic_vars.size());
    eta_matrices.clear_and_resize(basic_vars.size());
    rhs.clear_and_resize(basic_vars.size());
    BasisCols<Mat> cols{&csc, &basic_vars};
    lu_factors = lu_factorize(basic_vars.size(), cols, 0.1, scratch);  // singular => reference panics (1301)
    lu_factors_transp = lu_factors.transpose();
  }
  // FTRAN, solver.rs:1305-1319.  Caller has already filled `rhs` via begin_set/put.
  ScatteredVec& solve_loaded() {
    lu_factors.solve(rhs, scratch);
    const SparseMat& E = eta_matrices.coeff_cols;
    for (usize idx = 0; idx < eta_matrices.len(); ++idx) {
      usize r_leaving = eta_matrices.leaving_rows[idx];
      double coeff = rhs.get(r_leaving);
      for (usize p = E.indptr[idx]; p < E.indptr[idx + 1]; ++p) rhs.get_mut(E.indices[p]) -= coeff * E.data[p];
    }
    return rhs;
  }
  // BTRAN, solver.rs:1322-1338
  ScatteredVec& solve_transp_loaded() {
    const SparseMat& E = eta_matrices.coeff_cols;
    for (usize idx = eta_matrices.len(); idx-- > 0;) {
      double coeff = 0.0;
      for (usize p = E.indptr[idx]; p < E.indptr[idx + 1]; ++p) coeff += E.data[p] * rhs.get(E.indices[p]);
      usize r_leaving = eta_matrices.leaving_rows[idx];
      rhs.get_mut(r_leaving) -= coeff;
    }
    lu_factors_transp.solve(rhs, scratch);
    return rhs;
  }
};

struct SolverOptions {
  // GPU engine tie-break rules: where the reference resolves an
  // exact tie by list order, take the lowest row / variable index instead.
  // Off = reference behaviour.  Exact ties are counted either way.
  bool tie_lowest_index = false;
  bool record_trace = true;
};

template <class Mat>
struct Solver {  // solver.rs:14-58
  usize num_vars = 0;
  std::vector<double> orig_obj_coeffs, orig_var_mins, orig_var_maxs;
  Mat mat;  // orig_constraints + orig_constraints_csc
  std::vector<double> orig_rhs;
  bool enable_primal_steepest_edge = false, enable_dual_steepest_edge = false;
  bool is_primal_feasible = false, is_dual_feasible = false;
  std::vector<VarState> var_states;
  BasisSolver<Mat> basis_solver;
  std::vector<usize> basic_vars;
  std::vector<double> basic_var_vals, basic_var_mins, basic_var_maxs, dual_edge_sq_norms;
  std::vector<usize> nb_vars;
  std::vector<double> nb_var_obj_coeffs, nb_var_vals;
  std::vector<NonBasicVarState> nb_var_states;
  std::vector<uint8_t> nb_var_is_fixed;
  std::vector<double> primal_edge_sq_norms;
  double cur_obj_val = 0.0;
  SparseVec col_coeffs;
  std::vector<double> sq_norms_update_helper;
  SparseVec inv_basis_row_coeffs;
  ScatteredVec row_coeffs;

  // instrumentation (not in the reference)
  SolverOptions opts;
  std::vector<PivotRecord> trace;
  int64_t pivots_done = 0, refactor_count = 0, tie_events = 0;
  // Ties that MATTER for the pass-2 winner of the two ratio tests (804-823, 982-1002): pivots whose winning |coeff| is
  // shared exactly by another eligible candidate (tied_pivots: the reference then decides by list order) or approached
  // within NEAR_TIE relative (near_tie_pivots >= tied_pivots: the decision is within rounding of the summation order).
  static constexpr double NEAR_TIE = 1e-9;
  int64_t tied_pivots = 0, near_tie_pivots = 0, first_tied_pivot = -1, first_near_tie_pivot = -1;
  // The two SELECTIONS (pricing 696-739, dual row 855-917) use the same rule on both sides — strict '>', lowest index on exact
  // ties — so only rounding can make them differ: a pivot is recorded here when the runner-up's score is within NEAR_TIE of the
  // winner's (the order of two such scores is not defined beyond the rounding of the sums that produced them).
  int64_t sel_near_tie_pivots = 0, first_sel_near_tie_pivot = -1;
  void note_selection(double best, double second) {
    if (second >= best * (1.0 - NEAR_TIE) && best > 0.0) {
      sel_near_tie_pivots += 1;
      if (first_sel_near_tie_pivot < 0) first_sel_near_tie_pivot = pivots_done;
    }
  }
  void note_ties(int64_t exact_cnt, int64_t near_cnt) {  // counts include the winner itself
    if (exact_cnt > 1) { tied_pivots += 1; if (first_tied_pivot < 0) first_tied_pivot = pivots_done; }
    if (near_cnt > 1) { near_tie_pivots += 1; if (first_near_tie_pivot < 0) first_near_tie_pivot = pivots_done; }
  }
  int32_t cur_phase = 1;
  bool last_refactored = false;

  usize num_constraints() const { return mat.rows(); }
  usize num_total_vars() const { return num_vars + num_constraints(); }

  // The part of try_new that does not depend on matrix storage, solver.rs:116-187.
  struct InitVars { std::vector<double> vals; double obj_val; bool dual_feasible; };
  void init_vars(const std::vector<double>& obj, const std::vector<double>& mins, const std::vector<double>& maxs,
                 InitVars& out) {
    num_vars = obj.size();
    orig_var_mins = mins;
    orig_var_maxs = maxs;
    out.obj_val = 0.0;
    out.dual_feasible = true;
    for (usize v = 0; v < num_vars; ++v) {
      double mn = mins[v], mx = maxs[v];
      if (mn > mx) throw SolveError(Error::Infeasible, "min > max");  // 138-140
      var_states.push_back({false, nb_vars.size()});
      nb_vars.push_back(v);
      double init_val;
      if (mn == mx) init_val = mn;
      else if (std::isinf(mn) && std::isinf(mx)) { if (obj[v] != 0.0) out.dual_feasible = false; init_val = 0.0; }
      else if (obj[v] > 0.0) { if (std::isfinite(mn)) init_val = mn; else { out.dual_feasible = false; init_val = mx; } }
      else if (obj[v] < 0.0) { if (std::isfinite(mx)) init_val = mx; else { out.dual_feasible = false; init_val = mn; } }
      else if (std::isfinite(mn)) init_val = mn;
      else init_val = mx;
      nb_var_vals.push_back(init_val);
      out.obj_val += init_val * obj[v];
      nb_var_states.push_back({init_val == mn, init_val == mx});
    }
  }

  static void slack_bounds(ComparisonOp op, double& mn, double& mx) {  // 218-222
    switch (op) {
      case ComparisonOp::Le: mn = 0.0; mx = INF; break;
      case ComparisonOp::Ge: mn = -INF; mx = 0.0; break;
      default: mn = 0.0; mx = 0.0; break;
    }
  }
  static bool tautology_or_throw(ComparisonOp op, double rhs) {  // 201-213 / 558-570
    bool t = (op == ComparisonOp::Eq) ? (0.0 == rhs) : (op == ComparisonOp::Le) ? (0.0 <= rhs) : (0.0 >= rhs);
    if (!t) throw SolveError(Error::Infeasible, "empty infeasible constraint");
    return true;
  }

  // Second half of try_new, solver.rs:241-357, after `mat` has been filled.
  void finish_init(const std::vector<double>& obj, const InitVars& iv) {
    usize num_constraints_ = num_constraints();
    usize total = num_vars + num_constraints_;
    orig_obj_coeffs = obj;
    orig_obj_coeffs.resize(total, 0.0);
    is_dual_feasible = iv.dual_feasible;
    is_primal_feasible = true;
    for (usize r = 0; r < basic_var_vals.size(); ++r)
      if (!(basic_var_vals[r] >= basic_var_mins[r] && basic_var_vals[r] <= basic_var_maxs[r])) is_primal_feasible = false;
    bool need_artificial_obj = !is_primal_feasible && !is_dual_feasible;  // 261
    bool enable_steepest_edge = true;                                     // 114
    enable_dual_steepest_edge = enable_steepest_edge;
    if (enable_dual_steepest_edge) dual_edge_sq_norms.assign(basic_vars.size(), 1.0);
    enable_primal_steepest_edge = enable_steepest_edge && !is_dual_feasible;  // 272
    if (enable_primal_steepest_edge) sq_norms_update_helper.assign(total - num_constraints_, 0.0);
    std::vector<double> col_norms;
    if (enable_primal_steepest_edge) mat.col_sq_norms(total, col_norms);
    for (usize k = 0; k < nb_vars.size(); ++k) {  // 281-300
      usize var = nb_vars[k];
      const NonBasicVarState& st = nb_var_states[k];
      if (need_artificial_obj) {
        double c = (st.at_min && !st.at_max) ? 1.0 : (st.at_max && !st.at_min) ? -1.0 : 0.0;
        nb_var_obj_coeffs.push_back(c);
      } else nb_var_obj_coeffs.push_back(orig_obj_coeffs[var]);
      if (enable_primal_steepest_edge) primal_edge_sq_norms.push_back(col_norms[var] + 1.0);  // 297-299
    }
    cur_obj_val = need_artificial_obj ? 0.0 : iv.obj_val;
    basis_solver.scratch = ScratchSpace(num_constraints_);  // 304
    BasisCols<Mat> cols{&mat, &basic_vars};
    basis_solver.lu_factors = lu_factorize(basic_vars.size(), cols, 0.1, basis_solver.scratch);  // 305-316
    basis_solver.lu_factors_transp = basis_solver.lu_factors.transpose();
    basis_solver.eta_matrices.clear_and_resize(num_constraints_);
    basis_solver.rhs = ScatteredVec(num_constraints_);
    nb_var_is_fixed.assign(nb_vars.size(), 0);
    row_coeffs = ScatteredVec(total - num_constraints_);
  }

  double get_value(usize var) const {  // 371-376
    const VarState& s = var_states[var];
    return s.basic ? basic_var_vals[s.idx] : nb_var_vals[s.idx];
  }

  // ---- loops.  The reference's `optimize` (487-511) and `restore_feasibility`
  // (513-547) are unbounded loops; here one call performs one iteration so that
  // benches and parity tests can stop after a pivot budget.
  // Returns true if a pivot was performed, false if the loop terminated.
  bool primal_iteration() {  // body of optimize(), 497-506
    cur_phase = 1;
    PivotInfo pi;
    if (!choose_pivot(pi)) return false;
    pivot(pi);
    return true;
  }
  bool dual_iteration() {  // body of restore_feasibility(), 529-542
    cur_phase = 0;
    usize row; double leaving_new_val;
    if (!choose_pivot_row_dual(row, leaving_new_val)) return false;
    calc_row_coeffs(row);
    PivotInfo pi = choose_entering_col_dual(row, leaving_new_val);
    calc_col_coeffs(pi.col);
    pivot(pi);
    return true;
  }
  void optimize() { while (primal_iteration()) {} is_dual_feasible = true; }            // 487-511
  void restore_feasibility() { while (dual_iteration()) {} is_primal_feasible = true; } // 513-547

  // initial_solve, solver.rs:470-485, as a resumable state machine with a pivot budget.
  int solve_stage = 0;  // 0 start, 1 dual loop, 2 recalc, 3 primal loop, 4 done
  // Runs until done or until `max_pivots` more pivots have been made. Returns true when done.
  bool initial_solve_budget(int64_t max_pivots) {
    int64_t target = (max_pivots < 0) ? -1 : pivots_done + max_pivots;
    auto budget_left = [&] { return target < 0 || pivots_done < target; };
    for (;;) {
      switch (solve_stage) {
        case 0: solve_stage = is_primal_feasible ? 2 : 1; break;
        case 1:
          while (budget_left()) { if (!dual_iteration()) { is_primal_feasible = true; solve_stage = 2; break; } }
          if (solve_stage == 1) return false;
          break;
        case 2:
          if (!is_dual_feasible) { recalc_obj_coeffs(); solve_stage = 3; } else solve_stage = 4;
          break;
        case 3:
          while (budget_left()) { if (!primal_iteration()) { is_dual_feasible = true; solve_stage = 4; break; } }
          if (solve_stage == 3) return false;
          break;
        default:
          enable_primal_steepest_edge = false;  // 482
          return true;
      }
    }
  }
  void initial_solve() { initial_solve_budget(-1); }

  // calc_col_coeffs, 671-677
  void calc_col_coeffs(usize c_var) {
    usize var = nb_vars[c_var];
    basis_solver.rhs.begin_set();
    mat.for_col(var, [&](usize r, double v) { basis_solver.rhs.put(r, v); });
    basis_solver.solve_loaded().to_sparse_vec(col_coeffs);
  }
  // calc_row_coeffs, 680-693
  void calc_row_coeffs(usize r_constr) {
    basis_solver.rhs.begin_set();
    basis_solver.rhs.put(r_constr, 1.0);
    basis_solver.solve_transp_loaded().to_sparse_vec(inv_basis_row_coeffs);
    row_coeffs.clear_and_resize(nb_vars.size());
    for (usize k = 0; k < inv_basis_row_coeffs.len(); ++k) {
      usize r = inv_basis_row_coeffs.indices[k];
      double coeff = inv_basis_row_coeffs.values[k];
      mat.for_row(r, [&](usize v, double val) {
        const VarState& s = var_states[v];
        if (!s.basic) row_coeffs.get_mut(s.idx) += val * coeff;
      });
    }
  }

  // choose_pivot, 695-853.  Returns false when no entering column exists (optimal).
  bool choose_pivot(PivotInfo& out) {
    bool have_col = false;
    usize entering_c = 0;
    double best_score = -INF, second_score = -INF;
    for (usize col = 0; col < nb_var_obj_coeffs.size(); ++col) {
      double obj_coeff = nb_var_obj_coeffs[col];
      const NonBasicVarState& st = nb_var_states[col];
      if ((st.at_min && obj_coeff > -EPS) || (st.at_max && obj_coeff < EPS)) continue;  // 705-708
      double score = enable_primal_steepest_edge ? obj_coeff * obj_coeff / primal_edge_sq_norms[col] : std::fabs(obj_coeff);
      if (score > best_score) { have_col = true; entering_c = col; second_score = best_score; best_score = score; }
      else if (score > second_score) second_score = score;  // instrumentation only
    }
    if (!have_col) return false;
    note_selection(best_score, second_score);

    double entering_cur_val = nb_var_vals[entering_c];
    bool entering_diff_sign = nb_var_obj_coeffs[entering_c] < 0.0;  // 743
    double entering_other_val = entering_diff_sign ? orig_var_maxs[nb_vars[entering_c]] : orig_var_mins[nb_vars[entering_c]];
    calc_col_coeffs(entering_c);  // 750

    auto toward_max = [&](double coeff) { return (entering_diff_sign && coeff < 0.0) || (!entering_diff_sign && coeff > 0.0); };
    auto leaving_step = [&](usize r, double coeff) -> double {  // 752-771
      double val = basic_var_vals[r];
      if (toward_max(coeff)) { double mx = basic_var_maxs[r]; return val < mx ? mx - val : 0.0; }
      double mn = basic_var_mins[r];
      return val > mn ? val - mn : 0.0;
    };

    double max_step = std::fabs(entering_other_val - entering_cur_val);  // 782
    for (usize k = 0; k < col_coeffs.len(); ++k) {
      double coeff = col_coeffs.values[k];
      double coeff_abs = std::fabs(coeff);
      if (coeff_abs < EPS) continue;
      double cur_step = (leaving_step(col_coeffs.indices[k], coeff) + EPS) / coeff_abs;  // 791
      if (cur_step < max_step) max_step = cur_step;
    }

    bool have_row = false;
    usize leaving_r = 0;
    double leaving_new_val = 0.0, pivot_coeff_abs = -INF, pivot_coeff = 0.0;
    for (usize k = 0; k < col_coeffs.len(); ++k) {  // 804-823
      usize r = col_coeffs.indices[k];
      double coeff = col_coeffs.values[k];
      double coeff_abs = std::fabs(coeff);
      if (coeff_abs < EPS) continue;
      double cur_step = leaving_step(r, coeff) / coeff_abs;
      if (!(cur_step <= max_step)) continue;
      bool take = coeff_abs > pivot_coeff_abs;
      if (have_row && coeff_abs == pivot_coeff_abs) { tie_events += 1; if (opts.tie_lowest_index && r < leaving_r) take = true; }
      if (take) {
        have_row = true;
        leaving_r = r;
        leaving_new_val = toward_max(coeff) ? basic_var_maxs[r] : basic_var_mins[r];
        pivot_coeff = coeff;
        pivot_coeff_abs = coeff_abs;
      }
    }

    if (have_row) {
      int64_t ex = 0, nr = 0;  // instrumentation only: how contested was the winner
      for (usize k = 0; k < col_coeffs.len(); ++k) {
        double coeff = col_coeffs.values[k], coeff_abs = std::fabs(coeff);
        if (coeff_abs < EPS) continue;
        if (!(leaving_step(col_coeffs.indices[k], coeff) / coeff_abs <= max_step)) continue;
        if (coeff_abs == pivot_coeff_abs) ex += 1;
        if (coeff_abs >= pivot_coeff_abs * (1.0 - NEAR_TIE)) nr += 1;
      }
      note_ties(ex, nr);
      calc_row_coeffs(leaving_r);  // 826
      double entering_diff = (basic_var_vals[leaving_r] - leaving_new_val) / pivot_coeff;
      out = PivotInfo{entering_c, entering_cur_val + entering_diff, entering_diff, true, {leaving_r, pivot_coeff, leaving_new_val}};
    } else {
      if (std::isinf(entering_other_val)) throw SolveError(Error::Unbounded, "unbounded");  // 842-844
      out = PivotInfo{entering_c, entering_other_val, entering_other_val - entering_cur_val, false, {0, 0.0, 0.0}};
    }
    return true;
  }

  // choose_pivot_row_dual, 855-917
  bool choose_pivot_row_dual(usize& row_out, double& new_val_out) const {
    bool have = false;
    usize leaving_r = 0;
    double max_score = -INF, second = -INF;
    for (usize r = 0; r < basic_var_vals.size(); ++r) {
      double val = basic_var_vals[r], mn = basic_var_mins[r], mx = basic_var_maxs[r];
      double infeas;
      if (val < mn - EPS) infeas = mn - val;
      else if (val > mx + EPS) infeas = val - mx;
      else continue;
      double score = enable_dual_steepest_edge ? infeas * infeas / dual_edge_sq_norms[r] : infeas;
      if (score > max_score) { have = true; leaving_r = r; second = max_score; max_score = score; }
      else if (score > second) second = score;  // instrumentation only
    }
    if (!have) return false;
    const_cast<Solver*>(this)->note_selection(max_score, second);
    double val = basic_var_vals[leaving_r];
    if (val < basic_var_mins[leaving_r]) new_val_out = basic_var_mins[leaving_r];
    else if (val > basic_var_maxs[leaving_r]) new_val_out = basic_var_maxs[leaving_r];
    else throw Panic("choose_pivot_row_dual: unreachable");
    row_out = leaving_r;
    return true;
  }

  // choose_entering_col_dual, 919-1021
  PivotInfo choose_entering_col_dual(usize row, double leaving_new_val) {
    bool leaving_diff_sign = leaving_new_val > basic_var_vals[row];  // 925
    auto clamp_obj = [](double oc, const NonBasicVarState& st) {      // 927-935
      if (st.at_min && oc < 0.0) oc = 0.0;
      if (st.at_max && oc > 0.0) oc = 0.0;
      return oc;
    };
    auto eligible = [&](double coeff, const NonBasicVarState& st) -> bool {  // 937-951
      bool entering_diff_sign;
      if (coeff >= EPS) entering_diff_sign = !leaving_diff_sign;
      else if (coeff <= -EPS) entering_diff_sign = leaving_diff_sign;
      else return false;
      return entering_diff_sign ? !st.at_max : !st.at_min;
    };
    double max_step = INF;
    for (usize c : row_coeffs.nonzero) {  // 963-974
      double coeff = row_coeffs.values[c];
      const NonBasicVarState& st = nb_var_states[c];
      if (!eligible(coeff, st)) continue;
      double oc = clamp_obj(nb_var_obj_coeffs[c], st);
      double cur_step = (std::fabs(oc) + EPS) / std::fabs(coeff);
      if (cur_step < max_step) max_step = cur_step;
    }
    bool have = false;
    usize entering_c = 0;
    double pivot_coeff_abs = -INF, pivot_coeff = 0.0;
    for (usize c : row_coeffs.nonzero) {  // 982-1002
      double coeff = row_coeffs.values[c];
      const NonBasicVarState& st = nb_var_states[c];
      if (!eligible(coeff, st)) continue;
      double oc = clamp_obj(nb_var_obj_coeffs[c], st);
      double cur_step = std::fabs(oc) / std::fabs(coeff);
      if (cur_step <= max_step) {
        double coeff_abs = std::fabs(coeff);
        bool take = coeff_abs > pivot_coeff_abs;
        if (have && coeff_abs == pivot_coeff_abs) { tie_events += 1; if (opts.tie_lowest_index && nb_vars[c] < nb_vars[entering_c]) take = true; }
        if (take) { have = true; entering_c = c; pivot_coeff_abs = coeff_abs; pivot_coeff = coeff; }
      }
    }
    if (!have) throw SolveError(Error::Infeasible, "infeasible");  // 1019
    {
      int64_t ex = 0, nr = 0;  // instrumentation only
      for (usize c : row_coeffs.nonzero) {
        double coeff = row_coeffs.values[c];
        const NonBasicVarState& st = nb_var_states[c];
        if (!eligible(coeff, st)) continue;
        double oc = clamp_obj(nb_var_obj_coeffs[c], st);
        if (!(std::fabs(oc) / std::fabs(coeff) <= max_step)) continue;
        if (std::fabs(coeff) == pivot_coeff_abs) ex += 1;
        if (std::fabs(coeff) >= pivot_coeff_abs * (1.0 - NEAR_TIE)) nr += 1;
      }
      note_ties(ex, nr);
    }
    double entering_diff = (basic_var_vals[row] - leaving_new_val) / pivot_coeff;
    return PivotInfo{entering_c, nb_var_vals[entering_c] + entering_diff, entering_diff, true, {row, pivot_coeff, leaving_new_val}};
  }

  // pivot, 1023-1104
  void pivot(const PivotInfo& pi) {
    cur_obj_val += nb_var_obj_coeffs[pi.col] * pi.entering_diff;  // 1027
    usize entering_var = nb_vars[pi.col];
    last_refactored = false;
    if (!pi.has_elem) {  // 1031-1042
      nb_var_vals[pi.col] = pi.entering_new_val;
      for (usize k = 0; k < col_coeffs.len(); ++k) basic_var_vals[col_coeffs.indices[k]] -= pi.entering_diff * col_coeffs.values[k];
      nb_var_states[pi.col].at_min = pi.entering_new_val == orig_var_mins[entering_var];
      nb_var_states[pi.col].at_max = pi.entering_new_val == orig_var_maxs[entering_var];
      record(pi, entering_var, -1);
      return;
    }
    const PivotElem& pe = pi.elem;
    double pivot_coeff = pe.coeff;
    for (usize k = 0; k < col_coeffs.len(); ++k) {  // 1049-1055
      usize r = col_coeffs.indices[k];
      if (r == pe.row) basic_var_vals[r] = pi.entering_new_val;
      else basic_var_vals[r] -= pi.entering_diff * col_coeffs.values[k];
    }
    basic_var_mins[pe.row] = orig_var_mins[entering_var];
    basic_var_maxs[pe.row] = orig_var_maxs[entering_var];
    if (enable_dual_steepest_edge) update_dual_sq_norms(pe.row, pivot_coeff);  // 1060-1062

    usize leaving_var = basic_vars[pe.row];
    nb_var_vals[pi.col] = pe.leaving_new_val;
    nb_var_states[pi.col].at_min = pe.leaving_new_val == orig_var_mins[leaving_var];
    nb_var_states[pi.col].at_max = pe.leaving_new_val == orig_var_maxs[leaving_var];

    double pivot_obj = nb_var_obj_coeffs[pi.col] / pivot_coeff;  // 1073
    for (usize c : row_coeffs.nonzero) {
      if (c == pi.col) nb_var_obj_coeffs[c] = -pivot_obj;
      else nb_var_obj_coeffs[c] -= pivot_obj * row_coeffs.values[c];
    }
    if (enable_primal_steepest_edge) update_primal_sq_norms(pi.col, pivot_coeff);  // 1082-1084

    basic_vars[pe.row] = entering_var;  // 1088-1091
    var_states[entering_var] = {true, pe.row};
    nb_vars[pi.col] = leaving_var;
    var_states[leaving_var] = {false, pi.col};

    usize eta_nnz = basis_solver.eta_matrices.coeff_cols.nnz();  // 1096-1103
    if (eta_nnz < basis_solver.lu_factors.nnz()) {
      basis_solver.push_eta_matrix(col_coeffs, pe.row, pivot_coeff);
    } else {
      basis_solver.reset(mat, basic_vars);
      refactor_count += 1;
      last_refactored = true;
    }
    record(pi, entering_var, (int64_t)leaving_var);
  }

  void record(const PivotInfo& pi, usize entering_var, int64_t leaving_var) {
    pivots_done += 1;
    if (!opts.record_trace) return;
    PivotRecord r;
    r.phase = cur_phase;
    r.entering_var = (int64_t)entering_var;
    r.entering_col = (int64_t)pi.col;
    r.leaving_row = pi.has_elem ? (int64_t)pi.elem.row : -1;
    r.leaving_var = leaving_var;
    r.pivot_coeff = pi.has_elem ? pi.elem.coeff : 0.0;
    r.entering_diff = pi.entering_diff;
    r.obj_after = cur_obj_val;
    r.eta_count = (int64_t)basis_solver.eta_matrices.len();
    r.lu_nnz = (int64_t)basis_solver.lu_factors.nnz();
    r.nnz_col = (int64_t)col_coeffs.len();
    r.nnz_rho = (int64_t)inv_basis_row_coeffs.len();
    r.refactored = last_refactored ? 1 : 0;
    trace.push_back(r);
  }

  // update_primal_sq_norms, 1106-1151
  void update_primal_sq_norms(usize entering_col, double pivot_coeff) {
    basis_solver.rhs.begin_set();
    for (usize k = 0; k < col_coeffs.len(); ++k) basis_solver.rhs.put(col_coeffs.indices[k], col_coeffs.values[k]);
    ScatteredVec& tmp = basis_solver.solve_transp_loaded();
    for (usize r : tmp.nonzero)
      mat.for_row(r, [&](usize v, double) { const VarState& s = var_states[v]; if (!s.basic) sq_norms_update_helper[s.idx] = 0.0; });
    for (usize r : tmp.nonzero) {
      double coeff = tmp.values[r];
      mat.for_row(r, [&](usize v, double val) { const VarState& s = var_states[v]; if (!s.basic) sq_norms_update_helper[s.idx] += val * coeff; });
    }
    double pivot_sq_norm = col_coeffs.sq_norm() + 1.0;  // 1136
    double pivot_coeff_sq = pivot_coeff * pivot_coeff;
    for (usize c : row_coeffs.nonzero) {
      double r_coeff = row_coeffs.values[c];
      if (c == entering_col) primal_edge_sq_norms[c] = pivot_sq_norm / pivot_coeff_sq;
      else
        primal_edge_sq_norms[c] += -2.0 * r_coeff * sq_norms_update_helper[c] / pivot_coeff + pivot_sq_norm * r_coeff * r_coeff / pivot_coeff_sq;
      if (!std::isfinite(primal_edge_sq_norms[c])) throw SolveError(Error::NonFinite, "primal sq norm not finite");  // 1149
    }
  }
  // update_dual_sq_norms, 1153-1174
  void update_dual_sq_norms(usize leaving_row, double pivot_coeff) {
    basis_solver.rhs.begin_set();
    for (usize k = 0; k < inv_basis_row_coeffs.len(); ++k) basis_solver.rhs.put(inv_basis_row_coeffs.indices[k], inv_basis_row_coeffs.values[k]);
    ScatteredVec& tau = basis_solver.solve_loaded();
    double pivot_sq_norm = inv_basis_row_coeffs.sq_norm();
    double pivot_coeff_sq = pivot_coeff * pivot_coeff;
    for (usize k = 0; k < col_coeffs.len(); ++k) {
      usize r = col_coeffs.indices[k];
      double col_coeff = col_coeffs.values[k];
      if (r == leaving_row) dual_edge_sq_norms[r] = pivot_sq_norm / pivot_coeff_sq;
      else dual_edge_sq_norms[r] += -2.0 * col_coeff * tau.get(r) / pivot_coeff + pivot_sq_norm * col_coeff * col_coeff / pivot_coeff_sq;
      if (!std::isfinite(dual_edge_sq_norms[r])) throw SolveError(Error::NonFinite, "dual sq norm not finite");  // 1172
    }
  }
  // recalc_obj_coeffs, 1199-1231
  void recalc_obj_coeffs() {
    if (basis_solver.eta_matrices.len() > 0) { basis_solver.reset(mat, basic_vars); refactor_count += 1; }
    std::vector<double> multipliers(num_constraints(), 0.0);
    for (usize c = 0; c < basic_vars.size(); ++c) multipliers[c] = orig_obj_coeffs[basic_vars[c]];
    basis_solver.lu_factors_transp.solve_dense(multipliers, basis_solver.scratch);
    nb_var_obj_coeffs.clear();
    for (usize var : nb_vars) {
      double dot = 0.0;
      mat.for_col(var, [&](usize r, double val) { dot += val * multipliers[r]; });
      nb_var_obj_coeffs.push_back(orig_obj_coeffs[var] - dot);
    }
    cur_obj_val = 0.0;
    for (usize r = 0; r < basic_vars.size(); ++r) cur_obj_val += orig_obj_coeffs[basic_vars[r]] * basic_var_vals[r];
    for (usize c = 0; c < nb_vars.size(); ++c) cur_obj_val += orig_obj_coeffs[nb_vars[c]] * nb_var_vals[c];
  }

  // fix_var, 378-415
  void fix_var(usize var, double val) {
    if (val < orig_var_mins[var] || val > orig_var_maxs[var]) throw SolveError(Error::Infeasible, "fix_var out of bounds");
    usize col;
    if (var_states[var].basic) {
      usize row = var_states[var].idx;
      cur_phase = 0;
      calc_row_coeffs(row);
      PivotInfo pi = choose_entering_col_dual(row, val);
      calc_col_coeffs(pi.col);
      pivot(pi);
      col = pi.col;
    } else {
      col = var_states[var].idx;
      calc_col_coeffs(col);
      double diff = val - nb_var_vals[col];
      for (usize k = 0; k < col_coeffs.len(); ++k) basic_var_vals[col_coeffs.indices[k]] -= diff * col_coeffs.values[k];
      cur_obj_val += diff * nb_var_obj_coeffs[col];
      nb_var_vals[col] = val;
    }
    nb_var_states[col] = {true, true};
    nb_var_is_fixed[col] = 1;
    is_primal_feasible = false;
    restore_feasibility();
  }
  // unfix_var, 418-438
  bool unfix_var(usize var) {
    if (var_states[var].basic) return false;
    usize col = var_states[var].idx;
    bool was = nb_var_is_fixed[col];
    nb_var_is_fixed[col] = 0;
    if (!was) return false;
    double cur_val = nb_var_vals[col];
    nb_var_states[col] = {cur_val == orig_var_mins[var], cur_val == orig_var_maxs[var]};
    is_dual_feasible = false;
    optimize();
    return true;
  }
};

// add_constraint (solver.rs:549-634) and add_gomory_cut (440-460) need the
// faithful storage (the reference rebuilds the CSR with one more column).
inline void solver_add_constraint(Solver<CsMatrix>& s, CsVec coeffs, ComparisonOp op, double rhs) {
  if (!s.is_primal_feasible || !s.is_dual_feasible) throw Panic("add_constraint: not optimal");  // 555-556
  if (coeffs.indices.empty()) { Solver<CsMatrix>::tautology_or_throw(op, rhs); return; }
  usize slack_var = s.num_total_vars();
  double smin, smax;
  Solver<CsMatrix>::slack_bounds(op, smin, smax);
  s.orig_obj_coeffs.push_back(0.0);
  s.orig_var_mins.push_back(smin);
  s.orig_var_maxs.push_back(smax);
  s.var_states.push_back({true, s.basic_vars.size()});
  s.basic_vars.push_back(slack_var);
  s.basic_var_mins.push_back(smin);
  s.basic_var_maxs.push_back(smax);
  double lhs_val = 0.0;
  for (usize k = 0; k < coeffs.indices.size(); ++k) lhs_val += s.get_value(coeffs.indices[k]) * coeffs.data[k];  // 587-594
  s.basic_var_vals.push_back(rhs - lhs_val);

  usize new_total = s.num_total_vars() + 1;  // 597 (num_constraints not yet grown)
  // every stored row keeps its entries (all indices < old total < new_total); new row appended
  while (!coeffs.indices.empty() && coeffs.indices.back() >= new_total) { coeffs.indices.pop_back(); coeffs.data.pop_back(); }
  coeffs.indices.push_back(slack_var);
  coeffs.data.push_back(1.0);
  s.mat.n_cols = new_total;
  s.mat.append_row(coeffs.indices, coeffs.data);
  s.orig_rhs.push_back(rhs);
  s.mat.build_csc();  // 610
  s.basis_solver.reset(s.mat, s.basic_vars);  // 612
  s.refactor_count += 1;
  if (s.enable_primal_steepest_edge || s.enable_dual_steepest_edge) {  // 615-630
    s.calc_row_coeffs(s.num_constraints() - 1);
    if (s.enable_primal_steepest_edge)
      for (usize c : s.row_coeffs.nonzero) s.primal_edge_sq_norms[c] += s.row_coeffs.values[c] * s.row_coeffs.values[c];
    if (s.enable_dual_steepest_edge) s.dual_edge_sq_norms.push_back(s.inv_basis_row_coeffs.sq_norm());
  }
  s.is_primal_feasible = false;
  s.restore_feasibility();
}

inline void solver_add_gomory_cut(Solver<CsMatrix>& s, usize var) {  // 440-460
  if (!s.var_states[var].basic) throw Panic("add_gomory_cut: var is not basic");
  usize row = s.var_states[var].idx;
  s.calc_row_coeffs(row);
  std::vector<usize> idx;
  std::vector<double> val;
  for (usize c : s.row_coeffs.nonzero) {
    double coeff = s.row_coeffs.values[c];
    idx.push_back(s.nb_vars[c]);
    val.push_back(std::floor(coeff) - coeff);
  }
  double cut_bound = std::floor(s.basic_var_vals[row]) - s.basic_var_vals[row];
  solver_add_constraint(s, CsVec::make(s.num_total_vars(), idx, val), ComparisonOp::Le, cut_bound);
}

// try_new for the faithful storage, solver.rs:108-369.
inline void solver_init_sparse(Solver<CsMatrix>& s, const std::vector<double>& obj, const std::vector<double>& mins,
                               const std::vector<double>& maxs, const std::vector<Constraint>& constraints) {
  Solver<CsMatrix>::InitVars iv;
  s.init_vars(obj, mins, maxs, iv);
  std::vector<const CsVec*> kept;
  for (const Constraint& c : constraints) {  // 198-239
    if (c.coeffs.indices.empty()) { Solver<CsMatrix>::tautology_or_throw(c.op, c.rhs); continue; }
    kept.push_back(&c.coeffs);
    s.orig_rhs.push_back(c.rhs);
    double smin, smax;
    Solver<CsMatrix>::slack_bounds(c.op, smin, smax);
    s.orig_var_mins.push_back(smin);
    s.orig_var_maxs.push_back(smax);
    s.basic_var_mins.push_back(smin);
    s.basic_var_maxs.push_back(smax);
    usize slack = s.var_states.size();
    s.var_states.push_back({true, s.basic_vars.size()});
    s.basic_vars.push_back(slack);
    double lhs = 0.0;
    for (usize k = 0; k < c.coeffs.indices.size(); ++k) lhs += c.coeffs.data[k] * s.nb_var_vals[c.coeffs.indices[k]];
    s.basic_var_vals.push_back(c.rhs - lhs);
  }
  usize total = s.num_vars + kept.size();
  s.mat.n_cols = total;  // 247-253
  for (usize i = 0; i < kept.size(); ++i) {
    std::vector<usize> idx = kept[i]->indices;
    std::vector<double> val = kept[i]->data;
    idx.push_back(s.num_vars + i);
    val.push_back(1.0);
    s.mat.append_row(idx, val);
  }
  s.mat.build_csc();
  s.finish_init(obj, iv);
}

// try_new for dense storage: every row has all n structural entries; no empty rows.
inline void solver_init_dense(Solver<DenseMatrix>& s, const std::vector<double>& obj, const std::vector<double>& mins,
                              const std::vector<double>& maxs, usize m, const double* a_rowmajor,
                              const std::vector<int>& ops, const std::vector<double>& rhs) {
  Solver<DenseMatrix>::InitVars iv;
  s.init_vars(obj, mins, maxs, iv);
  usize n = s.num_vars;
  for (usize i = 0; i < m; ++i) {
    s.orig_rhs.push_back(rhs[i]);
    double smin, smax;
    Solver<DenseMatrix>::slack_bounds((ComparisonOp)ops[i], smin, smax);
    s.orig_var_mins.push_back(smin);
    s.orig_var_maxs.push_back(smax);
    s.basic_var_mins.push_back(smin);
    s.basic_var_maxs.push_back(smax);
    usize slack = s.var_states.size();
    s.var_states.push_back({true, s.basic_vars.size()});
    s.basic_vars.push_back(slack);
    double lhs = 0.0;
    const double* row = a_rowmajor + i * n;
    for (usize v = 0; v < n; ++v) lhs += row[v] * s.nb_var_vals[v];
    s.basic_var_vals.push_back(rhs[i] - lhs);
  }
  s.mat.n_rows = m;
  s.mat.n_struct = n;
  s.mat.a = a_rowmajor;
  s.finish_init(obj, iv);
}

// ------------------------------------------------------------------- lib.rs
struct Problem {  // lib.rs:192-305
  Direction direction = Direction::Minimize;
  std::vector<double> obj_coeffs, var_mins, var_maxs;
  std::vector<Constraint> constraints;
  usize add_var(double obj_coeff, double mn, double mx) {  // 233-243
    usize v = obj_coeffs.size();
    obj_coeffs.push_back(direction == Direction::Minimize ? obj_coeff : -obj_coeff);
    var_mins.push_back(mn);
    var_maxs.push_back(mx);
    return v;
  }
  void add_constraint(const std::vector<usize>& vars, const std::vector<double>& coeffs, ComparisonOp op, double rhs) {  // 276-283
    constraints.push_back({CsVec::make(obj_coeffs.size(), vars, coeffs), op, rhs});
  }
};

struct Solution {  // lib.rs:313-424
  Direction direction = Direction::Minimize;
  usize num_vars = 0;
  Solver<CsMatrix> solver;
  double objective() const { return direction == Direction::Minimize ? solver.cur_obj_val : -solver.cur_obj_val; }  // 334-339
  double var_value(usize v) const {
    if (v >= num_vars) throw Panic("var_value: out of range");
    return solver.get_value(v);
  }
  void add_constraint(const std::vector<usize>& vars, const std::vector<double>& coeffs, ComparisonOp op, double rhs) {  // 368-381
    solver_add_constraint(solver, CsVec::make(num_vars, vars, coeffs), op, rhs);
  }
  void fix_var(usize v, double val) { if (v >= num_vars) throw Panic("fix_var"); solver.fix_var(v, val); }
  bool unfix_var(usize v) { if (v >= num_vars) throw Panic("unfix_var"); return solver.unfix_var(v); }
  void add_gomory_cut(usize v) { if (v >= num_vars) throw Panic("gomory"); solver_add_gomory_cut(solver, v); }
};

inline void problem_begin_solve(const Problem& p, Solution& out, const SolverOptions& opts = SolverOptions()) {
  out.direction = p.direction;
  out.num_vars = p.obj_coeffs.size();
  out.solver.opts = opts;
  solver_init_sparse(out.solver, p.obj_coeffs, p.var_mins, p.var_maxs, p.constraints);
}
inline void problem_solve(const Problem& p, Solution& out, const SolverOptions& opts = SolverOptions()) {  // lib.rs:291-304
  problem_begin_solve(p, out, opts);
  out.solver.initial_solve();
}

// ------------------------------------------------------------------- mps.rs
// MpsFile::parse, mps.rs:39-329 (free format).
struct MpsFile {
  std::string problem_name;
  std::vector<std::string> var_names;  // index = Variable
  std::unordered_map<std::string, usize> variables;
  Problem problem;
};
struct MpsError : std::runtime_error { explicit MpsError(const std::string& s) : std::runtime_error(s) {} };

inline MpsFile parse_mps(const std::string& text, Direction direction) {
  struct Lines {
    const std::string& t; usize pos = 0; std::string cur; usize idx = 0;
    explicit Lines(const std::string& s) : t(s) {}
    void to_next() {  // mps.rs:339-357: skip '*' comments and blank lines; EOF => empty cur
      for (;;) {
        idx += 1;
        cur.clear();
        if (pos >= t.size()) return;
        usize e = t.find('\n', pos);
        if (e == std::string::npos) { cur = t.substr(pos); pos = t.size(); }
        else { cur = t.substr(pos, e - pos + 1); pos = e + 1; }
        if (cur.empty()) return;
        if (cur[0] == '*') continue;
        usize len = cur.size();
        while (len > 0 && std::isspace((unsigned char)cur[len - 1])) --len;
        if (len != 0) { cur.resize(len); return; }
      }
    }
    MpsError err(const std::string& m) const { return MpsError("line " + std::to_string(idx) + ": " + m); }
  } lines(text);
  auto tokenize = [](const std::string& s) {
    std::vector<std::string> out;
    usize i = 0;
    while (i < s.size()) {
      while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
      usize j = i;
      while (j < s.size() && !std::isspace((unsigned char)s[j])) ++j;
      if (j > i) out.push_back(s.substr(i, j - i));
      i = j;
    }
    return out;
  };
  auto parse_f64 = [&](const std::string& s) -> double {
    try {
      usize used = 0;
      double v = std::stod(s, &used);
      if (used != s.size()) throw std::invalid_argument("x");
      return v;
    } catch (...) { throw lines.err("couldn't parse float from string: `" + s + "`"); }
  };
  struct Tok {
    std::vector<std::string> v; usize i = 0; const Lines* L;
    const std::string& next() { if (i >= v.size()) throw L->err("unexpected end of line"); return v[i++]; }
    bool has() const { return i < v.size(); }
  };
  auto kv_pairs = [&](Tok& t) {  // mps.rs:402-431
    std::vector<std::pair<std::string, double>> out;
    std::string k1 = t.next();
    double v1 = parse_f64(t.next());
    out.push_back({k1, v1});
    if (t.has()) { std::string k2 = t.next(); double v2 = parse_f64(t.next()); out.push_back({k2, v2}); }
    return out;
  };
  auto starts_with_space = [](const std::string& s) { return !s.empty() && s[0] == ' '; };

  MpsFile mf;
  lines.to_next();
  { Tok t{tokenize(lines.cur), 0, &lines}; if (t.next() != "NAME") throw lines.err("expected NAME section"); mf.problem_name = t.has() ? t.next() : ""; }

  struct ConstraintDef { std::vector<usize> vars; std::vector<double> coeffs; ComparisonOp op; double rhs = 0.0, range = 0.0; };
  bool have_obj = false;
  std::string obj_name;
  std::unordered_set<std::string> free_rows;
  std::vector<ConstraintDef> cdefs;
  std::unordered_map<std::string, usize> cname2idx;
  lines.to_next();
  if (lines.cur != "ROWS") throw lines.err("expected ROWS section");
  for (;;) {
    lines.to_next();
    if (!starts_with_space(lines.cur)) break;
    Tok t{tokenize(lines.cur), 0, &lines};
    std::string row_type = t.next();
    std::string name = t.next();
    ComparisonOp op;
    if (row_type == "N") { if (!have_obj) { have_obj = true; obj_name = name; } else free_rows.insert(name); continue; }
    else if (row_type == "L") op = ComparisonOp::Le;
    else if (row_type == "G") op = ComparisonOp::Ge;
    else if (row_type == "E") op = ComparisonOp::Eq;
    else throw lines.err("unexpected row type " + row_type);
    if (!cname2idx.emplace(name, cdefs.size()).second) throw lines.err("row " + name + " already declared");
    ConstraintDef d; d.op = op; cdefs.push_back(d);
  }
  if (!have_obj) throw lines.err("objective function name not declared");

  struct VarDef { bool has_min = false, has_max = false; double mn = 0, mx = 0, obj = 0; };
  std::vector<VarDef> vdefs;
  if (lines.cur != "COLUMNS") throw lines.err("expected COLUMNS section");
  {
    usize cur_var = 0; std::string cur_name; VarDef cur_def;
    for (;;) {
      lines.to_next();
      if (!starts_with_space(lines.cur)) break;
      Tok t{tokenize(lines.cur), 0, &lines};
      std::string name = t.next();
      if (name != cur_name) {
        if (mf.variables.count(name)) throw lines.err("variable " + name + " already declared");
        if (!cur_name.empty()) { mf.variables[cur_name] = cur_var; mf.var_names.push_back(cur_name); vdefs.push_back(cur_def); cur_def = VarDef(); cur_var += 1; }
        cur_name = name;
      }
      for (auto& kv : kv_pairs(t)) {
        if (kv.first == obj_name) cur_def.obj = kv.second;
        else if (cname2idx.count(kv.first)) { ConstraintDef& d = cdefs[cname2idx[kv.first]]; d.vars.push_back(cur_var); d.coeffs.push_back(kv.second); }
        else if (!free_rows.count(kv.first)) throw lines.err("unknown constraint: " + kv.first);
      }
    }
    if (!cur_name.empty()) { mf.variables[cur_name] = cur_var; mf.var_names.push_back(cur_name); vdefs.push_back(cur_def); }
  }
  if (lines.cur != "RHS") throw lines.err("expected RHS section");
  {
    bool have_vec = false; std::string vec;
    for (;;) {
      lines.to_next();
      if (!starts_with_space(lines.cur)) break;
      Tok t{tokenize(lines.cur), 0, &lines};
      std::string vn = t.next();
      if (!have_vec) { have_vec = true; vec = vn; } else if (vec != vn) continue;  // first RHS vector only
      for (auto& kv : kv_pairs(t)) {
        if (kv.first == obj_name) throw lines.err("setting objective in RHS section is not supported");
        else if (cname2idx.count(kv.first)) cdefs[cname2idx[kv.first]].rhs = kv.second;
        else throw lines.err("unknown constraint: " + kv.first);
      }
    }
  }
  if (lines.cur == "RANGES") {
    bool have_vec = false; std::string vec;
    for (;;) {
      lines.to_next();
      if (!starts_with_space(lines.cur)) break;
      Tok t{tokenize(lines.cur), 0, &lines};
      std::string vn = t.next();
      if (!have_vec) { have_vec = true; vec = vn; } else if (vec != vn) continue;
      for (auto& kv : kv_pairs(t)) {
        if (cname2idx.count(kv.first)) cdefs[cname2idx[kv.first]].range = kv.second;
        else throw lines.err("unknown constraint: " + kv.first);
      }
    }
  }
  if (lines.cur == "BOUNDS") {
    bool have_vec = false; std::string vec;
    for (;;) {
      lines.to_next();
      if (!starts_with_space(lines.cur)) break;
      Tok t{tokenize(lines.cur), 0, &lines};
      std::string bt = t.next();
      std::string vn = t.next();
      if (!have_vec) { have_vec = true; vec = vn; } else if (vec != vn) continue;
      std::string var_name = t.next();
      if (!mf.variables.count(var_name)) throw lines.err("unknown variable: " + var_name);
      VarDef& d = vdefs[mf.variables[var_name]];
      if (bt == "FR") { d.has_min = d.has_max = true; d.mn = -INF; d.mx = INF; continue; }
      double val = parse_f64(t.next());
      if (bt == "LO") { d.has_min = true; d.mn = val; }
      else if (bt == "UP") { d.has_max = true; d.mx = val; }
      else if (bt == "FX") { d.has_min = d.has_max = true; d.mn = d.mx = val; }
      else throw lines.err("bound type " + bt + " is not supported");
    }
  }
  if (lines.cur != "ENDATA") throw lines.err("expected ENDATA section");

  mf.problem.direction = direction;
  for (const VarDef& d : vdefs) {  // mps.rs:294-303
    double mn, mx;
    if (d.has_min && d.has_max) { mn = d.mn; mx = d.mx; }
    else if (d.has_min) { mn = d.mn; mx = INF; }
    else if (d.has_max && d.mx < 0.0) { mn = -INF; mx = d.mx; }
    else if (d.has_max) { mn = 0.0; mx = d.mx; }
    else { mn = 0.0; mx = INF; }
    mf.problem.add_var(d.obj, mn, mx);
  }
  for (const ConstraintDef& c : cdefs) {  // mps.rs:305-322
    if (c.range == 0.0) mf.problem.add_constraint(c.vars, c.coeffs, c.op, c.rhs);
    else {
      double mn, mx;
      if (c.op == ComparisonOp::Ge) { mn = c.rhs; mx = c.rhs + std::fabs(c.range); }
      else if (c.op == ComparisonOp::Le) { mn = c.rhs - std::fabs(c.range); mx = c.rhs; }
      else if (c.range > 0.0) { mn = c.rhs; mx = c.rhs + c.range; }
      else { mn = c.rhs + c.range; mx = c.rhs; }
      mf.problem.add_constraint(c.vars, c.coeffs, ComparisonOp::Ge, mn);
      mf.problem.add_constraint(c.vars, c.coeffs, ComparisonOp::Le, mx);
    }
  }
  return mf;
}

}  // namespace mlo
