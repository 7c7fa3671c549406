// Synthetic dense LP families for the parity tests and the bench.  The reference ships no
// generator; the definitions are this repo's own and are documented in DESIGN.md §workloads.  This is the
// product-side implementation (array-at-a-time, row-range oriented so a caller can stream A to the device
// through small pinned buffers); oracle/synth_lp.hpp holds an independent one and
// tests/test_synth.py asserts they agree bit for bit.
#include "minilp_b200.h"

#include <cmath>
#include <limits>
#include <thread>
#include <vector>

namespace {
struct Rng {
  uint64_t key;
  static uint64_t fin(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
  }
  Rng(uint64_t seed, uint64_t stream) : key(fin(seed + 0x632BE59BD9B4E019ull * (stream + 1))) {}
  double operator()(uint64_t idx) const { return std::ldexp((double)(fin(key ^ idx) >> 11), -53); }
};
bool signed_kind(int kind) { return kind == 1 || kind == 3; }

// rows [row0, row0+nrows) x columns [col0, col0+ncols) of A, written row-major with ncols per row
void fill_block(int kind, int64_t m, int64_t n, uint64_t seed, int64_t row0, int64_t nrows, int64_t col0, int64_t ncols,
                double* out) {
  const Rng ra(seed, 0);
  const bool sg = signed_kind(kind);
  for (int64_t i = 0; i < nrows; ++i) {
    const int64_t r = row0 + i;
    double* dst = out + i * ncols;
    if (kind == 3 && r == m - 1) {
      for (int64_t j = 0; j < ncols; ++j) dst[j] = 1.0;
      continue;
    }
    const uint64_t base = (uint64_t)r * (uint64_t)n + (uint64_t)col0;
    for (int64_t j = 0; j < ncols; ++j) {
      const double u = ra(base + (uint64_t)j);
      dst[j] = sg ? 2.0 * u - 1.0 : u;
    }
  }
}
void fill_rows(int kind, int64_t m, int64_t n, uint64_t seed, int64_t row0, int64_t nrows, double* out) {
  fill_block(kind, m, n, seed, row0, nrows, 0, n, out);
}
}  // namespace

extern "C" {

void mlp_synth_rows(int32_t kind, int64_t m, int64_t n, uint64_t seed, int64_t row0, int64_t nrows, int32_t threads,
                    double* out_rows) {
  if (threads < 2 || nrows < 2 * threads) { fill_rows(kind, m, n, seed, row0, nrows, out_rows); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    const int64_t a = nrows * t / threads, b = nrows * (t + 1) / threads;
    th.emplace_back(fill_rows, kind, m, n, seed, row0 + a, b - a, out_rows + a * n);
  }
  for (auto& x : th) x.join();
}

void mlp_synth_block(int32_t kind, int64_t m, int64_t n, uint64_t seed, int64_t row0, int64_t nrows, int64_t col0,
                     int64_t ncols, int32_t threads, double* out_block) {
  if (threads < 2 || nrows < 2 * threads) { fill_block(kind, m, n, seed, row0, nrows, col0, ncols, out_block); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    const int64_t a = nrows * t / threads, b = nrows * (t + 1) / threads;
    th.emplace_back(fill_block, kind, m, n, seed, row0 + a, b - a, col0, ncols, out_block + a * ncols);
  }
  for (auto& x : th) x.join();
}

int32_t mlp_synth_vectors(int32_t kind, int64_t m, int64_t n, uint64_t seed, double* obj, double* mins, double* maxs,
                          int32_t* ops, double* rhs) {
  const double inf = std::numeric_limits<double>::infinity();
  const Rng rc(seed, 1), rb(seed, 2), rx(seed, 3), rt(seed, 4);
  for (int64_t j = 0; j < n; ++j) { mins[j] = 0.0; maxs[j] = inf; }
  for (int64_t i = 0; i < m; ++i) ops[i] = 1;
  if (kind == 0 || kind == 2) {
    for (int64_t j = 0; j < n; ++j) obj[j] = 0.5 + rc(j);
    for (int64_t i = 0; i < m; ++i) { rhs[i] = ((double)n / 4.0) * (0.5 + rb(i)); if (kind == 2) ops[i] = 2; }
    return kind == 0 ? 1 : 0;
  }
  if (kind == 1) {
    for (int64_t j = 0; j < n; ++j) { obj[j] = 0.5 + rc(j); maxs[j] = 1.0; }
    const double sq = std::sqrt((double)n);
    for (int64_t i = 0; i < m; ++i) rhs[i] = 0.25 * (0.5 + rb(i)) * sq;
    return 1;
  }
  // kind 3: rows built around a hidden feasible point x0
  std::vector<double> x0(n), row(n);
  for (int64_t j = 0; j < n; ++j) {
    x0[j] = rx(j);
    obj[j] = 2.0 * rc(j) - 1.0;
    if (j % 2 == 0) maxs[j] = x0[j] + 0.05;
  }
  for (int64_t i = 0; i + 1 < m; ++i) {
    fill_rows(kind, m, n, seed, i, 1, row.data());
    double ax = 0.0;
    for (int64_t j = 0; j < n; ++j) ax += row[j] * x0[j];
    const double t = rt(i), s = rb(i);
    if (t < 0.5) { ops[i] = 1; rhs[i] = ax + s; }
    else if (t < 0.8) { ops[i] = 2; rhs[i] = ax - s; }
    else { ops[i] = 0; rhs[i] = ax; }
  }
  ops[m - 1] = 1;
  rhs[m - 1] = 2.0 * (double)n;
  return 1;
}

}  // extern "C"
