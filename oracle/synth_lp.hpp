// ORACLE — TEST INFRASTRUCTURE ONLY.
// Synthetic dense LP families used by the parity tests and the bench
// (SURVEY.md §8d).  The reference ships no generator; these are this repo's own
// definitions.  The product library carries an independent implementation
// (minilp_b200/csrc/synth.cpp) and tests assert the two are bit-identical.
//
//   u(seed, stream, idx) = top 53 bits of mix(mix(seed + K*(stream+1)) ^ idx) * 2^-53
//   mix = splitmix64 finaliser.   Streams: 0 A (idx = i*n+j), 1 c, 2 b, 3 x0, 4 row type.
//
//   kind 0 dense_pos   : max c'x, Ax <= b, x >= 0;  a=u, c=0.5+u, b=(n/4)(0.5+u)
//                        -> primal-feasible / dual-infeasible start: primal loop with primal SE
//   kind 1 dense_box   : max c'x, Ax <= b, 0<=x<=1; a=2u-1, c=0.5+u, b=0.25(0.5+u)*sqrt(n)
//                        -> every x starts at its upper bound (dual-feasible), ~1/3 of the rows violated: dual loop
//                           over non-basic variables sitting at upper bounds
//   kind 2 dense_cover : min c'x, Ax >= b, x >= 0;  a=u, c=0.5+u, b=(n/4)(0.5+u)
//                        -> dual loop only
//   kind 3 dense_mixed : max c'x, rows Le/Ge/Eq around a hidden x0 in [0,1]^n, last row sum(x) <= 2n;
//                        0<=x_j<=x0_j+0.05 for even j, x_j>=0 for odd j; a=2u-1, c=2u-1
//                        -> both infeasible at start: artificial objective, dual loop, recalc_obj_coeffs, primal
//                           loop with bound flips (solver.rs:841-852, 1031-1042)
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <thread>
#include <vector>

namespace synth {

inline uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t stream_key(uint64_t seed, uint64_t stream) { return mix(seed + 0x632BE59BD9B4E019ull * (stream + 1)); }
inline double u01(uint64_t key, uint64_t idx) { return (double)(mix(key ^ idx) >> 11) * 0x1.0p-53; }

struct DenseLP {
  int direction = 0;  // 0 minimize, 1 maximize
  std::vector<double> obj, mins, maxs, rhs;
  std::vector<int> ops;  // 0 Eq, 1 Le, 2 Ge
};

inline double a_entry(int kind, uint64_t keyA, uint64_t idx) {
  double u = u01(keyA, idx);
  return (kind == 1 || kind == 3) ? 2.0 * u - 1.0 : u;
}

inline void generate(int kind, std::size_t m, std::size_t n, uint64_t seed, int threads, DenseLP& lp, std::vector<double>& A) {
  const double inf = std::numeric_limits<double>::infinity();
  uint64_t kA = stream_key(seed, 0), kc = stream_key(seed, 1), kb = stream_key(seed, 2), kx = stream_key(seed, 3),
           kt = stream_key(seed, 4);
  A.resize(m * n);
  if (threads < 1) threads = 1;
  auto fill = [&](std::size_t r0, std::size_t r1) {
    for (std::size_t i = r0; i < r1; ++i)
      for (std::size_t j = 0; j < n; ++j) A[i * n + j] = a_entry(kind, kA, (uint64_t)(i * n + j));
  };
  if (threads == 1 || m < 64) fill(0, m);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(fill, m * t / threads, m * (t + 1) / threads);
    for (auto& x : th) x.join();
  }
  lp.obj.resize(n); lp.mins.assign(n, 0.0); lp.maxs.assign(n, inf); lp.rhs.resize(m); lp.ops.assign(m, 1);
  switch (kind) {
    case 0:
      lp.direction = 1;
      for (std::size_t j = 0; j < n; ++j) lp.obj[j] = 0.5 + u01(kc, j);
      for (std::size_t i = 0; i < m; ++i) lp.rhs[i] = ((double)n / 4.0) * (0.5 + u01(kb, i));
      break;
    case 1:
      lp.direction = 1;
      for (std::size_t j = 0; j < n; ++j) { lp.obj[j] = 0.5 + u01(kc, j); lp.maxs[j] = 1.0; }
      for (std::size_t i = 0; i < m; ++i) lp.rhs[i] = 0.25 * (0.5 + u01(kb, i)) * std::sqrt((double)n);
      break;
    case 2:
      lp.direction = 0;
      for (std::size_t j = 0; j < n; ++j) lp.obj[j] = 0.5 + u01(kc, j);
      for (std::size_t i = 0; i < m; ++i) { lp.rhs[i] = ((double)n / 4.0) * (0.5 + u01(kb, i)); lp.ops[i] = 2; }
      break;
    default: {
      lp.direction = 1;
      for (std::size_t j = 0; j < n; ++j) { lp.obj[j] = 2.0 * u01(kc, j) - 1.0; if (j % 2 == 0) lp.maxs[j] = u01(kx, j) + 0.05; }
      for (std::size_t j = 0; j < n; ++j) A[(m - 1) * n + j] = 1.0;  // bounding row
      for (std::size_t i = 0; i + 1 < m; ++i) {
        double ax = 0.0;
        for (std::size_t j = 0; j < n; ++j) ax += A[i * n + j] * u01(kx, j);  // sequential, unfused
        double t = u01(kt, i), s = u01(kb, i);
        if (t < 0.5) { lp.ops[i] = 1; lp.rhs[i] = ax + s; }
        else if (t < 0.8) { lp.ops[i] = 2; lp.rhs[i] = ax - s; }
        else { lp.ops[i] = 0; lp.rhs[i] = ax; }
      }
      lp.ops[m - 1] = 1;
      lp.rhs[m - 1] = 2.0 * (double)n;
    }
  }
}

}  // namespace synth
