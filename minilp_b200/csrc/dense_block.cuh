// Throughput-bound dense building blocks for LARGE cores (k in the thousands to tens of thousands: config 4 deep into its
// solve).  The per-column substitution kernels of kernels_common.cuh (k_core_inverse*) are latency-bound — one CTA walks 2k
// dependent steps per column of the inverse — which is the right trade below k ~ 2000 and hopeless beyond (measured: 18 s per
// refactorization at k = 15 500, profiles/r02_deep_curve_config4.md).  Here the explicit inverse X = U^-1 L^-1 is formed by
// BLOCKED substitution on all columns at once: per 32-row block one small triangular solve (k_tri_block_*) and one rank-32
// update of the remaining rows (k_gemm_sub), i.e. 2 k^3/3 multiply-adds at FP64 pipe rate instead of k^2 latency chains.
// Deterministic (fixed tiling and summation order).  The inverse is an engine-internal quantity — the reference solves with
// L and U — so its rank-32 updates may contract a*b+c (__fma_rn, twice the pipe rate); k_gemm_sub<false> keeps the separate
// multiply and subtract of the LU itself.
#pragma once

constexpr int GB_T = 64;   // C tile: 64 x 64 per CTA of 256 threads, 4 x 4 per thread
constexpr int GB_K = 32;   // inner dimension of one update (<= 32)
// C[M x N] (ldc) -= A[M x kb] (lda) * B[kb x N] (ldb); column-major; per element the products are subtracted in ascending t.
template <bool FMA>
__global__ void __launch_bounds__(256) k_gemm_sub(int M, int N, int kb, const double* __restrict__ A, int64_t lda,
                                                  const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc) {
  pdl_wait();
  __shared__ double As[GB_K][GB_T + 1];  // As[t][i]
  __shared__ double Bs[GB_K][GB_T + 1];  // Bs[t][j]
  const int i0 = blockIdx.x * GB_T, j0 = blockIdx.y * GB_T;
  const int tid = threadIdx.x;
  for (int q = tid; q < GB_K * GB_T; q += 256) {
    const int i = q % GB_T, t = q / GB_T;  // consecutive threads: consecutive rows of A (coalesced)
    As[t][i] = (t < kb && i0 + i < M) ? A[(int64_t)t * lda + i0 + i] : 0.0;
  }
  for (int q = tid; q < GB_K * GB_T; q += 256) {
    const int t = q % GB_K, j = q / GB_K;  // consecutive threads: consecutive t of one column of B (coalesced)
    Bs[t][j] = (t < kb && j0 + j < N) ? B[(int64_t)(j0 + j) * ldb + t] : 0.0;
  }
  __syncthreads();
  const int ti = (tid & 15) * 4, tj = (tid >> 4) * 4;  // 16 x 16 threads, 4 x 4 outputs each
  double acc[4][4];
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ti + a, j = j0 + tj + b;
      acc[b][a] = (i < M && j < N) ? C[(int64_t)j * ldc + i] : 0.0;
    }
#pragma unroll 8
  for (int t = 0; t < GB_K; ++t) {
    double av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = As[t][ti + a];
#pragma unroll
    for (int b = 0; b < 4; ++b) bv[b] = Bs[t][tj + b];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (FMA) acc[b][a] = __fma_rn(-av[a], bv[b], acc[b][a]);
        else acc[b][a] -= av[a] * bv[b];
      }
  }
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ti + a, j = j0 + tj + b;
      if (i < M && j < N) C[(int64_t)j * ldc + i] = acc[b][a];
    }
}

// X[r0 .. r0+nb, j] <- T^-1 X[r0 .. r0+nb, j] for ncols columns j (one thread per column), T = the nb x nb diagonal block of
// the factors at (r0, r0): LOWER: unit lower triangle (L), forward substitution; else upper triangle with its diagonal (U),
// back substitution.  nb <= 32.
template <bool LOWER>
__global__ void __launch_bounds__(128) k_tri_block(const double* __restrict__ LU, int64_t ld, int r0, int nb, double* __restrict__ X,
                                                   int64_t ldx, int ncols) {
  pdl_wait();
  __shared__ double T[32][33];
  for (int q = threadIdx.x; q < nb * nb; q += blockDim.x) {
    const int i = q % nb, t = q / nb;
    T[i][t] = LU[(int64_t)(r0 + t) * ld + r0 + i];
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  double* xp = X + (int64_t)j * ldx + r0;
  double x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = i < nb ? xp[i] : 0.0;
  if (LOWER) {
#pragma unroll
    for (int i = 1; i < 32; ++i) {
      if (i < nb) {
        double v = x[i];
#pragma unroll
        for (int t = 0; t < i; ++t) v -= T[i][t] * x[t];
        x[i] = v;
      }
    }
  } else {
#pragma unroll
    for (int i = 31; i >= 0; --i) {
      if (i < nb) {
        double v = x[i];
#pragma unroll
        for (int t = 31; t > i; --t)
          if (t < nb) v -= T[i][t] * x[t];
        x[i] = v / T[i][i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < nb) xp[i] = x[i];
}
// X <- identity (k x k, leading dimension ld)
__global__ void k_set_identity(double* __restrict__ X, int64_t ld, int k) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < k && j < k) X[(int64_t)j * ld + i] = (i == j) ? 1.0 : 0.0;
}
