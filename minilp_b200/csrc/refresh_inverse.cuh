// Product-form refresh of the explicit core inverse: the eta file folded into C^-1 without a new factorization.
// Included by engine.cu only (sparse storage).
//
// The reference refactorizes from scratch whenever the eta file has grown as large as the factors (solver.rs:1096-1103 ->
// BasisSolver::reset 1286-1303 -> lu_factorize lu.rs:118-304): on netlib-like LPs every ~16 pivots.  A factorization is k
// dependent column steps plus O(k^3) work for the explicit inverse this engine solves with — 60 % of the step on config 4
// (profiles/r02d_*).  But between two refactorizations the basis changes in K positions only (tens on config 4's first thousands
// of pivots, up to RF_CAP handled here), and the engine already holds everything that describes the change:
//
//   B_new^-1 = E_K^-1 ... E_1^-1 B_old^-1                                  (solver.rs:1305-1319, the product form)
//   B_old^-1 [pos of core column t, row r in R_old]            =  C_old^-1[t, c(r)]
//   B_old^-1 [pos of the basic slack of row i, row r in R_old] = -D[i, J_old] . C_old^-1[:, c(r)]
//   B_old^-1 [ . , row r whose slack is basic]                 =  unit vector at that slack's position
//   C_new^-1 [t', c']  =  B_new^-1 [pos of new core column t', new core row R_new[c']]
//
// and the eta chain in closed form: X = X0 - E T with T = (I+G)^-1 X0[etaR, :].  So the inverse of the NEW core
// is a gather of the old one (plus <= K new rows, one sparse row-times-matrix product each, and unit columns) minus a rank-K
// product: O(k^2 K) throughput-bound work on all SMs instead of O(k) latency-bound steps + O(k^3).  It is the same
// arithmetic a longer eta file would do at every solve, done once; rounding accumulates like an eta file of that length,
// which is why every refresh is probed against the new core (k_rf_probe; above the tolerance it is redone as a true
// factorization) and why MLP_TUNE_LU_EVERY can bound the pivots between true factorizations.  C^-1 is an engine-internal
// quantity (the reference solves with L and U), so the rank-K product may contract a*b+c.
#pragma once

constexpr int RF_MAXK = 128;  // chunk of etas held in shared memory by k_rf_t; also the length of the device-map patch lists
constexpr int RF_CAP = 1024;  // etas folded by one refresh at most (rows of rf_W / rf_T / rf_Ep); a longer file takes the true factorization

// entry of B_old^-1: row descriptor rs (>= 0: old core column; < 0: row -1-rs of W) at basis position pos, column descriptor
// cs (>= 0: old core row; < 0: the unit vector at position -1-cs)
__device__ __forceinline__ double rf_x0(int rs, int pos, int cs, const double* __restrict__ Cinv, int64_t ld,
                                        const double* __restrict__ W, int64_t wld) {
  if (cs >= 0) return rs >= 0 ? Cinv[(int64_t)cs * ld + rs] : W[(int64_t)(-1 - rs) * wld + cs];
  return pos == -1 - cs ? 1.0 : 0.0;
}

// W[q, c] = -D[wrow[q], J_old] . C_old^-1[:, c]: the rows of B_old^-1 at positions that held a slack then and are needed now.
// D's rows come from the compact row-major copy of the old basic columns (a handful of entries each).
__global__ void __launch_bounds__(256) k_rf_w(const int64_t* __restrict__ dptr, const int32_t* __restrict__ didx,
                                               const double* __restrict__ dval, const int32_t* __restrict__ wrow, int k_old,
                                               const double* __restrict__ Cinv, int64_t ld, double* __restrict__ W, int64_t wld) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
  if (c >= k_old) return;
  const int i = wrow[q];
  double acc = 0.0;
  for (int64_t e = dptr[i]; e < dptr[i + 1]; ++e) acc -= dval[e] * Cinv[(int64_t)c * ld + didx[e]];
  W[(int64_t)q * wld + c] = acc;
}

// T[:, c'] = (I+G)^-1 X0[etaR, c'] for 32 columns c' per CTA (lower-triangular product, as k_mv_n<true>).  T is K x k_new,
// leading dimension RF_CAP.  The eta index j is walked in chunks of RF_MAXK held in shared memory; a chunk contributes to
// every row i at or below it, accumulated in T itself (each (i, c') is touched by one thread per chunk, chunks are
// separated by CTA barriers).  K <= RF_MAXK: one chunk, no accumulation through memory.
__global__ void __launch_bounds__(256) k_rf_t(const double* __restrict__ Ginv, int64_t Kld, int K, const int32_t* __restrict__ etasrc,
                                               const int32_t* __restrict__ etapos, const int32_t* __restrict__ colsrc, int k_new,
                                               const double* __restrict__ Cinv, int64_t ld, const double* __restrict__ W, int64_t wld,
                                               double* __restrict__ T) {
  pdl_wait();
  __shared__ double Us[RF_MAXK][33];
  const int c0 = blockIdx.x * 32;
  for (int jc = 0; jc < K; jc += RF_MAXK) {
    const int nj = min(RF_MAXK, K - jc);
    __syncthreads();  // the previous chunk's readers of Us are done (and its updates of T are visible)
    for (int q = threadIdx.x; q < nj * 32; q += 256) {
      const int j = q >> 5, cc = q & 31;
      Us[j][cc] = c0 + cc < k_new ? rf_x0(etasrc[jc + j], etapos[jc + j], colsrc[c0 + cc], Cinv, ld, W, wld) : 0.0;
    }
    __syncthreads();
    const int ni = K - jc;  // rows jc .. K-1 see this chunk
    for (int q = threadIdx.x; q < ni * 32; q += 256) {
      const int i = jc + q % ni, cc = q / ni;
      if (c0 + cc >= k_new) continue;
      const int jend = min(nj, i - jc + 1);  // j <= i
      double acc = 0.0;
      for (int j = 0; j < jend; ++j) acc += Ginv[(int64_t)(jc + j) * Kld + i] * Us[j][cc];
      double* dst = T + (int64_t)(c0 + cc) * RF_CAP + i;
      *dst = jc == 0 ? acc : *dst + acc;
    }
  }
}

// Cn[t', c'] = X0[pos of new core column t', c']  (column-major, leading dimension ld)
__global__ void __launch_bounds__(256) k_rf_x0(const int32_t* __restrict__ rowsrc, const int32_t* __restrict__ jposn,
                                                const int32_t* __restrict__ colsrc, int k_new, const double* __restrict__ Cinv,
                                                int64_t ld, const double* __restrict__ W, int64_t wld, double* __restrict__ Cn) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k_new) return;
  const int rs = rowsrc[t], pos = jposn[t];
  for (int c = blockIdx.y; c < k_new; c += gridDim.y) Cn[(int64_t)c * ld + t] = rf_x0(rs, pos, colsrc[c], Cinv, ld, W, wld);
}

// Ep[t', j] = E[pos of new core column t', j]  (k_new x K, leading dimension ld)
__global__ void __launch_bounds__(256) k_rf_ep(const double* __restrict__ E, int64_t mld, const int32_t* __restrict__ jposn, int k_new,
                                                double* __restrict__ Ep, int64_t ld) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (t < k_new) Ep[(int64_t)j * ld + t] = E[(int64_t)j * mld + jposn[t]];
}

// Accuracy probe of a refreshed inverse.  For RF_PROBE sampled columns c, x = C_new^-1[:, c] is checked against the new core
// itself (read from the compact row-major copy of the NEW basic columns; row Rp[i] of it is row i of the core):
//   out[1 + q] = bits of max_i |C x - e_c|_i          out[1 + RF_PROBE + q] = bits of max_i (|C||x| + |e_c|)_i
// and the host takes the largest quotient: a normwise backward error (a componentwise one is useless here — where the exact
// x_t is 0 a refresh leaves 1e-17 and a row with that single entry reports 100 %).  A fresh factorization sits at 1e-16..1e-12
// (MLP_REFACTOR_TRACE prints both); above the tolerance the refresh is redone as a true factorization.
// out[0] = entries of the core (for the estimate of LUFactors::nnz).
constexpr int RF_PROBE = 4;
__global__ void __launch_bounds__(256) k_rf_probe(const int64_t* __restrict__ dptr, const int32_t* __restrict__ didx,
                                                   const double* __restrict__ dval, const int32_t* __restrict__ Rp, int k,
                                                   const double* __restrict__ Cinv, int64_t ld, int c0, int cstep, int ncol,
                                                   unsigned long long* __restrict__ out) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double num[RF_PROBE], den[RF_PROBE];
#pragma unroll
  for (int q = 0; q < RF_PROBE; ++q) num[q] = den[q] = 0.0;
  unsigned len = 0;
  if (i < k) {
    const int r = Rp[i];
    len = (unsigned)(dptr[r + 1] - dptr[r]);
#pragma unroll
    for (int q = 0; q < RF_PROBE; ++q) {
      if (q >= ncol) break;
      const int c = (c0 + q * cstep) % k;
      const double* col = Cinv + (int64_t)c * ld;
      double acc = i == c ? -1.0 : 0.0, mag = i == c ? 1.0 : 0.0;
      for (int64_t e = dptr[r]; e < dptr[r + 1]; ++e) {
        const double t = dval[e] * col[didx[e]];
        acc += t;
        mag += fabs(t);
      }
      num[q] = acc == acc ? fabs(acc) : 1e300;  // a NaN counts as a failure
      den[q] = mag == mag ? mag : 0.0;
    }
  }
#pragma unroll
  for (int q = 0; q < RF_PROBE; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      num[q] = fmax(num[q], __shfl_down_sync(0xffffffffu, num[q], o));
      den[q] = fmax(den[q], __shfl_down_sync(0xffffffffu, den[q], o));
    }
  }
  len = __reduce_add_sync(0xffffffffu, len);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int q = 0; q < RF_PROBE; ++q) {
      if (num[q] > 0.0) atomicMax(out + 1 + q, (unsigned long long)__double_as_longlong(num[q]));
      if (den[q] > 0.0) atomicMax(out + 1 + RF_PROBE + q, (unsigned long long)__double_as_longlong(den[q]));
    }
    if (len) atomicAdd(out, (unsigned long long)len);
  }
}
