// Kernels of the pivot engine that do not depend on how the variables are sharded.  Included by engine.cu only.
#pragma once
// ------------------------------------------------------------------------------------------------ device helpers
// Programmatic dependent launch: every kernel of the pivot path is launched with programmaticStreamSerializationAllowed, so
// that its CTAs are scheduled while the previous kernel of the stream is still draining, and parks here until that kernel
// has completed and its writes are visible (griddepcontrol.wait = cudaGridDependencySynchronize).  First statement of every
// kernel: nothing before it touches memory, so the semantics are those of plain stream order.  Between the 25-40 short
// dependent kernels of a pivot this hides most of the launch gap (MLP_PDL=0 launches the ordinary way; the cooperative
// chain kernel is always launched the ordinary way).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULLMASK, v, o);
  return v;
}
// Deterministic block sum (result valid in thread 0). sm: >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sm) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = lane < nw ? sm[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}
struct KeyIdx {
  double key;
  long long idx;
};
// "better" orderings: max key then min idx / min key then min idx
__device__ __forceinline__ bool better_max(double k, long long i, double bk, long long bi) { return k > bk || (k == bk && i < bi); }
__device__ __forceinline__ KeyIdx warp_argmax(KeyIdx v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double k = __shfl_down_sync(FULLMASK, v.key, o);
    long long i = __shfl_down_sync(FULLMASK, v.idx, o);
    if (better_max(k, i, v.key, v.idx)) { v.key = k; v.idx = i; }
  }
  return v;
}
__device__ __forceinline__ KeyIdx block_argmax(KeyIdx v, double* smk, long long* smi) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_argmax(v);
  __syncthreads();
  if (lane == 0) { smk[wid] = v.key; smi[wid] = v.idx; }
  __syncthreads();
  KeyIdx r{-INFINITY, LLONG_MAX};
  if (wid == 0) {
    if (lane < nw) { r.key = smk[lane]; r.idx = smi[lane]; }
    r = warp_argmax(r);
  }
  return r;
}
// Arg-max that also counts how contested the winner is (pass 2 of the two Harris ratio tests, solver.rs:804-823 and
// 982-1002, where the reference resolves an exact tie in |coeff| by list order): ex = candidates whose key equals the
// winner's exactly, nr = candidates within NEAR_TIE relative of it; both include the winner.  The merge is exact for ex;
// nr may over-count slightly (a candidate within NEAR_TIE of a runner-up that is itself within NEAR_TIE of the winner).
constexpr double NEAR_TIE = 1e-9;
struct KeyIdxC {
  double key;
  long long idx;
  int ex, nr;
};
__device__ __forceinline__ void kic_merge(KeyIdxC& a, double bk, long long bi, int bex, int bnr) {
  if (bi == LLONG_MAX) return;
  if (a.idx == LLONG_MAX) { a.key = bk; a.idx = bi; a.ex = bex; a.nr = bnr; return; }
  if (better_max(bk, bi, a.key, a.idx)) {
    const int ex = bex + (a.key == bk ? a.ex : 0), nr = bnr + (a.key >= bk * (1.0 - NEAR_TIE) ? a.nr : 0);
    a.key = bk; a.idx = bi; a.ex = ex; a.nr = nr;
  } else {
    a.ex += (bk == a.key ? bex : 0);
    a.nr += (bk >= a.key * (1.0 - NEAR_TIE) ? bnr : 0);
  }
}
__device__ __forceinline__ KeyIdxC warp_argmax_c(KeyIdxC v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double k = __shfl_down_sync(FULLMASK, v.key, o);
    const long long i = __shfl_down_sync(FULLMASK, v.idx, o);
    const int ex = __shfl_down_sync(FULLMASK, v.ex, o), nr = __shfl_down_sync(FULLMASK, v.nr, o);
    kic_merge(v, k, i, ex, nr);
  }
  return v;
}
// smc: >= 32 long long (packed counts)
__device__ __forceinline__ KeyIdxC block_argmax_c(KeyIdxC v, double* smk, long long* smi, long long* smc) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_argmax_c(v);
  __syncthreads();
  if (lane == 0) { smk[wid] = v.key; smi[wid] = v.idx; smc[wid] = ((long long)v.ex << 32) | (unsigned)v.nr; }
  __syncthreads();
  KeyIdxC r{-INFINITY, LLONG_MAX, 0, 0};
  if (wid == 0) {
    if (lane < nw) { r.key = smk[lane]; r.idx = smi[lane]; r.ex = (int)(smc[lane] >> 32); r.nr = (int)(smc[lane] & 0xffffffffLL); }
    r = warp_argmax_c(r);
  }
  return r;
}
// tie counts (excluding the winner itself) packed into the low 32 bits of a candidate's `tie` word, saturating at 65535
__device__ __forceinline__ long long pack_ties(int ex, int nr) {
  const int e = min(max(ex - 1, 0), 65535), n = min(max(nr - 1, 0), 65535);
  return ((long long)n << 16) | (long long)e;
}
constexpr int RED_CNT_OFF = 2048;  // per-CTA packed counts live in red_i[RED_CNT_OFF + block] (ratio grids are <= 1024 CTAs)

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(FULLMASK, v, o));
  return v;
}
__device__ __forceinline__ double block_min(double v, double* sm) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double r = INFINITY;
  if (wid == 0) {
    r = lane < nw ? sm[lane] : INFINITY;
    r = warp_min(r);
  }
  return r;
}
// Grid-level "last block finishes" rendezvous. Returns true in every thread of the last-arriving block.
__device__ __forceinline__ bool last_block(unsigned* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// ------------------------------------------------------------------------------------------------ compaction
// ScatteredVec::to_sparse_vec (sparse.rs:115-121) for a device work vector: ordered list of the non-zero
// entries (ascending index), their count, and the sum of squares (SparseVec::sq_norm, sparse.rs:32-34).
// Pass 1: every CTA counts the non-zeros and sums the squares of its 1024-entry segment; the last CTA to
// finish adds the per-segment results in segment order (deterministic).  Pass 2 (only when the list is
// needed): each CTA derives its output offset from the segment counts and writes its entries in order.
constexpr int CP_SEG = 1024;
// mask != nullptr: the count is the number of set mask bytes instead (the STRUCTURAL size of col_coeffs, see k_touch_mark);
// the sum of squares is always the numeric one.
__global__ void __launch_bounds__(CP_SEG) k_compact_count(const double* __restrict__ x, int m, int32_t* __restrict__ seg_cnt,
                                                           double* __restrict__ seg_ss, unsigned* counter,
                                                           int32_t* __restrict__ count, double* __restrict__ sumsq,
                                                           const uint8_t* __restrict__ mask) {
  pdl_wait();
  __shared__ double sm[32];
  __shared__ int smi[32];
  const int i = blockIdx.x * CP_SEG + threadIdx.x;
  const double v = i < m ? x[i] : 0.0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(FULLMASK, mask ? (i < m && mask[i] != 0) : (v != 0.0));
  if (lane == 0) smi[wid] = __popc(bal);
  const double ss = block_sum(v * v, sm);  // has the barriers that publish smi
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w2 = 0; w2 < 32; ++w2) c += smi[w2];
    seg_cnt[blockIdx.x] = c;
    seg_ss[blockIdx.x] = ss;
  }
  if (!last_block(counter)) return;
  if (threadIdx.x < 32) {  // fixed order: lane-strided partial sums, then the shuffle tree
    int c = 0;
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) { c += __ldcg(seg_cnt + b); t += __ldcg(seg_ss + b); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULLMASK, c, o);
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      *count = c;
      *sumsq = t;
      *counter = 0;
    }
  }
}
__global__ void __launch_bounds__(CP_SEG) k_compact_write(const double* __restrict__ x, int m, const int32_t* __restrict__ seg_cnt,
                                                           int32_t* __restrict__ idx, double* __restrict__ val) {
  pdl_wait();
  __shared__ int warp_cnt[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (wid == 0) {  // offset of this segment = sum of the counts of the segments before it
    int acc = 0;
    for (int b = lane; b < (int)blockIdx.x; b += 32) acc += seg_cnt[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULLMASK, acc, o);
    if (lane == 0) base = acc;
  }
  const int i = blockIdx.x * CP_SEG + threadIdx.x;
  const double v = i < m ? x[i] : 0.0;
  const bool nz = v != 0.0;
  const unsigned bal = __ballot_sync(FULLMASK, nz);
  if (lane == 0) warp_cnt[wid] = __popc(bal);
  __syncthreads();
  if (nz) {
    int off = base;
    for (int w2 = 0; w2 < wid; ++w2) off += warp_cnt[w2];
    const int p = off + __popc(bal & ((1u << lane) - 1u));
    idx[p] = i;
    val[p] = v;
  }
}

// Structural size of col_coeffs.  The reference's FTRAN keeps a position in the result's non-zero list once it has been
// TOUCHED, whatever its value: BasisSolver::solve (solver.rs:1305-1319) applies every eta to all of that eta's stored
// positions (`*rhs.get_mut(r) -= coeff * val`, sparse.rs:75-80 marks r) even when coeff is 0, and push_eta_matrix
// (1274-1284) stores every listed position of col_coeffs, zeros included.  So the stored size of eta K is
// |pattern(LU solve of a_q) U stored positions of eta K-1|, and THAT is what the refactor rule (1096-1097) adds up — on
// sparse LPs several times the numeric count.  `touched` holds the stored positions of the newest eta (cleared at a
// refactorization), `touched_new` = touched U nz(alpha0) is written by the FTRAN of an entering column (alpha0: the result
// of the LU part, before the etas) and becomes `touched` when that column is pushed (k_pivot_rows).
__global__ void k_touch_mark(const double* __restrict__ alpha0, const uint8_t* __restrict__ touched, int m,
                             uint8_t* __restrict__ touched_new) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) touched_new[i] = (uint8_t)((alpha0[i] != 0.0) | (touched[i] != 0));
}

// ------------------------------------------------------------------------------------------------ explicit inverses
// The two small triangular systems of every FTRAN/BTRAN — the k x k core of the basis (L U = P C) and the K x K eta
// coupling matrix I+G — are latency-bound as substitutions (k/32 dependent steps on one CTA).  Their INVERSES are kept
// instead (C^-1 rebuilt at every refactorization, (I+G)^-1 extended by one row per eta), so that each solve becomes
// one grid-parallel matrix-vector product.
//   k_mv_n: y[i]        = sum_j M[i + j ld] x[gidx ? gidx[j] : j]     (TRI: j <= i)   CTA = 32 rows x 8 column groups
//   k_mv_t: y[sidx?[j]] = sum_i M[i + j ld] x[i]                      (TRI: i >= j)   one warp per column
template <bool TRI>
__global__ void __launch_bounds__(256) k_mv_n(const double* __restrict__ M, int64_t ld, int n, const double* __restrict__ xin,
                                              const int32_t* __restrict__ gidx, double* __restrict__ y) {
  pdl_wait();
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double acc = 0.0;
  if (i < n) {
    const int jend = TRI ? i + 1 : n;
    const double* p = M + i;
    int j = g;
    for (; j + 24 < jend; j += 32) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = p[(int64_t)(j + 8 * u) * ld];
        b[u] = xin[gidx ? gidx[j + 8 * u] : j + 8 * u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += a[u] * b[u];
    }
    for (; j < jend; j += 8) acc += p[(int64_t)j * ld] * xin[gidx ? gidx[j] : j];
  }
  part[g][lane] = acc;
  __syncthreads();
  if (g == 0 && i < n) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][lane];
    y[i] = t;
  }
}
template <bool TRI>
__global__ void __launch_bounds__(256) k_mv_t(const double* __restrict__ M, int64_t ld, int n, const double* __restrict__ xin,
                                              const int32_t* __restrict__ sidx, double* __restrict__ y) {
  pdl_wait();
  const int lane = threadIdx.x & 31, j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= n) return;
  const double* col = M + (int64_t)j * ld;
  double acc = 0.0;
  for (int i = (TRI ? j : 0) + lane; i < n; i += 32) acc += col[i] * xin[i];
  acc = warp_sum(acc);
  if (lane == 0) y[sidx ? sidx[j] : j] = acc;
}
// Column j of (L U)^-1 by substitution on e_j (the reference's column-oriented solves, lu.rs:450-463), one CTA per
// column; the work vector lives in shared memory when it fits (use_smem), else in the output column itself.
__global__ void __launch_bounds__(256) k_core_inverse(const double* __restrict__ LU, int64_t ld, int k, double* __restrict__ Cinv,
                                                       const int* __restrict__ flags, int use_smem) {
  pdl_wait();
  extern __shared__ double smem_v[];
  if (flags[1]) return;
  const int j = blockIdx.x, tid = threadIdx.x;
  double* out = Cinv + (int64_t)j * ld;
  double* v = use_smem ? smem_v : out;
  for (int i = tid; i < k; i += 256) v[i] = (i == j) ? 1.0 : 0.0;
  __syncthreads();
  for (int t = j; t < k - 1; ++t) {  // L y = e_j (unit diagonal; y_i = 0 for i < j)
    const double vt = v[t];
    if (vt != 0.0) {
      const double* c = LU + (int64_t)t * ld;
      for (int i = t + 1 + tid; i < k; i += 256) v[i] -= c[i] * vt;
    }
    __syncthreads();
  }
  for (int t = k - 1; t >= 0; --t) {  // U x = y
    const double* c = LU + (int64_t)t * ld;
    const double vt = v[t] / c[t];
    if (!use_smem) __syncthreads();  // out aliases v: every thread must have read v[t] before thread 0 overwrites it
    if (vt != 0.0)
      for (int i = tid; i < t; i += 256) v[i] -= c[i] * vt;
    if (tid == 0) out[t] = vt;
    __syncthreads();
  }
}
// The same substitution for k <= 256 RPT with every thread owning FIXED rows (tid, tid + 256, ...) so that the entries of
// the NEXT factor column can be fetched one step ahead: in k_core_inverse each of the ~2k sequential steps waits for a
// global load of its column (L2 latency, ~650 cycles per step: 138 us at k = 207); here that latency is off the
// critical path and a step costs one shared-memory round trip and a barrier.  Same operations in the same order.
template <int RPT>
__global__ void __launch_bounds__(256) k_core_inverse_pf(const double* __restrict__ LU, int64_t ld, int k, double* __restrict__ Cinv,
                                                          const int* __restrict__ flags) {
  pdl_wait();
  extern __shared__ double smem_v[];
  if (flags[1]) return;
  const int j = blockIdx.x, tid = threadIdx.x;
  double* out = Cinv + (int64_t)j * ld;
  double* v = smem_v;
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int i = tid + 256 * q;
    if (i < k) v[i] = (i == j) ? 1.0 : 0.0;
  }
  double nxt[RPT];
  // L y = e_j (unit diagonal; y_i = 0 for i < j)
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int i = tid + 256 * q;
    nxt[q] = (i < k && i > j) ? LU[(int64_t)j * ld + i] : 0.0;
  }
  __syncthreads();
  for (int t = j; t < k - 1; ++t) {
    double cur[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      cur[q] = nxt[q];
      const int i = tid + 256 * q;
      nxt[q] = (t + 2 < k && i < k && i > t + 1) ? LU[(int64_t)(t + 1) * ld + i] : 0.0;  // column of the next step
    }
    const double vt = v[t];
    if (vt != 0.0) {
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int i = tid + 256 * q;
        if (i < k && i > t) v[i] -= cur[q] * vt;
      }
    }
    __syncthreads();
  }
  // U x = y
  double dnx = LU[(int64_t)(k - 1) * ld + (k - 1)];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int i = tid + 256 * q;
    nxt[q] = (i < k - 1) ? LU[(int64_t)(k - 1) * ld + i] : 0.0;
  }
  for (int t = k - 1; t >= 0; --t) {
    double cur[RPT];
    const double diag = dnx;
    if (t > 0) dnx = LU[(int64_t)(t - 1) * ld + (t - 1)];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      cur[q] = nxt[q];
      const int i = tid + 256 * q;
      nxt[q] = (t > 0 && i < t - 1) ? LU[(int64_t)(t - 1) * ld + i] : 0.0;
    }
    const double vt = v[t] / diag;
    if (vt != 0.0) {
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int i = tid + 256 * q;
        if (i < t) v[i] -= cur[q] * vt;
      }
    }
    if (tid == 0) out[t] = vt;
    __syncthreads();
  }
}
// New row K of (I+G)^-1 when eta K with coupling row g (g[i] = E_i[r_K], i < K) is appended:
// [[T,0],[g^T,1]]^-1 = [[T^-1,0],[-g^T T^-1,1]].  One warp per column.
__global__ void __launch_bounds__(256) k_eta_inv_row(const double* __restrict__ g, double* __restrict__ Ginv, int64_t ld, int K) {
  pdl_wait();
  const int lane = threadIdx.x & 31, j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j > K) return;
  double* col = Ginv + (int64_t)j * ld;
  if (j == K) { if (lane == 0) col[K] = 1.0; return; }
  double acc = 0.0;
  for (int i = j + lane; i < K; i += 32) acc += g[i] * col[i];
  acc = warp_sum(acc);
  if (lane == 0) col[K] = -acc;
}

// y[i] = base[i] - sum_j M[i + j*ld] * t[j]   (column-major M: rows x cols; thread per row)
// Used for: FTRAN eta application rhs -= E t (solver.rs:1310-1316 in closed form) and the slack rows of
// the basis solve alpha_S = a_S - D1 x.
__global__ void __launch_bounds__(256) k_gemv_n_sub(const double* __restrict__ M, int64_t ld, int rows, int cols,
                                                     const double* __restrict__ t, double* __restrict__ y) {
  pdl_wait();
  __shared__ double ts[512];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = i < rows ? y[i] : 0.0;
  for (int j0 = 0; j0 < cols; j0 += 512) {
    const int nj = min(512, cols - j0);
    __syncthreads();
    for (int q = threadIdx.x; q < nj; q += blockDim.x) ts[q] = t[j0 + q];
    __syncthreads();
    if (i < rows) {
      const double* p = M + (int64_t)j0 * ld + i;
      int j = 0;
      for (; j + 16 <= nj; j += 16) {  // 16 loads in flight, subtraction still in column order
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = p[(int64_t)(j + u) * ld];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc -= ts[j + u] * v[u];
      }
      for (; j + 4 <= nj; j += 4) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p[(int64_t)(j + u) * ld];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc -= ts[j + u] * v[u];
      }
      for (; j < nj; ++j) acc -= ts[j] * p[(int64_t)j * ld];
    }
  }
  if (i < rows) y[i] = acc;
}

// The eta part of an FTRAN in ONE launch for short eta files (K <= FE_MAXK): every CTA first forms the K scalars
// t = (I+G)^-1 in[etaR] itself — K^2/2 multiply-adds, nothing next to a launch — and then applies out[i] = in[i] - sum_j t_j E[i,j]
// to its rows.  Same operations in the same order as k_mv_n<true> followed by k_gemv_n_sub (eight column groups per row of the
// triangular product, added in group order; columns subtracted in ascending order): bit-identical results.  `in` and `out`
// are different buffers (every CTA reads in[etaR[*]], which other CTAs would be overwriting in place).
constexpr int FE_MAXK = 128;
__global__ void __launch_bounds__(256) k_eta_apply(const double* __restrict__ E, int64_t ld, int rows, int K,
                                                    const double* __restrict__ Ginv, int64_t Kld, const int32_t* __restrict__ etaR,
                                                    const double* __restrict__ in, double* __restrict__ out) {
  pdl_wait();
  __shared__ double xs[FE_MAXK];
  __shared__ double ts[FE_MAXK];
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < K; j += 256) xs[j] = in[etaR[j]];
  __syncthreads();
  for (int i0 = 0; i0 < K; i0 += 32) {
    const int i = i0 + lane;
    double acc = 0.0;
    if (i < K) {
      const double* p = Ginv + i;
      for (int j = g; j <= i; j += 8) acc += p[(int64_t)j * Kld] * xs[j];
    }
    part[g][lane] = acc;
    __syncthreads();
    if (g == 0 && i < K) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += part[q][lane];
      ts[i] = t;
    }
    __syncthreads();
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double acc = in[i];
  const double* p = E + i;
  int j = 0;
  for (; j + 16 <= K; j += 16) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = p[(int64_t)(j + u) * ld];
#pragma unroll
    for (int u = 0; u < 16; ++u) acc -= ts[j + u] * v[u];
  }
  for (; j + 4 <= K; j += 4) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = p[(int64_t)(j + u) * ld];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc -= ts[j + u] * v[u];
  }
  for (; j < K; ++j) acc -= ts[j] * p[(int64_t)j * ld];
  out[i] = acc;
}
// BTRAN of a unit vector e_r through a short eta file, first half in ONE launch: c = e_r (cnt entries) and
// s = (I+G)^-T u with u = row r of E read in place (k_unit_and_gather + k_mv_t<true>; one warp per column, same order).
__global__ void __launch_bounds__(256) k_unit_eta_t(double* __restrict__ c, int64_t cnt, int64_t at, const double* __restrict__ E,
                                                     int64_t lde, const double* __restrict__ Ginv, int64_t Kld, int K,
                                                     double* __restrict__ s, const long long* __restrict__ at_dev) {
  pdl_wait();
  if (at_dev) at = *at_dev;  // the row left in device memory by k_select_row_dual (< 0: none)
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) c[t] = (t == at) ? 1.0 : 0.0;
  const int lane = threadIdx.x & 31, j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= K) return;
  const double* col = Ginv + (int64_t)j * Kld;
  double acc = 0.0;
  if (at >= 0)
    for (int i = j + lane; i < K; i += 32) acc += col[i] * E[(int64_t)i * lde + at];
  acc = warp_sum(acc);
  if (lane == 0) s[j] = acc;
}
// Eta push in ONE launch for short eta files: the bookkeeping of k_eta_grow and the new row K of (I+G)^-1 (k_eta_inv_row) with
// the coupling row g[i] = E_i[r_K] read in place.  Same order of operations: bit-identical.
__global__ void __launch_bounds__(256) k_eta_push(const double* __restrict__ E, int64_t lde, int K, int rK, double* __restrict__ Ginv,
                                                   int64_t ld, int32_t* etaR, int32_t* etaPrev, int32_t* etaHead, int32_t* etaLast, int prev) {
  pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    etaR[K] = rK;
    etaPrev[K] = prev;
    etaHead[K] = 1;
    etaLast[rK] = K;
    if (prev >= 0) etaHead[prev] = 0;
  }
  const int lane = threadIdx.x & 31, j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j > K) return;
  double* col = Ginv + (int64_t)j * ld;
  if (j == K) { if (lane == 0) col[K] = 1.0; return; }
  double acc = 0.0;
  for (int i = j + lane; i < K; i += 32) acc += E[(int64_t)i * lde + rK] * col[i];
  acc = warp_sum(acc);
  if (lane == 0) col[K] = -acc;
}

// Column-group split of the same tall-skinny products for wide basis blocks: with one thread per row a 50k-row pass
// has only ~340 threads per SM, too few loads in flight to saturate HBM once cols grows into the hundreds.  Grid
// (row tiles, G): CTA (x, g) accumulates column group g into part[g][i]; the finishing kernels add the G partials in
// group order.   col(j) = slots ? slots[j] : j;   rows with rowmask[i] < 0 are skipped.
__global__ void __launch_bounds__(256) k_tall_part(const double* __restrict__ M, int64_t ld, int rows, int cols,
                                                    const double* __restrict__ t, const int32_t* __restrict__ slots,
                                                    const int32_t* __restrict__ rowmask, double* __restrict__ part, int64_t pld) {
  pdl_wait();
  __shared__ double ts[512];
  __shared__ int32_t sl[512];
  const int G = gridDim.y, g = blockIdx.y;
  const int cg = (cols + G - 1) / G;
  const int c0 = g * cg, c1 = min(cols, c0 + cg);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < rows && (!rowmask || rowmask[i] >= 0);
  double acc = 0.0;
  for (int j0 = c0; j0 < c1; j0 += 512) {
    const int nj = min(512, c1 - j0);
    __syncthreads();
    for (int q = threadIdx.x; q < nj; q += blockDim.x) { ts[q] = t[j0 + q]; sl[q] = slots ? slots[j0 + q] : j0 + q; }
    __syncthreads();
    if (act) {
      const double* p = M + i;
      int j = 0;
      for (; j + 16 <= nj; j += 16) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = p[(int64_t)sl[j + u] * ld];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += ts[j + u] * v[u];
      }
      for (; j < nj; ++j) acc += ts[j] * p[(int64_t)sl[j] * ld];
    }
  }
  if (i < rows) part[(int64_t)g * pld + i] = acc;
}
// y[i] -= sum_g part[g][i]
__global__ void k_sub_parts(const double* __restrict__ part, int G, int64_t pld, int rows, double* __restrict__ y) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double tsum = 0.0;
  for (int g = 0; g < G; ++g) tsum += part[(int64_t)g * pld + i];
  y[i] -= tsum;
}

// out[j] = base[idx[j]] - sum_i M[i + j*ld] * x[i]   (negate) or the plain dot products.
// Grid (cols, S): CTA (j, s) reduces row slice s of column j; k_gemv_t_fin adds the S partials in order.
// Used for BTRAN: u = E^T rhs (solver.rs:1326-1330) and the right-hand side of the core solve.
constexpr int GT_MAXSPLIT = 16;
__global__ void __launch_bounds__(256) k_gemv_t_part(const double* __restrict__ M, int64_t ld, int rows, int cols,
                                                      const double* __restrict__ x, double* __restrict__ part) {
  pdl_wait();
  __shared__ double sm[32];
  const int j = blockIdx.x, S = gridDim.y, sidx = blockIdx.y;
  const int L = (rows + S - 1) / S;
  const int r0 = sidx * L, r1 = min(rows, r0 + L);
  const double* p = M + (int64_t)j * ld;
  double acc = 0.0;
  int i = r0 + threadIdx.x;
  const int st = blockDim.x;
  for (; i + 7 * st < r1; i += 8 * st) {  // 8 row pairs in flight per thread, added in row order
    double a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = p[i + u * st]; b[u] = x[i + u * st]; }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += a[u] * b[u];
  }
  for (; i < r1; i += st) acc += p[i] * x[i];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) part[(int64_t)sidx * cols + j] = tot;
}
__global__ void k_gemv_t_fin(const double* __restrict__ part, int S, int cols, const double* __restrict__ base,
                             const int32_t* __restrict__ base_idx, double* __restrict__ out, int negate) {
  pdl_wait();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  double tot = 0.0;
  for (int q = 0; q < S; ++q) tot += part[(int64_t)q * cols + j];
  const double b = base ? base[base_idx ? base_idx[j] : j] : 0.0;
  out[j] = negate ? b - tot : tot;
}

__global__ void k_gather_idx(const double* __restrict__ src, const int32_t* __restrict__ idx, int cnt, double* __restrict__ dst) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) dst[t] = src[idx[t]];
}
__global__ void k_scatter_idx(const double* __restrict__ src, const int32_t* __restrict__ idx, int cnt, double* __restrict__ dst) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) dst[idx[t]] = src[t];
}
// strided gather of one row of a column-major matrix: dst[j] = M[row + j*ld]
__global__ void k_gather_row(const double* __restrict__ M, int64_t ld, int row, int cnt, double* __restrict__ dst) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) dst[t] = M[(int64_t)t * ld + row];
}
// c = e_at (cnt entries) and dst[j] = M[at + j ld] for j < ncols, in one launch (BTRAN of a unit vector with an eta file)
// at_dev (optional): the row is read from device memory (left there by k_select_row_dual; < 0: no row — c = 0, nothing gathered)
__global__ void k_unit_and_gather(double* __restrict__ c, int64_t cnt, int64_t at, const double* __restrict__ M, int64_t ld,
                                  int ncols, double* __restrict__ dst, const long long* __restrict__ at_dev) {
  pdl_wait();
  if (at_dev) at = *at_dev;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) c[t] = (t == at) ? 1.0 : 0.0;
  if (t < ncols) dst[t] = at >= 0 ? M[t * ld + at] : 0.0;
}
__global__ void k_fill(double* p, int64_t cnt, double v) {
  pdl_wait();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) p[t] = v;
}
__global__ void k_set_unit(double* p, int64_t cnt, int64_t at) {
  pdl_wait();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) p[t] = (t == at) ? 1.0 : 0.0;
}
// BTRAN through the eta file, last step (solver.rs:1331-1332): rhs[r_leaving(idx)] -= coeff(idx), idx = K-1..0.
// Several etas may share a leaving row; thread j owns the chain headed by the LAST eta of a row and walks it in
// the reference's order (descending idx), so the subtraction order is the reference's.
__global__ void k_eta_scatter(const double* __restrict__ s, const int32_t* __restrict__ etaR,
                              const int32_t* __restrict__ etaPrev, const int32_t* __restrict__ etaHead, int K,
                              double* __restrict__ rhs) {
  pdl_wait();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K || !etaHead[j]) return;
  double v = rhs[etaR[j]];
  for (int q = j; q >= 0; q = etaPrev[q]) v -= s[q];
  rhs[etaR[j]] = v;
}

// BTRAN head: rho_i = c[pos of slack i] on covered rows (U^T solve over the identity block); cov copy with zeros elsewhere
__global__ void k_btran_start(const double* __restrict__ c, const int32_t* __restrict__ rowcover, int m,
                              double* __restrict__ out, double* __restrict__ cov) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int p = rowcover[i];
  const double v = p >= 0 ? c[p] : 0.0;
  out[i] = v;
  cov[i] = v;
}

// ------------------------------------------------------------------------------------------------ dense LU of the core
// lu_factorize (lu.rs:118-304) specialised to B = [D | E_S]: the unit columns come first in order_simple
// (ordering.rs:4-21) and pivot on their own rows; what remains is the k x k core C = D[R,:] whose columns
// are taken in basis-position order and whose pivots follow the reference's threshold rule:
// among rows with |x| >= 0.1 max|x| (lu.rs:224) — all have the same original-row count, lu.rs:225-229 —
// the first in list order, i.e. the lowest original row index.
//
// Right-looking, blocked in panels of nb <= 32 columns, three launches per panel:
//   k_lu_panel     one CTA factorizes the (k-j0) x nb panel held in shared memory (pivot search, row swap, scaling and
//                  the rank-1 updates inside the panel), records the panel's row permutation;
//   k_lu_swap_solve  one warp per column outside the panel applies that permutation and, right of the panel, solves
//                  U12 = L11^-1 A12 in registers;
//   k_lu_trailing  A22 -= L21 U12, thread per row with its L21 row in registers, 16 columns per CTA.
// Every element sees the same operations in the same order as column-by-column elimination (t ascending, separate
// multiply and subtract), so the factors are bit-identical to an unblocked factorization.
constexpr int LU_NB = 32;
// Warp reductions of the pivot search through the integer reduce unit (redux.sync: one instruction instead of a
// five-step shuffle tree per 32-bit word); every lane gets the result.
// max of non-negative doubles: for x >= 0 the IEEE bit pattern orders like the value.
__device__ __forceinline__ double warp_max_nonneg(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
  const unsigned mh = __reduce_max_sync(FULLMASK, hi);
  const unsigned ml = __reduce_max_sync(FULLMASK, hi == mh ? lo : 0u);
  return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}
// arg-min of distinct non-negative keys (UINT_MAX: none): returns the payload of the lane holding the smallest key
__device__ __forceinline__ int warp_argmin_u32(unsigned key, int payload, unsigned* min_out) {
  const unsigned mk = __reduce_min_sync(FULLMASK, key);
  const unsigned who = __ballot_sync(FULLMASK, key == mk);
  *min_out = mk;
  return __shfl_sync(FULLMASK, payload, __ffs(who) - 1);
}
// lexicographic arg-min of (hi, lo) key pairs with distinct lo (UINT_MAX, UINT_MAX: none)
__device__ __forceinline__ int warp_argmin_u32x2(unsigned hi, unsigned lo, int payload, unsigned* hi_out, unsigned* lo_out) {
  const unsigned mh = __reduce_min_sync(FULLMASK, hi);
  const unsigned ml = __reduce_min_sync(FULLMASK, hi == mh ? lo : 0xffffffffu);
  const unsigned who = __ballot_sync(FULLMASK, hi == mh && lo == ml);
  *hi_out = mh;
  *lo_out = ml;
  return __shfl_sync(FULLMASK, payload, __ffs(who) - 1);
}
// The panel factorization is column-sequential — per column: pivot search (two reductions: max |x|, then the lowest
// original row among the rows within 0.1 of it), row swap, scaling and rank-1 update of the rest of the panel — so it is
// bound by the LATENCY of one column step, not by throughput.  Per column: 4 CTA barriers (9 in the first version), the
// two reductions go through redux.sync, and the update is mapped one ROW per thread (thread r divides its entry of
// column c by the pivot, lu.rs:261, stores it and walks the remaining <= 31 columns of its row four at a time with all
// loads issued ahead of the stores): no integer div/mod per element as with a flat (row, column) index, no element is
// touched by two threads, stride-1 shared-memory accesses, the pivot row is a broadcast.  The launch uses as many
// threads as the panel has rows (<= 1024), which keeps the barriers cheap for small cores.
// Measured per 32-column panel at ~200 rows: 91 us (flat index, 9 barriers) -> 81 us (row per thread, shuffle trees)
// -> see profiles/ for this form.  The arithmetic and the pivot rule are unchanged: factors are bit-identical.
// Rcnt (sparse storage; nullptr for a dense A, where every row has the same count): number of entries of each core row in
// the basis matrix, the reference's orig_row2elt_count (lu.rs:139-144) — among the rows that pass the threshold the one
// with the FEWEST entries is taken (lu.rs:219-231); equal counts go to the lowest original row.
__global__ void __launch_bounds__(1024) k_lu_panel(double* __restrict__ C, int64_t ld, int k, int j0, int nb,
                                                    int32_t* __restrict__ Rp, int32_t* __restrict__ Rcnt, int* __restrict__ flags,
                                                    int32_t* __restrict__ aff_pos, int32_t* __restrict__ aff_src,
                                                    int32_t* __restrict__ aff_cnt, int32_t* __restrict__ perm_glob, int use_smem) {
  pdl_wait();
  extern __shared__ __align__(16) unsigned char lu_smem[];
  __shared__ double redk[32];
  __shared__ unsigned redrp[32];
  __shared__ unsigned redrc[32];
  __shared__ int redr[32];
  __shared__ int s_cnt;
  const int rows = k - j0, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
  const int T = blockDim.x;
  double* P;
  int64_t pld;
  int32_t *rp, *perm, *rc = nullptr;
  if (use_smem) {
    P = reinterpret_cast<double*>(lu_smem);
    pld = rows;
    rp = reinterpret_cast<int32_t*>(P + (size_t)rows * nb);
    perm = rp + rows;
    if (Rcnt) {
      rc = perm + rows;
      for (int r = tid; r < rows; r += T) rc[r] = Rcnt[j0 + r];
    }
    for (int r = tid; r < rows; r += T) {  // 8 global loads in flight per thread (a plain loop serialises them: ~12 us per panel)
      const double* src = C + (int64_t)j0 * ld + j0 + r;
      for (int c0 = 0; c0 < nb; c0 += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = c0 + u < nb ? src[(int64_t)(c0 + u) * ld] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u < nb) P[(size_t)(c0 + u) * pld + r] = v[u];
      }
      rp[r] = Rp[j0 + r];
    }
  } else {
    P = C + (int64_t)j0 * ld + j0;
    pld = ld;
    rp = Rp + j0;
    perm = perm_glob;
    if (Rcnt) rc = Rcnt + j0;
  }
  for (int r = tid; r < rows; r += T) perm[r] = r;
  if (tid == 0) s_cnt = 0;
  bool stop = flags[1] != 0;  // uniform: flags[1] only changes inside this kernel, by a uniform decision
  __syncthreads();
  for (int c = 0; c < nb && !stop; ++c) {
    double* col = P + (size_t)c * pld;
    // pivot search 1: max |x| from the diagonal down (lu.rs:194-206); a NaN never wins (fmax semantics)
    double mx = 0.0;
    for (int r = c + tid; r < rows; r += T) {
      const double x = fabs(col[r]);
      if (x > mx) mx = x;
    }
    mx = warp_max_nonneg(mx);
    if (lane == 0) redk[wid] = mx;
    __syncthreads();
    const double max_abs = warp_max_nonneg(lane < nw ? redk[lane] : 0.0);
    if (!(max_abs >= 1e-8) || isinf(max_abs)) {  // lu.rs:207-211; uniform across the CTA
      if (tid == 0) flags[1] = 1;
      stop = true;
      break;
    }
    // pivot search 2: among the eligible rows the one with the fewest entries in the basis matrix, then the lowest
    // original row (original rows are distinct)
    unsigned bc = 0xffffffffu, bk = 0xffffffffu;
    int br = -1;
    const double thr = 0.1 * max_abs;
    for (int r = c + tid; r < rows; r += T)
      if (fabs(col[r]) >= thr) {
        const unsigned cnt = rc ? (unsigned)rc[r] : 0u, key = (unsigned)rp[r];
        if (cnt < bc || (cnt == bc && key < bk)) { bc = cnt; bk = key; br = r; }
      }
    unsigned wc, wk;
    const int wr = warp_argmin_u32x2(bc, bk, br, &wc, &wk);
    if (lane == 0) { redrc[wid] = wc; redrp[wid] = wk; redr[wid] = wr; }
    __syncthreads();
    unsigned d0, d1;
    const int p = warp_argmin_u32x2(lane < nw ? redrc[lane] : 0xffffffffu, lane < nw ? redrp[lane] : 0xffffffffu,
                                    lane < nw ? redr[lane] : -1, &d0, &d1);
    if (p != c) {
      if (tid < nb) {
        double* q = P + (size_t)tid * pld;
        const double a = q[c], bb = q[p];
        q[c] = bb;
        q[p] = a;
      } else if (tid == nb) {
        const int a = rp[c]; rp[c] = rp[p]; rp[p] = a;
        const int bb = perm[c]; perm[c] = perm[p]; perm[p] = bb;
        if (rc) { const int cc2 = rc[c]; rc[c] = rc[p]; rc[p] = cc2; }
      }
    }
    __syncthreads();
    // scaling and rank-1 update, one row per thread
    const double pv = col[c];
    for (int r = c + 1 + tid; r < rows; r += T) {
      const double l = col[r] / pv;  // lu.rs:261
      col[r] = l;
      int cc = c + 1;
      for (; cc + 4 <= nb; cc += 4) {
        double* q0 = P + (size_t)cc * pld;
        double* q1 = q0 + pld;
        double* q2 = q1 + pld;
        double* q3 = q2 + pld;
        const double u0 = q0[c], u1 = q1[c], u2 = q2[c], u3 = q3[c];
        double x0 = q0[r], x1 = q1[r], x2 = q2[r], x3 = q3[r];
        x0 -= l * u0;
        x1 -= l * u1;
        x2 -= l * u2;
        x3 -= l * u3;
        q0[r] = x0;
        q1[r] = x1;
        q2[r] = x2;
        q3[r] = x3;
      }
      for (; cc < nb; ++cc) {
        double* q = P + (size_t)cc * pld;
        q[r] -= l * q[c];
      }
    }
    __syncthreads();
  }
  if (use_smem) {
    for (int r = tid; r < rows; r += T) {
      double* dst = C + (int64_t)j0 * ld + j0 + r;
      for (int c0 = 0; c0 < nb; c0 += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = c0 + u < nb ? P[(size_t)(c0 + u) * pld + r] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u < nb) dst[(int64_t)(c0 + u) * ld] = v[u];
      }
      Rp[j0 + r] = rp[r];
      if (rc) Rcnt[j0 + r] = rc[r];
    }
  }
  for (int r = tid; r < rows; r += T)
    if (perm[r] != r) {
      const int slot = atomicAdd(&s_cnt, 1);
      aff_pos[slot] = j0 + r;
      aff_src[slot] = j0 + perm[r];
    }
  __syncthreads();
  if (tid == 0) *aff_cnt = s_cnt;
}
__global__ void __launch_bounds__(256) k_lu_swap_solve(double* __restrict__ C, int64_t ld, int k, int j0, int nb,
                                                        const int32_t* __restrict__ aff_pos, const int32_t* __restrict__ aff_src,
                                                        const int32_t* __restrict__ aff_cnt, const int* __restrict__ flags) {
  pdl_wait();
  __shared__ double L11[LU_NB][LU_NB + 1];
  if (flags[1]) return;
  for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
    const int t = idx % nb, sc = idx / nb;
    L11[t][sc] = C[(int64_t)(j0 + sc) * ld + j0 + t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int j = blockIdx.x * 8 + (threadIdx.x >> 5);  // index over the k - nb columns outside the panel
  if (j >= k - nb) return;
  if (j >= j0) j += nb;
  double* colj = C + (int64_t)j * ld;
  const int na = *aff_cnt;  // <= 2 nb
  double v0 = 0.0, v1 = 0.0;
  if (lane < na) v0 = colj[aff_src[lane]];
  if (lane + 32 < na) v1 = colj[aff_src[lane + 32]];
  __syncwarp();
  if (lane < na) colj[aff_pos[lane]] = v0;
  if (lane + 32 < na) colj[aff_pos[lane + 32]] = v1;
  __syncwarp();
  if (j < j0) return;
  double a = lane < nb ? colj[j0 + lane] : 0.0;
  for (int sc = 0; sc < nb; ++sc) {
    const double us = __shfl_sync(FULLMASK, a, sc);
    if (lane > sc && lane < nb) a -= L11[lane][sc] * us;
  }
  if (lane < nb) colj[j0 + lane] = a;
}
// Entries per core row before the factorization (orig_row2elt_count of the reference restricted to the core: the slack of
// a core row is non-basic, so only structural columns count), and the number of stored entries of the core.
__global__ void __launch_bounds__(256) k_core_row_counts(const double* __restrict__ C, int64_t ld, int k, int32_t* __restrict__ rcnt,
                                                          unsigned long long* __restrict__ total) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  if (i < k)
    for (int t = 0; t < k; ++t) c += C[(int64_t)t * ld + i] != 0.0;
  if (i < k) rcnt[i] = c;
  unsigned s = __reduce_add_sync(FULLMASK, (unsigned)c);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, (unsigned long long)s);
}
// Off-diagonal non-zeros of the finished factors L\U (LUFactors::nnz counts lower.nondiag + upper.nondiag, lu.rs:52-54;
// exact zeros are not stored, lu.rs:253-255)
__global__ void __launch_bounds__(256) k_count_offdiag(const double* __restrict__ C, int64_t ld, int k,
                                                        unsigned long long* __restrict__ total) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t0 = blockIdx.y * 64, t1 = min(k, t0 + 64);
  int c = 0;
  if (i < k)
    for (int t = t0; t < t1; ++t) c += (t != i) && C[(int64_t)t * ld + i] != 0.0;
  unsigned s = __reduce_add_sync(FULLMASK, (unsigned)c);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, (unsigned long long)s);
}
constexpr int LU_NC = 16;
__global__ void __launch_bounds__(256) k_lu_trailing(double* __restrict__ C, int64_t ld, int k, int j0, int nb,
                                                      const int* __restrict__ flags) {
  pdl_wait();
  __shared__ double U[LU_NB][LU_NC];
  if (flags[1]) return;
  const int r0 = j0 + nb;
  const int jbase = r0 + blockIdx.x * LU_NC;
  const int ncol = min(LU_NC, k - jbase);
  for (int idx = threadIdx.x; idx < nb * LU_NC; idx += blockDim.x) {
    const int t = idx % nb, jj = idx / nb;
    U[t][jj] = jj < ncol ? C[(int64_t)(jbase + jj) * ld + j0 + t] : 0.0;
  }
  __syncthreads();
  const int r = r0 + blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= k) return;
  double Lr[LU_NB];
#pragma unroll
  for (int t = 0; t < LU_NB; ++t) Lr[t] = t < nb ? C[(int64_t)(j0 + t) * ld + r] : 0.0;
  for (int jj = 0; jj < ncol; ++jj) {
    double* q = C + (int64_t)(jbase + jj) * ld + r;
    double a = *q;
#pragma unroll
    for (int t = 0; t < LU_NB; ++t)
      if (t < nb) a -= Lr[t] * U[t][jj];
    *q = a;
  }
}
// ------------------------------------------------------------------------------------------------ K3 primal ratio test
// Harris pass 1 (solver.rs:782-795): max_step = min(max_step0, min_r (slack_r + EPS)/|alpha_r|)
__device__ __forceinline__ double leaving_step(double a, int sign, double val, double lo, double hi, bool& toward_max) {
  toward_max = (sign && a < 0.0) || (!sign && a > 0.0);  // 756
  if (toward_max) return val < hi ? hi - val : 0.0;
  return val > lo ? val - lo : 0.0;
}
__global__ void __launch_bounds__(256) k_ratio_primal_1(const double* __restrict__ alpha, const double* __restrict__ xB,
                                                         const double* __restrict__ loB, const double* __restrict__ hiB, int m,
                                                         int sign, double max_step0, double* __restrict__ red_f,
                                                         unsigned* counter, double* __restrict__ scal) {
  pdl_wait();
  __shared__ double sm[32];
  double best = INFINITY;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
    const double a = alpha[r], aa = fabs(a);
    if (aa < EPS) continue;
    bool tm;
    const double st = leaving_step(a, sign, xB[r], loB[r], hiB[r], tm);
    const double cur = (st + EPS) / aa;  // 791
    if (cur < best) best = cur;
  }
  best = block_min(best, sm);
  if (threadIdx.x == 0) red_f[blockIdx.x] = best;
  if (!last_block(counter)) return;
  double b = INFINITY;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) b = fmin(b, __ldcg(red_f + q));
  b = block_min(b, sm);
  if (threadIdx.x == 0) {
    *counter = 0;
    scal[0] = b < max_step0 ? b : max_step0;
  }
}
// Harris pass 2 (solver.rs:800-823): among rows with slack/|alpha| <= max_step the largest |alpha|;
// exact ties go to the lowest row (the reference: first in col_coeffs list order) and are COUNTED, so the
// caller knows whenever the reference's order-dependent rule would have been in play.
__global__ void __launch_bounds__(256) k_ratio_primal_2(const double* __restrict__ alpha, const double* __restrict__ xB,
                                                         const double* __restrict__ loB, const double* __restrict__ hiB, int m,
                                                         int sign, const double* __restrict__ scal, double* __restrict__ red_f,
                                                         long long* __restrict__ red_i, unsigned* counter, DevRes* res) {
  pdl_wait();
  __shared__ double smk[32];
  __shared__ long long smi[32];
  __shared__ long long smc[32];
  const double max_step = scal[0];
  KeyIdxC best{-INFINITY, LLONG_MAX, 0, 0};
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
    const double a = alpha[r], aa = fabs(a);
    if (aa < EPS) continue;
    bool tm;
    const double st = leaving_step(a, sign, xB[r], loB[r], hiB[r], tm);
    const double cur = st / aa;  // 810
    if (cur <= max_step) kic_merge(best, aa, (long long)r, 1, 1);
  }
  best = block_argmax_c(best, smk, smi, smc);
  if (threadIdx.x == 0) {
    red_f[blockIdx.x] = best.key;
    red_i[blockIdx.x] = best.idx;
    red_i[RED_CNT_OFF + blockIdx.x] = ((long long)best.ex << 32) | (unsigned)best.nr;
  }
  if (!last_block(counter)) return;
  KeyIdxC b{-INFINITY, LLONG_MAX, 0, 0};
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
    const double k = __ldcg(red_f + q);
    const long long i = __ldcg(red_i + q);
    const long long c = __ldcg(red_i + RED_CNT_OFF + q);
    kic_merge(b, k, i, (int)(c >> 32), (int)(c & 0xffffffffLL));
  }
  b = block_argmax_c(b, smk, smi, smc);
  if (threadIdx.x == 0) {
    *counter = 0;
    res->i[2] = 0;
    res->i[3] = 0;
    if (b.idx == LLONG_MAX) res->i[0] = -1;
    else {
      const int r = (int)b.idx;
      const double a = alpha[r];
      bool tm;
      leaving_step(a, sign, xB[r], loB[r], hiB[r], tm);
      res->i[0] = r;
      res->f[0] = a;
      res->f[1] = tm ? hiB[r] : loB[r];  // 813-819
      res->f[2] = xB[r];
      res->f[3] = max_step;
      res->i[2] = b.ex - 1;  // other rows sharing the winning |alpha| exactly: the reference would decide by list order
      res->i[3] = b.nr - 1;  // ... or within NEAR_TIE of it
    }
  }
}

// ------------------------------------------------------------------------------------------------ K11 dual selection
// choose_pivot_row_dual, solver.rs:855-917
__global__ void __launch_bounds__(256) k_select_row_dual(const double* __restrict__ xB, const double* __restrict__ loB,
                                                          const double* __restrict__ hiB, const double* __restrict__ w, int m,
                                                          int use_se, double* __restrict__ red_f, long long* __restrict__ red_i,
                                                          unsigned* counter, DevRes* res) {
  pdl_wait();
  __shared__ double smk[32];
  __shared__ long long smi[32];
  KeyIdx best{-INFINITY, LLONG_MAX};
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
    const double val = xB[r], mn = loB[r], mx = hiB[r];
    double infeas;
    if (val < mn - EPS) infeas = mn - val;
    else if (val > mx + EPS) infeas = val - mx;
    else continue;
    const double score = use_se ? infeas * infeas / w[r] : infeas;
    if (better_max(score, r, best.key, best.idx)) { best.key = score; best.idx = r; }
  }
  best = block_argmax(best, smk, smi);
  if (threadIdx.x == 0) { red_f[blockIdx.x] = best.key; red_i[blockIdx.x] = best.idx; }
  if (!last_block(counter)) return;
  KeyIdx b{-INFINITY, LLONG_MAX};
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
    const double k = __ldcg(red_f + q);
    const long long i = __ldcg(red_i + q);
    if (better_max(k, i, b.key, b.idx)) { b.key = k; b.idx = i; }
  }
  b = block_argmax(b, smk, smi);
  if (threadIdx.x == 0) {
    *counter = 0;
    if (b.idx == LLONG_MAX) res->i[0] = -1;
    else {
      const int r = (int)b.idx;
      res->i[0] = r;
      res->f[0] = xB[r];
      res->f[1] = loB[r];
      res->f[2] = hiB[r];
    }
  }
}

// Coupling row of a new eta K: g[j] = E_j[r_K] (j < K), a strided gather of row r_K of E, plus the per-row chains used by
// k_eta_scatter — see DESIGN.md "eta chain in closed form".
__global__ void k_eta_grow(const double* __restrict__ E, int64_t lde, int K, int rK, double* __restrict__ g,
                           int32_t* etaR, int32_t* etaPrev, int32_t* etaHead, int32_t* etaLast, int prev) {
  pdl_wait();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < K) g[j] = E[(int64_t)j * lde + rK];
  if (j == 0) {
    etaR[K] = rK;
    etaPrev[K] = prev;
    etaHead[K] = 1;
    etaLast[rK] = K;
    if (prev >= 0) etaHead[prev] = 0;
  }
}

