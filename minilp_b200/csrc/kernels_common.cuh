// Kernels of the pivot engine that do not depend on how the variables are sharded.  Included by engine.cu only.
#pragma once
// ------------------------------------------------------------------------------------------------ device helpers
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
__global__ void __launch_bounds__(CP_SEG) k_compact_count(const double* __restrict__ x, int m, int32_t* __restrict__ seg_cnt,
                                                           double* __restrict__ seg_ss, unsigned* counter,
                                                           int32_t* __restrict__ count, double* __restrict__ sumsq) {
  __shared__ double sm[32];
  __shared__ int smi[32];
  const int i = blockIdx.x * CP_SEG + threadIdx.x;
  const double v = i < m ? x[i] : 0.0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(FULLMASK, v != 0.0);
  if (lane == 0) smi[wid] = __popc(bal);
  const double ss = block_sum(v * v, sm);  // has the barriers that publish smi
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w2 = 0; w2 < 32; ++w2) c += smi[w2];
    seg_cnt[blockIdx.x] = c;
    seg_ss[blockIdx.x] = ss;
  }
  if (!last_block(counter)) return;
  if (threadIdx.x == 0) {
    int c = 0;
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) { c += __ldcg(seg_cnt + b); t += __ldcg(seg_ss + b); }
    *count = c;
    *sumsq = t;
    *counter = 0;
  }
}
__global__ void __launch_bounds__(CP_SEG) k_compact_write(const double* __restrict__ x, int m, const int32_t* __restrict__ seg_cnt,
                                                           int32_t* __restrict__ idx, double* __restrict__ val) {
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

// ------------------------------------------------------------------------------------------------ dense triangular solves
// Blocked (32-wide) triangular solve on a column-major matrix by ONE CTA of 1024 threads; used for the
// L/U factors of the basis core (LUFactors::solve lu.rs:79-106 / tri_solve_process_col 450-463) and for
// the eta-file coupling matrix G (see k_gemv_* below).
//   AXPY form (op(M) = M):   after a 32-block of unknowns is solved, every remaining row is updated
//                            (the reference's column-oriented substitution).
//   DOT form  (op(M) = M^T): before a 32-block is solved, each of its unknowns takes the dot product of
//                            its (contiguous) column with the already-solved part.
// FWD: unknowns 0..n-1, else n-1..0.  UNIT: unit diagonal.
template <bool FWD, bool AXPY, bool UNIT>
__global__ void __launch_bounds__(1024) k_trsv(const double* __restrict__ M, int64_t ld, int n, double* __restrict__ x) {
  __shared__ double xs[32];
  __shared__ double dots[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nblk = (n + 31) / 32;
  for (int bi = 0; bi < nblk; ++bi) {
    const int b = FWD ? bi * 32 : (nblk - 1 - bi) * 32;
    const int nb = min(32, n - b);
    if (!AXPY) {
      // dot products with the solved part, one warp per unknown of the block
      if (wid < nb) {
        const double* colp = M + (int64_t)(b + wid) * ld;
        double acc = 0.0;
        if (FWD) { for (int j = lane; j < b; j += 32) acc += colp[j] * x[j]; }
        else { for (int j = b + nb + lane; j < n; j += 32) acc += colp[j] * x[j]; }
        acc = warp_sum(acc);
        if (lane == 0) dots[wid] = acc;
      }
      __syncthreads();
    }
    if (wid == 0) {
      double v = 0.0, dg = 1.0;
      double coef[32];
      if (lane < nb) {
        v = x[b + lane];
        if (!AXPY) v -= dots[lane];
        if (!UNIT) dg = M[(int64_t)(b + lane) * ld + (b + lane)];
      }
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        // coefficient of unknown jj in equation `lane` of the diagonal block
        const bool need = lane < nb && jj < nb && (FWD ? (jj < lane) : (jj > lane));
        coef[jj] = need ? (AXPY ? M[(int64_t)(b + jj) * ld + (b + lane)] : M[(int64_t)(b + lane) * ld + (b + jj)]) : 0.0;
      }
#pragma unroll
      for (int s = 0; s < 32; ++s) {
        const int jj = FWD ? s : 31 - s;
        if (!UNIT && lane == jj) v = v / dg;
        const double xj = __shfl_sync(FULLMASK, v, jj);
        const bool upd = FWD ? (lane > jj) : (lane < jj);
        if (upd && jj < nb) v -= xj * coef[jj];
      }
      if (lane < nb) {
        x[b + lane] = v;
        xs[lane] = v;
      }
    }
    __syncthreads();
    if (AXPY) {
      // rhs[r] -= x_val * coeff for every remaining row (lu.rs:460-462)
      const int lo_i = FWD ? b + nb : 0;
      const int hi_i = FWD ? n : b;
      for (int i = lo_i + threadIdx.x; i < hi_i; i += 1024) {
        double acc = x[i];
        if (FWD) { for (int jj = 0; jj < nb; ++jj) acc -= xs[jj] * M[(int64_t)(b + jj) * ld + i]; }
        else { for (int jj = nb - 1; jj >= 0; --jj) acc -= xs[jj] * M[(int64_t)(b + jj) * ld + i]; }
        x[i] = acc;
      }
      __syncthreads();
    }
  }
}

// y[i] = base[i] - sum_j M[i + j*ld] * t[j]   (column-major M: rows x cols; thread per row)
// Used for: FTRAN eta application rhs -= E t (solver.rs:1310-1316 in closed form) and the slack rows of
// the basis solve alpha_S = a_S - D1 x.
__global__ void __launch_bounds__(256) k_gemv_n_sub(const double* __restrict__ M, int64_t ld, int rows, int cols,
                                                     const double* __restrict__ t, double* __restrict__ y) {
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
      for (; j + 8 <= nj; j += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[(int64_t)(j + u) * ld];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc -= ts[j + u] * v[u];
      }
      for (; j < nj; ++j) acc -= ts[j] * p[(int64_t)j * ld];
    }
  }
  if (i < rows) y[i] = acc;
}

// out[j] = base[idx[j]] - sum_i M[i + j*ld] * x[i]   (negate) or the plain dot products.
// Grid (cols, S): CTA (j, s) reduces row slice s of column j; k_gemv_t_fin adds the S partials in order.
// Used for BTRAN: u = E^T rhs (solver.rs:1326-1330) and the right-hand side of the core solve.
constexpr int GT_MAXSPLIT = 16;
__global__ void __launch_bounds__(256) k_gemv_t_part(const double* __restrict__ M, int64_t ld, int rows, int cols,
                                                      const double* __restrict__ x, double* __restrict__ part) {
  __shared__ double sm[32];
  const int j = blockIdx.x, S = gridDim.y, sidx = blockIdx.y;
  const int L = (rows + S - 1) / S;
  const int r0 = sidx * L, r1 = min(rows, r0 + L);
  const double* p = M + (int64_t)j * ld;
  double acc = 0.0;
  for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) acc += p[i] * x[i];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) part[(int64_t)sidx * cols + j] = tot;
}
__global__ void k_gemv_t_fin(const double* __restrict__ part, int S, int cols, const double* __restrict__ base,
                             const int32_t* __restrict__ base_idx, double* __restrict__ out, int negate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  double tot = 0.0;
  for (int q = 0; q < S; ++q) tot += part[(int64_t)q * cols + j];
  const double b = base ? base[base_idx ? base_idx[j] : j] : 0.0;
  out[j] = negate ? b - tot : tot;
}

__global__ void k_gather_idx(const double* __restrict__ src, const int32_t* __restrict__ idx, int cnt, double* __restrict__ dst) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) dst[t] = src[idx[t]];
}
__global__ void k_scatter_idx(const double* __restrict__ src, const int32_t* __restrict__ idx, int cnt, double* __restrict__ dst) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) dst[idx[t]] = src[t];
}
// strided gather of one row of a column-major matrix: dst[j] = M[row + j*ld]
__global__ void k_gather_row(const double* __restrict__ M, int64_t ld, int row, int cnt, double* __restrict__ dst) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) dst[t] = M[(int64_t)t * ld + row];
}
__global__ void k_fill(double* p, int64_t cnt, double v) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) p[t] = v;
}
__global__ void k_set_unit(double* p, int64_t cnt, int64_t at) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) p[t] = (t == at) ? 1.0 : 0.0;
}
// BTRAN through the eta file, last step (solver.rs:1331-1332): rhs[r_leaving(idx)] -= coeff(idx), idx = K-1..0.
// Several etas may share a leaving row; thread j owns the chain headed by the LAST eta of a row and walks it in
// the reference's order (descending idx), so the subtraction order is the reference's.
__global__ void k_eta_scatter(const double* __restrict__ s, const int32_t* __restrict__ etaR,
                              const int32_t* __restrict__ etaPrev, const int32_t* __restrict__ etaHead, int K,
                              double* __restrict__ rhs) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K || !etaHead[j]) return;
  double v = rhs[etaR[j]];
  for (int q = j; q >= 0; q = etaPrev[q]) v -= s[q];
  rhs[etaR[j]] = v;
}

// BTRAN head: rho_i = c[pos of slack i] on covered rows (U^T solve over the identity block); cov copy with zeros elsewhere
__global__ void k_btran_start(const double* __restrict__ c, const int32_t* __restrict__ rowcover, int m,
                              double* __restrict__ out, double* __restrict__ cov) {
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
__global__ void __launch_bounds__(1024) k_lu_pivot(double* __restrict__ C, int64_t ld, int k, int t,
                                                    int32_t* __restrict__ Rp, int* __restrict__ flags) {
  __shared__ double smk[32];
  __shared__ long long smi[32];
  __shared__ double s_max;
  __shared__ int s_piv;
  if (flags[1]) return;
  double* col = C + (int64_t)t * ld;
  double mx = 0.0;
  for (int i = t + threadIdx.x; i < k; i += blockDim.x) mx = fmax(mx, fabs(col[i]));
  KeyIdx r = block_argmax(KeyIdx{mx, 0}, smk, smi);
  if (threadIdx.x == 0) s_max = r.key;
  __syncthreads();
  const double max_abs = s_max;
  if (!(max_abs >= 1e-8) || isinf(max_abs)) {  // lu.rs:207-211
    if (threadIdx.x == 0) flags[1] = 1;
    return;
  }
  // lowest original row among eligible: maximise -Rp
  KeyIdx c{-INFINITY, LLONG_MAX};
  for (int i = t + threadIdx.x; i < k; i += blockDim.x)
    if (fabs(col[i]) >= 0.1 * max_abs) {
      const double key = -(double)Rp[i];
      if (better_max(key, i, c.key, c.idx)) { c.key = key; c.idx = i; }
    }
  c = block_argmax(c, smk, smi);
  if (threadIdx.x == 0) s_piv = (int)c.idx;
  __syncthreads();
  const int p = s_piv;
  if (p != t) {
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      const double a = C[(int64_t)j * ld + t], b = C[(int64_t)j * ld + p];
      C[(int64_t)j * ld + t] = b;
      C[(int64_t)j * ld + p] = a;
    }
    if (threadIdx.x == 0) { const int a = Rp[t]; Rp[t] = Rp[p]; Rp[p] = a; }
  }
  __syncthreads();
  const double pv = col[t];
  for (int i = t + 1 + threadIdx.x; i < k; i += blockDim.x) col[i] = col[i] / pv;  // lu.rs:261
}
__global__ void k_lu_update(double* __restrict__ C, int64_t ld, int k, int t, const int* __restrict__ flags) {
  if (flags[1]) return;
  const int i = t + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = t + 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i < k && j < k) C[(int64_t)j * ld + i] -= C[(int64_t)t * ld + i] * C[(int64_t)j * ld + t];
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
// exact ties go to the lowest row (the reference: first in col_coeffs list order).
__global__ void __launch_bounds__(256) k_ratio_primal_2(const double* __restrict__ alpha, const double* __restrict__ xB,
                                                         const double* __restrict__ loB, const double* __restrict__ hiB, int m,
                                                         int sign, const double* __restrict__ scal, double* __restrict__ red_f,
                                                         long long* __restrict__ red_i, unsigned* counter, DevRes* res) {
  __shared__ double smk[32];
  __shared__ long long smi[32];
  const double max_step = scal[0];
  KeyIdx best{-INFINITY, LLONG_MAX};
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
    const double a = alpha[r], aa = fabs(a);
    if (aa < EPS) continue;
    bool tm;
    const double st = leaving_step(a, sign, xB[r], loB[r], hiB[r], tm);
    const double cur = st / aa;  // 810
    if (cur <= max_step && better_max(aa, r, best.key, best.idx)) { best.key = aa; best.idx = r; }
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
      const double a = alpha[r];
      bool tm;
      leaving_step(a, sign, xB[r], loB[r], hiB[r], tm);
      res->i[0] = r;
      res->f[0] = a;
      res->f[1] = tm ? hiB[r] : loB[r];  // 813-819
      res->f[2] = xB[r];
      res->f[3] = max_step;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K11 dual selection
// choose_pivot_row_dual, solver.rs:855-917
__global__ void __launch_bounds__(256) k_select_row_dual(const double* __restrict__ xB, const double* __restrict__ loB,
                                                          const double* __restrict__ hiB, const double* __restrict__ w, int m,
                                                          int use_se, double* __restrict__ red_f, long long* __restrict__ red_i,
                                                          unsigned* counter, DevRes* res) {
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

// Coupling matrix of the eta file: G[i][j] = E_j[r_i] (j < i).  New eta K adds row K (a strided gather of row
// r_K of E) — see DESIGN.md "eta chain in closed form".
__global__ void k_eta_grow(const double* __restrict__ E, int64_t lde, int K, int rK, double* __restrict__ G, int64_t ldg,
                           int32_t* etaR, int32_t* etaPrev, int32_t* etaHead, int prev) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < K) G[(int64_t)j * ldg + K] = E[(int64_t)j * lde + rK];
  if (j == 0) {
    etaR[K] = rK;
    etaPrev[K] = prev;
    etaHead[K] = 1;
    if (prev >= 0) etaHead[prev] = 0;
  }
}

