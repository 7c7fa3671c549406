// Device kernels of the engine other than the dense price-out: column loads, sparse-storage kernels (CSC price-out, row dots,
// core extraction), pricing / ratio scans with tie accounting, the candidate exchange, pivot updates, set-up and incremental-API
// kernels.  Included by engine.cu only.
#pragma once

// ------------------------------------------------------------------------------------------------ columns
// rhs.set(column of var) (solver.rs:672-675, sparse.rs:103): column of [A|I] of LOCAL variable lv as a dense m-vector
__global__ void k_load_col(const double* __restrict__ A, int64_t lda, int64_t n, int m, int64_t lv, double* __restrict__ dst) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  dst[i] = lv < n ? A[(int64_t)i * lda + lv] : ((int64_t)i == lv - n ? 1.0 : 0.0);
}
// same, for the variable named by a candidate header that is still on the device (no host round trip)
__global__ void k_cand_load_col(const double* __restrict__ A, int64_t lda, int64_t n, int64_t c0, int64_t ng, int m,
                                const Cand* __restrict__ cand, double* __restrict__ dst, Cand* __restrict__ win_out) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && win_out) *win_out = *cand;  // single shard: the candidate IS the winner
  if (i >= m) return;
  const long long g = cand->var;
  if (g < 0) { dst[i] = 0.0; return; }
  const int64_t lv = g >= ng ? n + (g - ng) : g - c0;
  dst[i] = lv < n ? A[(int64_t)i * lda + lv] : ((int64_t)i == lv - n ? 1.0 : 0.0);
}

// FTRAN tail: alpha[pos] for slack positions = a_i - (D1 x)_i ; alpha[Jpos[t]] = x[t]   (U-solve of the
// identity-bordered basis, lu.rs:93 with B = [D | E_S]).  The t-th dense column lives in cache slot Jslot[t].
__global__ void __launch_bounds__(256) k_ftran_finish(const double* __restrict__ Bcols, int64_t ldb, int m, int k,
                                                       const double* __restrict__ xk, const double* __restrict__ rhs0,
                                                       const int32_t* __restrict__ rowcover, const int32_t* __restrict__ Jpos,
                                                       const int32_t* __restrict__ Jslot, double* __restrict__ out,
                                                       const uint8_t* __restrict__ touched, uint8_t* __restrict__ touched_new) {
  pdl_wait();
  __shared__ double ts[512];
  __shared__ int32_t sl[512];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = i < m ? rhs0[i] : 0.0;
  const int cov = i < m ? rowcover[i] : -1;
  for (int j0 = 0; j0 < k; j0 += 512) {
    const int nj = min(512, k - j0);
    __syncthreads();
    for (int q = threadIdx.x; q < nj; q += blockDim.x) { ts[q] = xk[j0 + q]; sl[q] = Jslot[j0 + q]; }
    __syncthreads();
    if (cov >= 0) {
      const double* p = Bcols + i;
      int j = 0;
      for (; j + 16 <= nj; j += 16) {  // 16 loads in flight, subtraction still in column order
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = p[(int64_t)sl[j + u] * ldb];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc -= ts[j + u] * v[u];
      }
      for (; j + 4 <= nj; j += 4) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p[(int64_t)sl[j + u] * ldb];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc -= ts[j + u] * v[u];
      }
      for (; j < nj; ++j) acc -= ts[j] * p[(int64_t)sl[j] * ldb];
    }
  }
  if (cov >= 0) { out[cov] = acc; if (touched_new) touched_new[cov] = (uint8_t)((acc != 0.0) | (touched[cov] != 0)); }
  if (i < k) {
    const int p = Jpos[i];
    const double xv = xk[i];
    out[p] = xv;
    if (touched_new) touched_new[p] = (uint8_t)((xv != 0.0) | (touched[p] != 0));  // k_touch_mark, folded in
  }
}
// FTRAN tail after a column-group split (k_tall_part): alpha_slack = a_S - sum_g part[g], alpha[Jpos[t]] = x[t]
__global__ void k_ftran_finish_parts(const double* __restrict__ part, int G, int64_t pld, int m, int k,
                                     const double* __restrict__ xk, const double* __restrict__ rhs0,
                                     const int32_t* __restrict__ rowcover, const int32_t* __restrict__ Jpos,
                                     double* __restrict__ out, const uint8_t* __restrict__ touched,
                                     uint8_t* __restrict__ touched_new) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) {
    const int cov = rowcover[i];
    if (cov >= 0) {
      double tsum = 0.0;
      for (int g = 0; g < G; ++g) tsum += part[(int64_t)g * pld + i];
      const double v = rhs0[i] - tsum;
      out[cov] = v;
      if (touched_new) touched_new[cov] = (uint8_t)((v != 0.0) | (touched[cov] != 0));
    }
  }
  if (i < k) {
    const int p = Jpos[i];
    const double xv = xk[i];
    out[p] = xv;
    if (touched_new) touched_new[p] = (uint8_t)((xv != 0.0) | (touched[p] != 0));
  }
}
// core C = D[R,:] (k x k, column-major) from the column cache
__global__ void k_extract_core(const double* __restrict__ Bcols, int64_t ldb, int k, const int32_t* __restrict__ Rp,
                               const int32_t* __restrict__ Jslot, double* __restrict__ C, int64_t ld) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (i < k && t < k) C[(int64_t)t * ld + i] = Bcols[(int64_t)Jslot[t] * ldb + Rp[i]];
}
// BTRAN: right-hand side of the core solve, rhs_t = c[Jpos[t]] - sum_i Bcols[i, slot_t] cov_i; CTA (t, s) reduces a row
// slice, k_gemv_t_fin adds the slices in order.
__global__ void __launch_bounds__(256) k_core_rhs_part(const double* __restrict__ Bcols, int64_t ldb, int rows, int k,
                                                        const int32_t* __restrict__ Jslot, const double* __restrict__ x,
                                                        double* __restrict__ part) {
  pdl_wait();
  __shared__ double sm[32];
  const int j = blockIdx.x, S = gridDim.y, sidx = blockIdx.y;
  const int L = (rows + S - 1) / S;
  const int r0 = sidx * L, r1 = min(rows, r0 + L);
  const double* p = Bcols + (int64_t)Jslot[j] * ldb;
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
  if (threadIdx.x == 0) part[(int64_t)sidx * k + j] = tot;
}

// ------------------------------------------------------------------------------------------------ sparse storage
// Sparse A (CSR + CSC, u32 indices).  The basis-inverse machinery is shared with the dense engine: basis columns are
// expanded into the dense column cache when they enter, so only three things read the sparse matrix — the column
// load, the price-out and the set-up passes.
// rhs.set(column) (solver.rs:672-675): dst is zero-filled by the caller; var < 0 comes from a candidate header.
// Every shard of a sparse-storage engine holds the WHOLE matrix (12 nnz bytes: small next to HBM; the basis operations
// need the basic columns wherever they price): n and lv are GLOBAL here.
__global__ void k_load_col_csc(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                               int64_t n, int64_t lv_arg, const Cand* __restrict__ cand, double* __restrict__ dst,
                               Cand* __restrict__ win_out) {
  pdl_wait();
  int64_t lv = lv_arg;
  if (cand && win_out && blockIdx.x == 0 && threadIdx.x == 0) *win_out = *cand;
  if (cand) {
    if (cand->var < 0) return;
    lv = cand->var;  // GLOBAL variable index
  }
  if (lv >= n) {
    if (blockIdx.x == 0 && threadIdx.x == 0) dst[lv - n] = 1.0;
    return;
  }
  const int64_t b = ptr[lv], e = ptr[lv + 1];
  for (int64_t t = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < e; t += (int64_t)gridDim.x * blockDim.x) dst[idx[t]] = val[t];
}
// Price-out over the CSC copy (calc_row_coeffs 685-692, update_primal_sq_norms 1117-1132, recalc_obj_coeffs 1216-1222,
// column norms 297-299): gathers the DENSE multiplier vector w at each column's row indices — 12 bytes per stored entry.
// Column lengths are power-law distributed (one column of a netlib-like LP can hold 10^5 entries), so the unit of work is
// a SEGMENT of at most CSC_SEG consecutive entries of one column (table built once at creation): pass 1, one warp per
// segment, rows ascending, fixed shuffle tree; pass 2 adds a column's segment sums in order.  Bit-reproducible, no atomics.
// MODE 0: out[v] = sum_i A[i,v] w[i] (slack v: w[v-n]; basic v: 0)     MODE 1: out[v] = |a_v|^2 + 1
constexpr int CSC_SEG = 1024;
// Column-sharded engines price out only the segments [sg0, sg1) of their own column block.
//
// Latency, not bandwidth, bounds this kernel: a column of the config-4 LP holds ~100 entries, so a warp spends its time in
// the dependent chain descriptor -> (row index, value) -> multiplier[row] -> shuffle tree, three global round trips per
// segment (first version: five — segment column, its offset and end, the basic flag, then the entries — 89 us per launch
// = 1.3 TB/s, profiles/r02_price_csc_full.md).  Here a segment is ONE 16-byte descriptor, the basic flag is left to
// k_price_csc_fin, and a warp works on PR_CSC_U segments at once with the first two strides of each in flight together.
// Per segment the sum is unchanged: lane-strided partial sums in ascending entry order, then the fixed shuffle tree.
struct SegDesc {
  int64_t begin;
  int32_t len, col;
};
static_assert(sizeof(SegDesc) == 16, "SegDesc is one 16-byte load");
constexpr int PR_CSC_U = 4;      // short segments (<= 64 entries) in flight per warp
constexpr int PR_CSC_LONG = 64;  // a segment with more entries is "long": one per warp, eight strides in flight
// Where the entries are: the column counts are power-law distributed, so on config 4 ~8 % of the segments (the full
// 1024-entry pieces of the ~1 % longest columns) hold ~85 % of the entries, while ~90 % of the segments are short columns of a
// few dozen entries.  Two work lists (built with the segment table): a LONG segment goes to one warp that keeps eight
// strides (256 entries: values, row indices, then the gathers) in flight — bandwidth; SHORT ones are taken four at a time
// with both strides of each issued together — latency.  Per segment the summation order is the same in both: lane-strided
// partial sums in ascending entry order, then the fixed shuffle tree (bit-identical to the first version of the kernel).
// STREAM: matrix entries are loaded with the evict-first policy (__ldcs: a matrix larger than L2 is streamed once per pivot) or
// with the default policy (the 117 MB of config 4 can stay partly L2-resident between two price-outs; MLP_CSC_STREAM picks).
template <int MODE, bool STREAM>
__global__ void __launch_bounds__(256) k_price_csc_seg(const SegDesc* __restrict__ desc, const int32_t* __restrict__ idx,
                                                       const double* __restrict__ val, const int32_t* __restrict__ long_ids,
                                                       int nlong, const int32_t* __restrict__ short_ids, int nshort,
                                                       const double* __restrict__ w, double* __restrict__ seg_sum) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  const int wid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  for (int i = wid; i < nlong; i += warps) {
    const int sg = long_ids[i];
    const int4 d = __ldg(reinterpret_cast<const int4*>(desc + sg));
    const int64_t b = ((int64_t)(unsigned)d.x) | ((int64_t)d.y << 32);
    const int len = d.z;
    double acc = 0.0;
    for (int o0 = lane; o0 < len; o0 += 256) {
      double a[8];
      int r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int o = o0 + 32 * u;
        const bool ok = o < len;
        a[u] = ok ? (STREAM ? __ldcs(val + b + o) : __ldg(val + b + o)) : 0.0;
        r[u] = (MODE == 0 && ok) ? (STREAM ? __ldcs(idx + b + o) : __ldg(idx + b + o)) : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (o0 + 32 * u < len) acc += (MODE == 0) ? a[u] * w[r[u]] : a[u] * a[u];
    }
    const double tot = warp_sum(acc);
    if (lane == 0) seg_sum[sg] = tot;
  }
  for (int base = wid; base < nshort; base += warps * PR_CSC_U) {
    int64_t b[PR_CSC_U];
    int len[PR_CSC_U], sgq[PR_CSC_U];
#pragma unroll
    for (int q = 0; q < PR_CSC_U; ++q) {
      const int i = base + q * warps;
      b[q] = 0;
      len[q] = 0;
      sgq[q] = -1;
      if (i < nshort) {
        sgq[q] = short_ids[i];
        const int4 d = __ldg(reinterpret_cast<const int4*>(desc + sgq[q]));
        b[q] = ((int64_t)(unsigned)d.x) | ((int64_t)d.y << 32);
        len[q] = d.z;
      }
    }
    double a[PR_CSC_U][2];
    int r[PR_CSC_U][2];
#pragma unroll
    for (int q = 0; q < PR_CSC_U; ++q)
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int o = lane + 32 * it;
        const bool ok = o < len[q];
        a[q][it] = ok ? (STREAM ? __ldcs(val + b[q] + o) : __ldg(val + b[q] + o)) : 0.0;
        r[q][it] = (MODE == 0 && ok) ? (STREAM ? __ldcs(idx + b[q] + o) : __ldg(idx + b[q] + o)) : 0;
      }
#pragma unroll
    for (int q = 0; q < PR_CSC_U; ++q) {
      double acc = 0.0;
#pragma unroll
      for (int it = 0; it < 2; ++it)
        if (lane + 32 * it < len[q]) acc += (MODE == 0) ? a[q][it] * w[r[q][it]] : a[q][it] * a[q][it];
      const double tot = warp_sum(acc);
      if (lane == 0 && sgq[q] >= 0) seg_sum[sgq[q]] = tot;
    }
  }
}
template <int MODE>
__global__ void __launch_bounds__(256) k_price_csc_fin(const int64_t* __restrict__ col_seg, const double* __restrict__ seg_sum,
                                                       int64_t n, int64_t m, int64_t c0, const double* __restrict__ w,
                                                       const uint8_t* __restrict__ vflag, double* __restrict__ out) {
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // LOCAL variable
  if (v < n) {
    double t = 0.0;
    for (int64_t sg = col_seg[c0 + v]; sg < col_seg[c0 + v + 1]; ++sg) t += seg_sum[sg];
    out[v] = (MODE == 1) ? t + 1.0 : ((vflag[v] & MLP_BASIC) ? 0.0 : t);
  } else if (v < n + m) {
    out[v] = (MODE == 1) ? 2.0 : ((vflag[v] & MLP_BASIC) ? 0.0 : w[v - n]);
  }
}
// rows of A x_N over the CSR copy (solver.rs:234-238): one warp per row
// (a shard sums only the entries of its own column block [c0, c0 + n_loc); xnb is indexed by LOCAL variable)
__global__ void __launch_bounds__(256) k_row_dot_csr(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                     const double* __restrict__ val, int64_t m, int64_t c0, int64_t n_loc,
                                                     const double* __restrict__ xnb, double* __restrict__ out) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= m) return;
  double acc = 0.0;
  for (int64_t t = ptr[r] + lane; t < ptr[r + 1]; t += 32) {
    const int64_t j = (int64_t)idx[t] - c0;
    if (j >= 0 && j < n_loc) acc += val[t] * xnb[j];
  }
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc;
}

// The three places where the basis machinery touches the basic structural columns, read from the sparse matrix itself
// instead of a dense m x k column cache (12 bytes per stored entry instead of 8 m k):
//   corevar[t]  structural variable of core column t        corepos[v]  core column of variable v, or -1
//   rowcore[i]  core row of constraint row i (i in R), or -1
// FTRAN tail: alpha[cov_i] = a_i - sum_{j in row i, j in the core} A[i,j] x[corepos[j]]  (warp per CSR row)
__global__ void __launch_bounds__(256) k_ftran_finish_csr(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                          const double* __restrict__ val, int m, int k,
                                                          const double* __restrict__ xk, const double* __restrict__ rhs0,
                                                          const int32_t* __restrict__ rowcover, const int32_t* __restrict__ Jpos,
                                                          const int32_t* __restrict__ corepos, double* __restrict__ out) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt < k) out[Jpos[gt]] = xk[gt];
  const int64_t i = gt >> 5;
  if (i >= m) return;
  const int cov = rowcover[i];
  if (cov < 0) return;
  double acc = 0.0;
  for (int64_t t = ptr[i] + lane; t < ptr[i + 1]; t += 32) {
    const int c = corepos[idx[t]];
    if (c >= 0) acc += val[t] * xk[c];
  }
  acc = warp_sum(acc);
  if (lane == 0) out[cov] = rhs0[i] - acc;
}
// BTRAN core right-hand side: x[t] = c[Jpos[t]] - sum_i A[i, corevar[t]] cov[i], over the segments of the core columns
// (cseg_id[j] = global segment, cseg_first[t] = first entry of core column t in that list; built at each refactorization)
__global__ void __launch_bounds__(256) k_core_rhs_seg(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                      const double* __restrict__ val, const int32_t* __restrict__ seg_col,
                                                      const int64_t* __restrict__ seg_off, const int32_t* __restrict__ cseg_id,
                                                      int ncseg, const double* __restrict__ cov, double* __restrict__ csum) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= ncseg) return;
  const int64_t sg = cseg_id[j];
  const int v = seg_col[sg];
  const int64_t b = seg_off[sg], e = min(b + (int64_t)CSC_SEG, ptr[v + 1]);
  double acc = 0.0;
  for (int64_t q = b + lane; q < e; q += 32) acc += val[q] * cov[idx[q]];
  acc = warp_sum(acc);
  if (lane == 0) csum[j] = acc;
}
__global__ void k_core_rhs_fin(const double* __restrict__ csum, const int32_t* __restrict__ cseg_first, int k,
                               const double* __restrict__ c, const int32_t* __restrict__ Jpos, double* __restrict__ x) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k) return;
  double tsum = 0.0;
  for (int j = cseg_first[t]; j < cseg_first[t + 1]; ++j) tsum += csum[j];
  x[t] = c[Jpos[t]] - tsum;
}
// core C = D[R,:]: scatter the stored entries of each core column that fall into core rows (C zero-filled before)
__global__ void __launch_bounds__(256) k_extract_core_seg(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                          const double* __restrict__ val, const int32_t* __restrict__ seg_col,
                                                          const int64_t* __restrict__ seg_off, const int32_t* __restrict__ cseg_id,
                                                          int ncseg, const int32_t* __restrict__ corepos,
                                                          const int32_t* __restrict__ rowcore, double* __restrict__ C, int64_t ld) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= ncseg) return;
  const int64_t sg = cseg_id[j];
  const int v = seg_col[sg];
  const int t = corepos[v];
  const int64_t b = seg_off[sg], e = min(b + (int64_t)CSC_SEG, ptr[v + 1]);
  for (int64_t q = b + lane; q < e; q += 32) {
    const int r = rowcore[idx[q]];
    if (r >= 0) C[(int64_t)t * ld + r] = val[q];
  }
}
__global__ void k_set_corepos(int32_t* __restrict__ corepos, const int32_t* __restrict__ corevar, int k, int clear) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < k) corepos[corevar[t]] = clear ? -1 : t;
}

// ------------------------------------------------------------------------------------------------ K1 pricing scan
// choose_pivot, solver.rs:696-739: arg-max of d^2/gamma (or |d|) over eligible non-basic variables, strict '>' in
// ascending position order => lowest position wins ties.  Writes this shard's candidate header.
__global__ void __launch_bounds__(256) k_select_primal(const double* __restrict__ d, const double* __restrict__ gam,
                                                        const uint8_t* __restrict__ vflag, const int32_t* __restrict__ vpos,
                                                        int64_t nt, int64_t n, int64_t c0, int64_t ng, int use_se,
                                                        double* __restrict__ red_f, long long* __restrict__ red_i,
                                                        unsigned* counter, const double* __restrict__ xnb,
                                                        const int* __restrict__ flags, Cand* out) {
  pdl_wait();
  __shared__ double smk[32];
  __shared__ long long smi[32];
  KeyIdx best{-INFINITY, LLONG_MAX};
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nt; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned f = vflag[v];
    if (f & MLP_BASIC) continue;
    const double dv = d[v];
    if (((f & MLP_AT_MIN) && dv > -EPS) || ((f & MLP_AT_MAX) && dv < EPS)) continue;  // 705-708
    const double score = use_se ? dv * dv / gam[v] : fabs(dv);
    const long long key2 = ((long long)vpos[v] << 32) | (long long)v;  // position decides ties
    if (better_max(score, key2, best.key, best.idx)) { best.key = score; best.idx = key2; }
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
    out->f[4] = flags[2] ? 2.0 : (double)flags[0];
    if (b.idx == LLONG_MAX) { out->var = -1; out->key = -INFINITY; out->tie = LLONG_MAX; }
    else {
      const long long v = b.idx & 0xffffffffLL;
      out->key = b.key;
      out->tie = (b.idx >> 32) << 32;
      out->var = v < n ? c0 + v : ng + (v - n);
      out->f[0] = d[v];
      out->f[1] = xnb[v];
    }
  }
}

// low word of the winner's `tie` after absorbing a losing candidate of another shard: that candidate and its own ties
// count towards the winner's when the keys agree exactly (low 16 bits) / within NEAR_TIE (next 16 bits)
__host__ __device__ __forceinline__ long long merge_ties(long long wt, double wk, long long ct, double ck) {
  long long e = wt & 0xffff, n = (wt >> 16) & 0xffff;
  if (ck == wk) e += 1 + (ct & 0xffff);
  if (ck >= wk * (1.0 - 1e-9)) n += 1 + ((ct >> 16) & 0xffff);
  if (e > 65535) e = 65535;
  if (n > 65535) n = 65535;
  return (wt & ~0xffffffffLL) | (n << 16) | e;
}
// Arg-reduce of the gathered candidate headers ON THE DEVICE (larger key wins, ties go to the smaller `tie`: lowest
// position, solver.rs:719) and copy of the winner's column into colq, so that the FTRAN of the entering column can be
// queued behind the selection without a host round trip.  Every thread repeats the <= 8-way comparison.
__global__ void __launch_bounds__(256) k_pick_winner(const char* __restrict__ recv, size_t xbytes, int world, int m,
                                                      double* __restrict__ colq, Cand* __restrict__ win) {
  pdl_wait();
  int best = -1;
  double err = 0.0;
  for (int r = 0; r < world; ++r) {
    const Cand* c = reinterpret_cast<const Cand*>(recv + (size_t)r * xbytes);
    if (c->f[4] != 0.0) err = 1.0;
    if (c->var < 0) continue;
    if (best < 0) { best = r; continue; }
    const Cand* b = reinterpret_cast<const Cand*>(recv + (size_t)best * xbytes);
    if (c->key > b->key || (c->key == b->key && c->tie < b->tie)) best = r;
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) colq[i] = best < 0 ? 0.0 : reinterpret_cast<const double*>(recv + (size_t)best * xbytes + sizeof(Cand))[i];
  if (i == 0) {
    if (best < 0) { win->var = -1; win->key = -INFINITY; win->tie = LLONG_MAX; }
    else {
      Cand w = *reinterpret_cast<const Cand*>(recv + (size_t)best * xbytes);
      for (int r = 0; r < world; ++r) {  // ties of the winner across shards (dual ratio test)
        const Cand* c = reinterpret_cast<const Cand*>(recv + (size_t)r * xbytes);
        if (r != best && c->var >= 0) w.tie = merge_ties(w.tie, w.key, c->tie, c->key);
      }
      *win = w;
    }
    win->f[4] = err;
  }
}

// The same exchange as ONE kernel over NVLink peer memory (one process per GPU, buffers mapped with CUDA IPC): every
// rank stores its 64-byte candidate header into a mailbox slot of every peer and raises the slot's sequence flag
// (system-scope fence in between); every rank then polls its OWN mailbox, arg-reduces the headers, and pulls the
// winner's column (8 m bytes) straight out of the owner's memory.  Compared with the all-gather it moves one column
// instead of `world` and has no collective launch latency.  Buffers and mailboxes are double-buffered by exchange
// parity: a rank can be at most one exchange ahead of a peer, because finishing exchange s needs every peer's flag s.
struct PeerTable { char* base[8]; };
constexpr int P2P_SLOT = 128;  // 64 B header + flag, padded
__global__ void __launch_bounds__(256) k_exchange_p2p(PeerTable pt, int rank, int world, unsigned long long seq, int parity,
                                                       const Cand* __restrict__ mine, int m, size_t col_bytes, size_t box_off,
                                                       double* __restrict__ colq, Cand* __restrict__ win) {
  pdl_wait();
  __shared__ Cand hdr[8];
  __shared__ int s_best;
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  if (blockIdx.x == 0 && threadIdx.x < world) {  // publish into peer `threadIdx.x`
    char* slot = pt.base[threadIdx.x] + box_off + ((size_t)parity * world + rank) * P2P_SLOT;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(mine);
    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(slot);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = src[q];
    __threadfence_system();
    dst[8] = seq;
  }
  __syncthreads();
  if (threadIdx.x < world) {  // wait for peer `threadIdx.x`'s header in my own mailbox
    const char* slot = pt.base[rank] + box_off + ((size_t)parity * world + threadIdx.x) * P2P_SLOT;
    const volatile unsigned long long* src = reinterpret_cast<const volatile unsigned long long*>(slot);
    const long long t0 = clock64();
    bool ok = true;
    while (src[8] != seq) {
      if (clock64() - t0 > 120000000000LL) { ok = false; break; }  // ~60 s: a peer died; report instead of hanging forever
      __nanosleep(64);
    }
    __threadfence_system();
    unsigned long long* d = reinterpret_cast<unsigned long long*>(&hdr[threadIdx.x]);
#pragma unroll
    for (int q = 0; q < 8; ++q) d[q] = src[q];
    if (!ok) s_bad = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = -1;
    for (int r = 0; r < world; ++r) {
      if (hdr[r].var < 0) continue;
      if (best < 0 || hdr[r].key > hdr[best].key || (hdr[r].key == hdr[best].key && hdr[r].tie < hdr[best].tie)) best = r;
    }
    s_best = s_bad ? -1 : best;
  }
  __syncthreads();
  const int best = s_best;
  const double* src = best < 0 ? nullptr : reinterpret_cast<const double*>(pt.base[best] + (size_t)parity * col_bytes);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    colq[i] = best < 0 ? 0.0 : __ldcv(src + i);  // peer memory: never served from a stale cache line
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double err = s_bad ? 2.0 : 0.0;
    for (int r = 0; r < world; ++r) if (hdr[r].f[4] != 0.0 && err == 0.0) err = 1.0;
    if (best < 0) { win->var = -1; win->key = -INFINITY; win->tie = LLONG_MAX; }
    else {
      Cand w = hdr[best];
      for (int r = 0; r < world; ++r)
        if (r != best && hdr[r].var >= 0) w.tie = merge_ties(w.tie, w.key, hdr[r].tie, hdr[r].key);
      *win = w;
    }
    win->f[4] = err;
  }
}

// ------------------------------------------------------------------------------------------------ K11 dual column
// choose_entering_col_dual, solver.rs:919-1021
__device__ __forceinline__ bool dual_eligible(double coeff, unsigned f, int leaving_diff_sign) {
  bool entering_diff_sign;
  if (coeff >= EPS) entering_diff_sign = !leaving_diff_sign;
  else if (coeff <= -EPS) entering_diff_sign = leaving_diff_sign;
  else return false;
  return entering_diff_sign ? !(f & MLP_AT_MAX) : !(f & MLP_AT_MIN);
}
__device__ __forceinline__ double clamp_obj(double oc, unsigned f) {
  if ((f & MLP_AT_MIN) && oc < 0.0) oc = 0.0;
  if ((f & MLP_AT_MAX) && oc > 0.0) oc = 0.0;
  return oc;
}
__global__ void __launch_bounds__(256) k_ratio_dual_1(const double* __restrict__ rc, const double* __restrict__ d,
                                                       const uint8_t* __restrict__ vflag, int64_t nt, int lds,
                                                       double* __restrict__ red_f, unsigned* counter, double* __restrict__ scal,
                                                       const DevRes* __restrict__ dr) {
  pdl_wait();
  __shared__ double sm[32];
  if (dr) lds = dr->f[0] < dr->f[1] ? 1 : 0;  // the row's record left by k_select_row_dual: leaving_new_val > basic value (solver.rs:908-915, 925)
  double best = INFINITY;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nt; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned f = vflag[v];
    if (f & MLP_BASIC) continue;
    const double coeff = rc[v];
    if (!dual_eligible(coeff, f, lds)) continue;
    const double oc = clamp_obj(d[v], f);
    const double cur = (fabs(oc) + EPS) / fabs(coeff);  // 970
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
    scal[0] = b;
  }
}
__global__ void k_min_small(const double* __restrict__ vals, int cnt, double* __restrict__ out) {
  pdl_wait();
  double b = INFINITY;
  for (int q = 0; q < cnt; ++q) b = fmin(b, vals[q]);
  *out = b;
}
// pass 2: exact ties in |coeff| go to the lowest GLOBAL variable index (the reference: first-touch order) and
// are counted (candidate header `tie`, low word)
__global__ void __launch_bounds__(256) k_ratio_dual_2(const double* __restrict__ rc, const double* __restrict__ d,
                                                       const uint8_t* __restrict__ vflag, const int32_t* __restrict__ vpos,
                                                       const double* __restrict__ xnb, int64_t nt, int64_t n, int64_t c0,
                                                       int64_t ng, int lds, const double* __restrict__ scal,
                                                       double* __restrict__ red_f, long long* __restrict__ red_i,
                                                       unsigned* counter, const int* __restrict__ flags, Cand* out,
                                                       int scan_slacks, const DevRes* __restrict__ dr) {
  pdl_wait();
  if (dr) lds = dr->f[0] < dr->f[1] ? 1 : 0;  // as in k_ratio_dual_1
  __shared__ double smk[32];
  __shared__ long long smi[32];
  __shared__ long long smc[32];
  const double max_step = scal[0];
  KeyIdxC best{-INFINITY, LLONG_MAX, 0, 0};
  // The slack variables are replicated on every shard: only ONE shard (rank 0) proposes and counts them, so that every
  // variable is scanned exactly once across the shards and the tie counts add up to the single-shard ones.
  const int64_t vend = scan_slacks ? nt : n;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < vend; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned f = vflag[v];
    if (f & MLP_BASIC) continue;
    const double coeff = rc[v];
    if (!dual_eligible(coeff, f, lds)) continue;
    const double oc = clamp_obj(d[v], f);
    const double cur = fabs(oc) / fabs(coeff);  // 993
    const long long g = v < n ? c0 + v : ng + (v - n);
    if (cur <= max_step) kic_merge(best, fabs(coeff), g, 1, 1);
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
    out->f[4] = flags[2] ? 2.0 : (double)flags[0];
    if (b.idx == LLONG_MAX) { out->var = -1; out->key = -INFINITY; out->tie = LLONG_MAX; }
    else {
      const long long g = b.idx;
      const long long v = g >= ng ? n + (g - ng) : g - c0;
      out->key = b.key;
      out->tie = (g << 32) | pack_ties(b.ex, b.nr);
      out->var = g;
      out->f[0] = rc[v];
      out->f[1] = d[v];
      out->f[2] = xnb[v];
      out->f[3] = (double)vpos[v];
    }
  }
}

// ------------------------------------------------------------------------------------------------ pivot updates
// Row half of Solver::pivot: basic values (solver.rs:1049-1055), dual steepest-edge norms (update_dual_sq_norms
// 1163-1173) and the new eta column (push_eta_matrix 1274-1284).  Replicated on every shard.
__global__ void __launch_bounds__(256) k_pivot_rows(const double* __restrict__ alpha, const double* __restrict__ tau,
                                                     double* __restrict__ xB, double* __restrict__ w, int m, int row,
                                                     double entering_new_val, double entering_diff, double coeff, int has_elem,
                                                     int dse, const double* __restrict__ scal, double* __restrict__ eta_col,
                                                     int* __restrict__ flags, uint8_t* __restrict__ touched,
                                                     const uint8_t* __restrict__ touched_new) {
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  if (eta_col) touched[r] = touched_new[r];  // the pushed eta stores every listed position of col_coeffs (solver.rs:1274-1284)
  const double a = alpha[r];
  if (!has_elem) {  // bound flip, solver.rs:1035-1037
    if (a != 0.0) xB[r] -= entering_diff * a;
    return;
  }
  if (r == row) xB[r] = entering_new_val;
  else if (a != 0.0) xB[r] -= entering_diff * a;
  if (dse) {
    const double pivot_sq_norm = scal[1];  // |rho|^2, solver.rs:1160
    const double pcs = coeff * coeff;
    if (r == row) {
      w[r] = pivot_sq_norm / pcs;
      if (!isfinite(w[r])) flags[0] = 1;
    } else if (a != 0.0) {
      const double nw = w[r] + (-2.0 * a * tau[r] / coeff + pivot_sq_norm * a * a / pcs);  // 1168-1169
      w[r] = nw;
      if (!isfinite(nw)) flags[0] = 1;
    }
  }
  if (eta_col) eta_col[r] = (r == row) ? 1.0 - 1.0 / coeff : a / coeff;  // 1276-1280
}
__global__ void k_flip_var(double* xnb, uint8_t* vflag, const double* lo, const double* hi, int64_t q, int64_t ql, double new_val) {
  pdl_wait();
  xnb[ql] = new_val;
  unsigned f = vflag[ql] & MLP_FIXED;
  if (new_val == lo[q]) f |= MLP_AT_MIN;
  if (new_val == hi[q]) f |= MLP_AT_MAX;
  vflag[ql] = (uint8_t)f;  // solver.rs:1038-1040
}

// The variable half of Solver::pivot and the NEXT pricing scan in one pass over this shard's variables (SURVEY K1
// "fused update+select"): per variable — finish the N^T v price-out (sum of the chunk partials in chunk order, as
// k_price_finish), update reduced cost and primal steepest-edge norm (k_pivot_vars), apply the basis swap to the two
// variables concerned (k_pivot_swap), then score the variable for choose_pivot (k_select_primal) with its new state.
// Block partial arg-max -> last block finishes and leaves the candidate header for the exchange step.
struct UpdSel {
  // price-out finish (pse only; sparse storage has helper already)
  const double* partial; const int32_t* count_ptr; int64_t lda; const double* slack_vals; double* helper; int finish;
  // pivot
  int64_t q, ql, lvl, lv; int col, row; double pivot_obj, coeff, leaving_new_val; int pse;
};
__global__ void __launch_bounds__(256) k_update_select(UpdSel a, double* __restrict__ d, double* __restrict__ gam,
                                                        const double* __restrict__ rc, double* __restrict__ xnb,
                                                        uint8_t* __restrict__ vflag, int32_t* __restrict__ vpos,
                                                        int32_t* __restrict__ bvar, double* __restrict__ loB,
                                                        double* __restrict__ hiB, const double* __restrict__ lo,
                                                        const double* __restrict__ hi, int64_t nt, int64_t n, int64_t m,
                                                        int64_t c0, int64_t ng, const double* __restrict__ scal,
                                                        int* __restrict__ flags, double* __restrict__ red_f,
                                                        long long* __restrict__ red_i, unsigned* counter, DevRes* res, Cand* out) {
  pdl_wait();
  __shared__ double smk[32];
  __shared__ long long smi[32];
  KeyIdx best{-INFINITY, LLONG_MAX};
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nt) {
    unsigned f = vflag[v];
    double h = 0.0;
    if (a.pse) {
      if (a.finish) {
        if (v < n) {
          const int C = price_chunks_for(*a.count_ptr);
          for (int c = 0; c < C; ++c) h += a.partial[(int64_t)c * a.lda + v];
        } else h = a.slack_vals[v - n];
        if (f & MLP_BASIC) h = 0.0;
        a.helper[v] = h;
      } else h = a.helper[v];
    }
    double dv = d[v], gv = gam[v];
    if (v == a.lvl) {  // the leaving variable takes the non-basic slot (solver.rs:1066-1071, 1076, 1142)
      xnb[v] = a.leaving_new_val;
      f = 0;
      if (a.leaving_new_val == lo[a.lv]) f |= MLP_AT_MIN;
      if (a.leaving_new_val == hi[a.lv]) f |= MLP_AT_MAX;
      vflag[v] = (uint8_t)f;
      vpos[v] = a.col;
      dv = -a.pivot_obj;
      d[v] = dv;
      if (a.pse) {
        gv = (scal[2] + 1.0) / (a.coeff * a.coeff);
        gam[v] = gv;
        if (!isfinite(gv)) flags[0] = 1;
      }
    } else if (v == a.ql) {  // the entering variable becomes basic (1088-1091)
      f = MLP_BASIC;
      vflag[v] = MLP_BASIC;
      vpos[v] = a.row;
    } else if (!(f & MLP_BASIC)) {
      const double c = rc[v];
      if (c != 0.0) {
        dv -= a.pivot_obj * c;  // 1073-1080
        d[v] = dv;
        if (a.pse) {
          const double psn = scal[2] + 1.0;  // 1136
          gv = gv + (-2.0 * c * h / a.coeff + psn * c * c / (a.coeff * a.coeff));  // 1144-1146
          gam[v] = gv;
          if (!isfinite(gv)) flags[0] = 1;
        }
      }
    }
    if (v == 0) {  // row-side bookkeeping, identical on every shard (1057-1058, 1088)
      loB[a.row] = lo[a.q];
      hiB[a.row] = hi[a.q];
      res->i[0] = bvar[a.row];  // the device's idea of the leaving variable, cross-checked by the host
      bvar[a.row] = (int32_t)a.q;
    }
    // choose_pivot's scan (696-739) on the updated state
    if (!(f & MLP_BASIC) && !(((f & MLP_AT_MIN) && dv > -EPS) || ((f & MLP_AT_MAX) && dv < EPS))) {
      best.key = a.pse ? dv * dv / gv : fabs(dv);
      best.idx = ((long long)vpos[v] << 32) | (long long)v;
    }
  }
  best = block_argmax(best, smk, smi);
  if (threadIdx.x == 0) { red_f[blockIdx.x] = best.key; red_i[blockIdx.x] = best.idx; }
  if (!last_block(counter)) return;
  KeyIdx b{-INFINITY, LLONG_MAX};
  for (int qd = threadIdx.x; qd < (int)gridDim.x; qd += blockDim.x) {
    const double k = __ldcg(red_f + qd);
    const long long i = __ldcg(red_i + qd);
    if (better_max(k, i, b.key, b.idx)) { b.key = k; b.idx = i; }
  }
  b = block_argmax(b, smk, smi);
  if (threadIdx.x == 0) {
    *counter = 0;
    const int nf = *((volatile int*)flags);
    res->flags[0] = nf;
    res->flags[1] = flags[1];
    out->f[4] = flags[2] ? 2.0 : (double)nf;
    if (b.idx == LLONG_MAX) { out->var = -1; out->key = -INFINITY; out->tie = LLONG_MAX; }
    else {
      const long long vv = b.idx & 0xffffffffLL;
      out->key = b.key;
      out->tie = (b.idx >> 32) << 32;
      out->var = vv < n ? c0 + vv : ng + (vv - n);
      out->f[0] = __ldcg(d + vv);
      out->f[1] = __ldcg(xnb + vv);
    }
  }
}

// ------------------------------------------------------------------------------------------------ init kernels
// partial of A x_N over this shard's columns (solver.rs:234-238). One CTA per row.
__global__ void __launch_bounds__(256) k_row_dot(const double* __restrict__ A, int64_t lda, int64_t n,
                                                  const double* __restrict__ xnb, double* __restrict__ out) {
  pdl_wait();
  __shared__ double sm[32];
  const int r = blockIdx.x;
  const double* row = A + (int64_t)r * lda;
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) acc += row[j] * xnb[j];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) out[r] = tot;
}
// basic_var_vals = rhs - sum over shards (in rank order) of the partial products
__global__ void k_init_basic_vals(const double* __restrict__ parts, int world, int m, const double* __restrict__ rhs,
                                  double* __restrict__ xB) {
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  double tot = 0.0;
  for (int g = 0; g < world; ++g) tot += parts[(int64_t)g * m + r];
  xB[r] = rhs[r] - tot;
}
// d_N = c_N - N^T y (recalc_obj_coeffs, solver.rs:1216-1222)
__global__ void k_recalc_d(const double* __restrict__ cobj, const double* __restrict__ rc, const uint8_t* __restrict__ vflag,
                           int64_t nt, int64_t n, int64_t c0, int64_t ng, double* __restrict__ d) {
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nt || (vflag[v] & MLP_BASIC)) return;
  const int64_t g = v < n ? c0 + v : ng + (v - n);
  d[v] = cobj[g] - rc[v];
}
// objective from scratch (solver.rs:1224-1230) in three parts: basic rows, non-basic slacks (both replicated),
// non-basic structurals of this shard.  Single CTA, deterministic.
__global__ void __launch_bounds__(1024) k_recalc_obj(const double* __restrict__ cobj, const int32_t* __restrict__ bvar,
                                                      const double* __restrict__ xB, int m, const double* __restrict__ xnb,
                                                      const uint8_t* __restrict__ vflag, int64_t n, int64_t c0, int64_t ng,
                                                      double* __restrict__ out3) {
  pdl_wait();
  __shared__ double sm[32];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int r = threadIdx.x; r < m; r += blockDim.x) a += cobj[bvar[r]] * xB[r];
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x)
    if (!(vflag[n + i] & MLP_BASIC)) b += cobj[ng + i] * xnb[n + i];
  for (int64_t v = threadIdx.x; v < n; v += blockDim.x)
    if (!(vflag[v] & MLP_BASIC)) c += cobj[c0 + v] * xnb[v];
  const double ta = block_sum(a, sm);
  const double tb = block_sum(b, sm);
  const double tc = block_sum(c, sm);
  if (threadIdx.x == 0) { out3[0] = ta; out3[1] = tb; out3[2] = tc; }
}
__global__ void k_gather_cB(const double* __restrict__ cobj, const int32_t* __restrict__ bvar, int m, double* __restrict__ out) {
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m) out[r] = cobj[bvar[r]];
}

// ------------------------------------------------------------------------------------------------ incremental API (row f2)
// Solver::add_constraint (solver.rs:549-634) pieces.  A cut may carry coefficients g_i on slack variables
// (add_gomory_cut, 440-460); slack columns stay unit columns here, so s_i = rhs_i - a_i x is substituted:
// row' = c - A^T g, rhs' = rhs - g . rhs_old (the same constraint; see DESIGN.md §8 for what that changes).
__global__ void k_row_combine(double* __restrict__ row, const double* __restrict__ partial, const int32_t* __restrict__ count_ptr,
                              int64_t lda, int64_t n) {
  pdl_wait();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int C = price_chunks_for(*count_ptr);
  double t = 0.0;
  for (int c = 0; c < C; ++c) t += partial[(int64_t)c * lda + j];
  row[j] -= t;
}
// out[0] = base - sum_i a_i b_i (single CTA, deterministic); used for rhs' and for the new basic value rhs - a . x
__global__ void __launch_bounds__(1024) k_sub_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t cnt, double base,
                                                   const double* __restrict__ base_ptr, double* __restrict__ out) {
  pdl_wait();
  __shared__ double sm[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) acc += a[i] * b[i];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) out[0] = (base_ptr ? *base_ptr : base) - tot;
}
// current value of every structural variable (Solver::get_value, 371-376) as a dense vector
__global__ void k_struct_values(const double* __restrict__ xnb, const double* __restrict__ xB, const uint8_t* __restrict__ vflag,
                                const int32_t* __restrict__ vpos, int64_t n, double* __restrict__ out) {
  pdl_wait();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = (vflag[j] & MLP_BASIC) ? xB[vpos[j]] : xnb[j];
}
// state of the appended row and of its slack variable (563-571, 591)
__global__ void k_new_row_state(int64_t r, int64_t lv, int64_t gv, double smin, double smax, const double* __restrict__ val,
                                const double* __restrict__ rhs_new, double* lo, double* hi, double* cobj, double* d, double* gam,
                                double* xnb, uint8_t* vflag, int32_t* vpos, int32_t* bvar, double* xB, double* loB, double* hiB,
                                double* w, double* rhs, int32_t* rowcover) {
  pdl_wait();
  lo[gv] = smin; hi[gv] = smax; cobj[gv] = 0.0;
  d[lv] = 0.0; gam[lv] = 0.0; xnb[lv] = 0.0;
  vflag[lv] = MLP_BASIC; vpos[lv] = (int32_t)r;
  bvar[r] = (int32_t)gv; xB[r] = *val; loB[r] = smin; hiB[r] = smax; w[r] = 1.0; rhs[r] = *rhs_new;
  rowcover[r] = (int32_t)r;
}
// the cached basis columns get their entry of the new row
__global__ void k_bcols_new_row(const double* __restrict__ rowA, const int32_t* __restrict__ slots, const int32_t* __restrict__ vars,
                                int cnt, int64_t ldb, int64_t r, double* __restrict__ Bcols) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) Bcols[(int64_t)slots[t] * ldb + r] = rowA[vars[t]];
}
// primal_edge_sq_norms[c] += coeff^2 over the new tableau row (618-622)
__global__ void k_add_sq(double* __restrict__ gam, const double* __restrict__ rc, const uint8_t* __restrict__ vflag, int64_t nt) {
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nt && !(vflag[v] & MLP_BASIC)) gam[v] += rc[v] * rc[v];
}
__global__ void k_copy1(double* dst, const double* src) {
  pdl_wait(); *dst = *src; }
__global__ void k_set_var_state(uint8_t* vflag, int64_t lv, unsigned f) {
  pdl_wait(); vflag[lv] = (uint8_t)f; }

