// On-device CSR -> CSC (second half): CsMat::to_csc (solver.rs:253, 610) / SparseMat::transpose
// (sparse.rs:230-269) as a segmented counting transpose.  The result must list every column's rows in ASCENDING order —
// that is the order the reference's transpose produces and the order every column-oriented kernel here sums in — so a
// scatter with atomics is not enough.  Rows are cut into C chunks of consecutive rows:
//   1. k_t_hist     hist[c][j] = entries of column j in chunk c (integer atomics: the count does not depend on their order)
//   2. k_t_colscan  per column, exclusive scan over the chunks (start of chunk c's entries inside column j) + column count
//   3. k_scan_excl  column counts -> csc_ptr; segment counts -> col_seg  (one CTA, three-phase scan)
//   4. k_t_fill     one CTA per chunk walks ITS rows in order; the threads take the entries of one row in parallel — a row
//                   holds a column at most once, so no two threads touch the same counter — and a barrier separates rows.
//   5. k_t_segs     segment table of the CSC copy (<= CSC_SEG consecutive entries of one column per segment)
// Deterministic, bit-identical to the host counting transpose (tests/test_sparse_gpu.py compares them).
#pragma once

// `list` (optional): the input rows are list[0..m) instead of 0..m — used to transpose a SUBSET of the columns of the CSC
// copy (the basic structural columns, at every refactorization) into a compact row-major copy whose column ids are the
// positions in the list.
__global__ void __launch_bounds__(256) k_t_hist(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t m,
                                                int64_t n, int rows_per_chunk, int32_t* __restrict__ hist,
                                                const int32_t* __restrict__ list) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= m) return;
  const int64_t src = list ? list[r] : r;
  int32_t* h = hist + (r / rows_per_chunk) * n;
  for (int64_t t = ptr[src] + lane; t < ptr[src + 1]; t += 32) atomicAdd(h + idx[t], 1);
}
__global__ void __launch_bounds__(256) k_t_colscan(int32_t* __restrict__ hist, int64_t n, int chunks, int64_t* __restrict__ cnt,
                                                   int64_t* __restrict__ segs, int seg_len) {
  pdl_wait();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int32_t run = 0;
  for (int c0 = 0; c0 < chunks; c0 += 8) {  // eight loads in flight (one per iteration: 258 us per pass at 128 chunks x 100k rows)
    int32_t t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = c0 + u < chunks ? hist[(int64_t)(c0 + u) * n + j] : 0;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (c0 + u < chunks) { hist[(int64_t)(c0 + u) * n + j] = run; run += t[u]; }
  }
  cnt[j] = run;
  segs[j] = run == 0 ? 1 : (run + seg_len - 1) / seg_len;  // an empty column keeps one empty segment
}
// out[0..n] = exclusive prefix sums of in[0..n-1] (out[n] = total).  One CTA of 1024 threads = 32 warps; every warp owns a
// contiguous slice and walks it in rows of 32 with coalesced loads and a shuffle scan per row (a thread-per-slice version with
// strided serial loads took 140 us at n = 100k).
__global__ void __launch_bounds__(1024) k_scan_excl(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  pdl_wait();
  __shared__ int64_t part[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t per = ((n + 31) / 32 + 31) / 32 * 32;  // slice length, a multiple of 32
  const int64_t b = (int64_t)w * per, e = b + per < n ? b + per : n;
  int64_t s = 0;
  for (int64_t i = b + lane; i < e; i += 32) s += in[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) part[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t run = 0;
    for (int q = 0; q < 32; ++q) { const int64_t v = part[q]; part[q] = run; run += v; }
    out[n] = run;
  }
  __syncthreads();
  int64_t carry = part[w];
  for (int64_t i0 = b; i0 < e; i0 += 32) {
    const int64_t i = i0 + lane;
    const int64_t v = i < e ? in[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (i < e) out[i] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}
__global__ void __launch_bounds__(256) k_t_fill(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                const double* __restrict__ val, int64_t m, int64_t n, int rows_per_chunk,
                                                int32_t* __restrict__ hist, const int64_t* __restrict__ csc_ptr,
                                                int32_t* __restrict__ csc_idx, double* __restrict__ csc_val,
                                                const int32_t* __restrict__ list) {
  pdl_wait();
  const int c = blockIdx.x;
  int32_t* h = hist + (int64_t)c * n;
  const int64_t r0 = (int64_t)c * rows_per_chunk, r1 = r0 + rows_per_chunk < m ? r0 + rows_per_chunk : m;
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t src = list ? list[r] : r;
    for (int64_t t = ptr[src] + threadIdx.x; t < ptr[src + 1]; t += blockDim.x) {
      const int32_t j = idx[t];
      const int32_t k = h[j];
      h[j] = k + 1;
      const int64_t pos = csc_ptr[j] + k;
      csc_idx[pos] = (int32_t)r;
      csc_val[pos] = val[t];
    }
    __syncthreads();
  }
}
// Compact row-major copy of the basic structural columns, built from their SEGMENTS (<= CSC_SEG entries each; the basis of a
// netlib-like LP holds columns of 10^4..10^5 entries next to columns of two, so the unit of work must not be a column).
// cseg_id[j]: j-th segment of the core in core-column order; chunk = j / segs_per_chunk.  Splitting a column over chunks is
// harmless: a (row, column) pair occurs once, so the order of a row's entries — ascending core column — only depends on the
// order of the segments.
__global__ void __launch_bounds__(256) k_d_hist(const int4* __restrict__ seg_desc, const int32_t* __restrict__ cseg_id, int ncseg,
                                                const int32_t* __restrict__ idx, int64_t m, int segs_per_chunk,
                                                int32_t* __restrict__ hist) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (j >= ncseg) return;
  const int4 d = seg_desc[cseg_id[j]];
  const int64_t b = ((int64_t)(unsigned)d.x) | ((int64_t)d.y << 32);
  int32_t* h = hist + (int64_t)(j / segs_per_chunk) * m;
  for (int o = lane; o < d.z; o += 32) atomicAdd(h + idx[b + o], 1);
}
__global__ void __launch_bounds__(256) k_d_fill(const int4* __restrict__ seg_desc, const int32_t* __restrict__ cseg_id, int ncseg,
                                                const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t m,
                                                int segs_per_chunk, const int32_t* __restrict__ corepos, int32_t* __restrict__ hist,
                                                const int64_t* __restrict__ dptr, int32_t* __restrict__ didx, double* __restrict__ dval) {
  pdl_wait();
  const int c = blockIdx.x;
  int32_t* h = hist + (int64_t)c * m;
  const int j0 = c * segs_per_chunk, j1 = min(ncseg, j0 + segs_per_chunk);
  for (int j = j0; j < j1; ++j) {
    const int4 d = seg_desc[cseg_id[j]];
    const int64_t b = ((int64_t)(unsigned)d.x) | ((int64_t)d.y << 32);
    const int t = corepos[d.w];  // core column of this segment's variable
    for (int o = threadIdx.x; o < d.z; o += blockDim.x) {
      const int32_t r = idx[b + o];
      const int32_t q = h[r];
      h[r] = q + 1;
      const int64_t pos = dptr[r] + q;
      didx[pos] = t;
      dval[pos] = val[b + o];
    }
    __syncthreads();
  }
}
// FTRAN tail over the compact row-major copy of the basic structural columns (dptr / didx / dval, column ids = core
// columns): alpha[cov_i] = a_i - sum_q dval[q] x[didx[q]], entries of a row in ascending core-column order.  A row holds
// nnz(D) / m entries on average (2 at k = 2000 on config 4, 15 at k = 15 000): one thread per row.  Replaces the pass over
// the WHOLE CSR copy (12 nnz bytes = 117 MB per FTRAN on config 4, two FTRANs per pivot) by 12 nnz(D) bytes.
__global__ void __launch_bounds__(256) k_ftran_finish_dcsr(const int64_t* __restrict__ dptr, const int32_t* __restrict__ didx,
                                                           const double* __restrict__ dval, int m, int k,
                                                           const double* __restrict__ xk, const double* __restrict__ rhs0,
                                                           const int32_t* __restrict__ rowcover, const int32_t* __restrict__ Jpos,
                                                           double* __restrict__ out, const uint8_t* __restrict__ touched,
                                                           uint8_t* __restrict__ touched_new) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) {
    const int p = Jpos[i];
    const double xv = xk[i];
    out[p] = xv;
    if (touched_new) touched_new[p] = (uint8_t)((xv != 0.0) | (touched[p] != 0));  // k_touch_mark, folded in
  }
  if (i >= m) return;
  const int cov = rowcover[i];
  if (cov < 0) return;
  double acc = 0.0;
  for (int64_t q = dptr[i]; q < dptr[i + 1]; ++q) acc += dval[q] * xk[didx[q]];
  const double v = rhs0[i] - acc;
  out[cov] = v;
  if (touched_new) touched_new[cov] = (uint8_t)((v != 0.0) | (touched[cov] != 0));
}
__global__ void __launch_bounds__(256) k_t_segs(const int64_t* __restrict__ csc_ptr, const int64_t* __restrict__ col_seg, int64_t n,
                                                int seg_len, int32_t* __restrict__ seg_col, int64_t* __restrict__ seg_off,
                                                int4* __restrict__ seg_desc) {
  pdl_wait();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t b = csc_ptr[j], e = csc_ptr[j + 1];
  for (int64_t sg = col_seg[j]; sg < col_seg[j + 1]; ++sg) {
    const int64_t o = b + (sg - col_seg[j]) * seg_len;
    seg_col[sg] = (int32_t)j;
    seg_off[sg] = o;
    const int64_t len = e - o < seg_len ? e - o : seg_len;
    seg_desc[sg] = make_int4((int)(unsigned)(o & 0xffffffffLL), (int)(o >> 32), (int)(len > 0 ? len : 0), (int)j);  // SegDesc
  }
}
