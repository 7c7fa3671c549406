// The lane-0 chain of a primal pivot with steepest edge as ONE cooperative kernel.  Included by engine.cu only.
//
//   alpha_q = B^-1 a_q            calc_col_coeffs, solver.rs:671-677 -> BasisSolver::solve 1305-1319
//   nnz(alpha_q), |alpha_q|^2     SparseVec::sq_norm (sparse.rs:32-34), eta bookkeeping 1096-1099
//   v = B^-T alpha_q              update_primal_sq_norms, solver.rs:1114 -> BasisSolver::solve_transp 1322-1338
//   ordered non-zero list of v    ScatteredVec::to_sparse_vec, sparse.rs:115-121
//
// As separate launches (ftran / compact / btran / compact in engine.cu) this is 17 dependent kernels of 3-20 us whose
// bodies move a few hundred KB each: at k, K ~ 100 the chain cost 0.15 ms per pivot, all of it launch and drain
// latency, and it sits between two HBM-bound price-outs.  Here one persistent grid (one CTA per SM, launched
// cooperatively so that every CTA is resident) walks the same steps separated by grid-wide barriers (~2 us each).
//
// Arithmetic: every element-wise update is the one of the separate kernels, in the same order (acc -= t_j * M[i,j] with j
// ascending; chains of k_eta_scatter in descending eta index).  The reductions (small mat-vecs, transposed tall-skinny
// products, sums of squares) use fixed shapes that depend only on (m, k, K) and the SM count, so a result is
// reproducible and identical on every shard of a column-sharded engine.
//
// Memory model: data produced in one phase and consumed in a later one by OTHER CTAs is read with __ldcg (L2): L1 is
// not coherent across SMs and a plain load could hit a line cached in an earlier phase.  Inputs written before the
// launch (factors, eta file, index maps, the entering column) use ordinary loads.
#pragma once

constexpr int FZ_T = 512;     // threads per CTA
constexpr int FZ_WARPS = FZ_T / 32;
constexpr int FZ_MAX = 512;   // largest k and K handled here; beyond, the basis passes are bandwidth-bound (>= 200 MB each) and the
                              // separate kernels with their column-group splits take over.  A/B at 1024 (r02s3, config 3 pivots
                              // 2500-4500, k 710-844, K up to 756): 3.36-3.46 ms/pivot fused vs 3.35-3.42 separate — no gain, the deep
                              // regime is bandwidth-bound as a whole (profiles/r02_deep_curve_config3.md)
constexpr int FZ_G = 32;      // column groups of the small mat-vecs
constexpr int FZ_MAXS = 64;   // row slices of the transposed tall-skinny products
constexpr int FZ_SEG = 1024;  // compaction segment (= CP_SEG)

struct ChainArgs {
  int m, k, K, S_k, S_K;
  int64_t mld, kcap, Kcap;
  const double *Cinv, *Ginv, *Bcols, *E, *colq;
  const int32_t *Rp, *Jpos, *Jslot, *rowcover, *etaR, *etaPrev, *etaLast;
  double *alpha, *vvec, *cov;
  double *px, *pt, *pu, *pr;  // partials: FZ_G x k, FZ_G x K, S_K x K, S_k x k
  double *uK, *sK, *rk;       // K, K, k
  int32_t* cta_cnt;           // statistics of alpha per CTA
  double* cta_ss;
  int32_t* seg_cnt;           // statistics of v per 1024-row segment
  double* seg_ss;
  int32_t* vidx;
  double* vval;
  int32_t* icnt;  // [1] nnz(alpha) [2] nnz(v)
  double* scal;   // [2] |alpha|^2  [3] |v|^2
  unsigned* bar;
  int* flags;     // [2]: a grid barrier timed out
  const uint8_t* touched;  // stored positions of the newest eta (k_touch_mark in kernels_common.cuh)
  uint8_t* touched_new;    // touched U nz(alpha0): the stored size of col_coeffs is its population count -> icnt[1]
};

// Grid-wide barrier (the cooperative-groups scheme: CTA 0 adds the complement, the top bit flips when all have arrived).
// Bounded spin: if the grid is ever not co-resident the kernel gives up and raises flags[2] instead of hanging the GPU.
__device__ __forceinline__ bool fz_grid_bar(unsigned* bar, int* flags) {
  __shared__ int ok_s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned nb = 1;
    if (blockIdx.x == 0) nb = 0x80000000u - (gridDim.x - 1);
    __threadfence();
    const unsigned old = atomicAdd(bar, nb);
    const long long t0 = clock64();
    int ok = 1;
    while (((old ^ *((volatile unsigned*)bar)) & 0x80000000u) == 0) {
      if (clock64() - t0 > 4000000000LL) { ok = 0; flags[2] = 1; break; }  // ~2 s
    }
    __threadfence();
    ok_s = ok;
  }
  __syncthreads();
  return ok_s != 0;
}
#define FZ_BAR()                                   \
  do {                                             \
    if (!fz_grid_bar(a.bar, a.flags)) return;      \
  } while (0)

__device__ __forceinline__ int fz_block_sum_int(int v, int* sm) {  // result valid in thread 0
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULLMASK, v, o);
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  int r = 0;
  if (wid == 0) {
    r = lane < FZ_WARPS ? sm[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(FULLMASK, r, o);
  }
  return r;
}

// c[p] of BTRAN after the eta file (solver.rs:1325-1333): alpha[p] minus the coefficients of the etas whose leaving row
// is p, newest first — the chain k_eta_scatter walks.
__device__ __forceinline__ double fz_after_etas(const ChainArgs& a, int p) {
  double v = __ldcg(a.alpha + p);
  if (a.K > 0)
    for (int q = a.etaLast[p]; q >= 0; q = a.etaPrev[q]) v -= __ldcg(a.sK + q);
  return v;
}

// part[g][i] = sum_{j = g, g+G, ...; j < n (TRI: j <= i)} M[i + j ld] * x[gidx[j]]       warp = (32 rows) x (one column group)
template <bool TRI, bool XCG>
__device__ __forceinline__ void fz_mv_n_part(const double* __restrict__ M, int64_t ld, int n, const double* x,
                                             const int32_t* __restrict__ gidx, double* part, int gw, int W, int lane) {
  const int nrb = (n + 31) >> 5;
  for (int u = gw; u < nrb * FZ_G; u += W) {
    const int rb = u / FZ_G, g = u % FZ_G;
    const int i = rb * 32 + lane;
    if (i >= n) continue;
    const int jend = TRI ? i + 1 : n;
    const double* p = M + i;
    double acc = 0.0;
    int j = g;
    for (; j + 3 * FZ_G < jend; j += 4 * FZ_G) {
      double mv[4], xv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        mv[q] = p[(int64_t)(j + q * FZ_G) * ld];
        const double* xp = x + gidx[j + q * FZ_G];
        xv[q] = XCG ? __ldcg(xp) : *xp;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) acc += mv[q] * xv[q];
    }
    for (; j < jend; j += FZ_G) {
      const double* xp = x + gidx[j];
      acc += p[(int64_t)j * ld] * (XCG ? __ldcg(xp) : *xp);
    }
    part[(int64_t)g * n + i] = acc;
  }
}

// part[s][j] = sum over row slice s of M[i + col(j) ld] * x[i]      warp = one (column, slice); x read through L2
__device__ __forceinline__ void fz_mv_t_part(const double* __restrict__ M, int64_t ld, int rows, int cols, int S,
                                             const int32_t* __restrict__ slots, const double* x, double* part, int gw, int W,
                                             int lane) {
  const int L = (rows + S - 1) / S;
  for (int u = gw; u < cols * S; u += W) {
    const int j = u % cols, s = u / cols;
    const int r0 = s * L, r1 = min(rows, r0 + L);
    const double* p = M + (int64_t)(slots ? slots[j] : j) * ld;
    double acc = 0.0;
    int i = r0 + lane;
    for (; i + 7 * 32 < r1; i += 8 * 32) {
      double mv[8], xv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { mv[q] = p[i + q * 32]; xv[q] = __ldcg(x + i + q * 32); }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc += mv[q] * xv[q];
    }
    for (; i < r1; i += 32) acc += p[i] * __ldcg(x + i);
    acc = warp_sum(acc);
    if (lane == 0) part[(int64_t)s * cols + j] = acc;
  }
}

// y[i] -= sum_j t[j] * M[i + col(j) ld], j ascending (the loop of k_ftran_finish / k_gemv_n_sub)
__device__ __forceinline__ double fz_row_sub(double acc, const double* __restrict__ p, int64_t ld, int n, const double* ts,
                                             const int32_t* sl) {
  int j = 0;
  for (; j + 16 <= n; j += 16) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = p[(int64_t)(sl ? sl[j + u] : j + u) * ld];
#pragma unroll
    for (int u = 0; u < 16; ++u) acc -= ts[j + u] * v[u];
  }
  for (; j + 4 <= n; j += 4) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = p[(int64_t)(sl ? sl[j + u] : j + u) * ld];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc -= ts[j + u] * v[u];
  }
  for (; j < n; ++j) acc -= ts[j] * p[(int64_t)(sl ? sl[j] : j) * ld];
  return acc;
}

__global__ void __launch_bounds__(FZ_T, 1) k_chain_primal(ChainArgs a) {
  __shared__ double xs[FZ_MAX];
  __shared__ int32_t sl[FZ_MAX];
  __shared__ double smd[32];
  __shared__ int smi[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gw = blockIdx.x * FZ_WARPS + warp, W = gridDim.x * FZ_WARPS;
  const int m = a.m, k = a.k, K = a.K;
  int st_cnt = 0;       // statistics of the alpha entries this thread finalises
  double st_ss = 0.0;

  // ---- FTRAN ----------------------------------------------------------------------------------------------------
  // 1. x = C^-1 a_R (lu.rs:92-93 through the explicit inverse of the core): column-group partials
  if (k > 0) {
    fz_mv_n_part<false, false>(a.Cinv, a.kcap, k, a.colq, a.Rp, a.px, gw, W, lane);
    FZ_BAR();
  }
  // 2. alpha_slack = a_S - D[S,:] x, alpha[Jpos[t]] = x_t
  for (int j = tid; j < k; j += FZ_T) {
    double t = 0.0;
    for (int g = 0; g < FZ_G; ++g) t += __ldcg(a.px + (int64_t)g * k + j);
    xs[j] = t;
    sl[j] = a.Jslot[j];
  }
  __syncthreads();
  for (int i = blockIdx.x * FZ_T + tid; i < m; i += gridDim.x * FZ_T) {
    const int cv = a.rowcover[i];
    if (cv < 0) continue;
    const double acc = fz_row_sub(a.colq[i], a.Bcols + i, a.mld, k, xs, sl);
    a.alpha[cv] = acc;
    const int tch = (acc != 0.0) | (a.touched[cv] != 0);  // structural size of col_coeffs: see k_touch_mark
    a.touched_new[cv] = (uint8_t)tch;
    st_cnt += tch;
    if (K == 0) st_ss += acc * acc;
  }
  if (blockIdx.x == 0)
    for (int t = tid; t < k; t += FZ_T) {
      const double xv = xs[t];
      const int p = a.Jpos[t];
      a.alpha[p] = xv;
      const int tch = (xv != 0.0) | (a.touched[p] != 0);
      a.touched_new[p] = (uint8_t)tch;
      st_cnt += tch;
      if (K == 0) st_ss += xv * xv;
    }
  FZ_BAR();
  if (K > 0) {
    // 3. t = (I+G)^-1 alpha0[r]   (solver.rs:1310-1316 in closed form)
    fz_mv_n_part<true, true>(a.Ginv, a.Kcap, K, a.alpha, a.etaR, a.pt, gw, W, lane);
    FZ_BAR();
    // 4. alpha -= E t
    for (int j = tid; j < K; j += FZ_T) {
      double t = 0.0;
      for (int g = 0; g < FZ_G; ++g) t += __ldcg(a.pt + (int64_t)g * K + j);
      xs[j] = t;
    }
    __syncthreads();
    for (int i = blockIdx.x * FZ_T + tid; i < m; i += gridDim.x * FZ_T) {
      const double acc = fz_row_sub(__ldcg(a.alpha + i), a.E + i, a.mld, K, xs, nullptr);
      a.alpha[i] = acc;
      st_ss += acc * acc;
    }
  }
  {  // nnz(alpha), |alpha|^2: per-CTA partials, added in CTA order in step 10
    const int c = fz_block_sum_int(st_cnt, smi);
    const double s = block_sum(st_ss, smd);
    if (tid == 0) { a.cta_cnt[blockIdx.x] = c; a.cta_ss[blockIdx.x] = s; }
  }
  FZ_BAR();

  // ---- BTRAN of alpha -------------------------------------------------------------------------------------------
  if (K > 0) {
    // 5. u = E^T alpha (solver.rs:1326-1330)
    fz_mv_t_part(a.E, a.mld, m, K, a.S_K, nullptr, a.alpha, a.pu, gw, W, lane);
    FZ_BAR();
    for (int j = gw; j < K; j += W) {
      double v = 0.0;
      for (int s = lane; s < a.S_K; s += 32) v += __ldcg(a.pu + (int64_t)s * K + j);
      v = warp_sum(v);
      if (lane == 0) a.uK[j] = v;
    }
    FZ_BAR();
    // 6. s = (I+G)^-T u
    for (int j = gw; j < K; j += W) {
      const double* col = a.Ginv + (int64_t)j * a.Kcap;
      double acc = 0.0;
      for (int i = j + lane; i < K; i += 32) acc += col[i] * __ldcg(a.uK + i);
      acc = warp_sum(acc);
      if (lane == 0) a.sK[j] = acc;
    }
    FZ_BAR();
  }
  // 7. rho_S: v_i = c[position of slack i] on covered rows, 0 on the core's rows (k_btran_start)
  for (int i = blockIdx.x * FZ_T + tid; i < m; i += gridDim.x * FZ_T) {
    const int p = a.rowcover[i];
    const double v = p >= 0 ? fz_after_etas(a, p) : 0.0;
    a.vvec[i] = v;
    a.cov[i] = v;
  }
  FZ_BAR();
  if (k > 0) {
    // 8. right-hand side of the core solve: rhs_t = c[Jpos[t]] - D[S,t] . rho_S
    fz_mv_t_part(a.Bcols, a.mld, m, k, a.S_k, a.Jslot, a.cov, a.pr, gw, W, lane);
    FZ_BAR();
    for (int t = gw; t < k; t += W) {
      double v = 0.0;
      for (int s = lane; s < a.S_k; s += 32) v += __ldcg(a.pr + (int64_t)s * k + t);
      v = warp_sum(v);
      if (lane == 0) a.rk[t] = fz_after_etas(a, a.Jpos[t]) - v;
    }
    FZ_BAR();
    // 9. y = C^-T rhs, scattered to the core's constraint rows (lu_factors_transp, lu.rs:108-115)
    for (int j = gw; j < k; j += W) {
      const double* col = a.Cinv + (int64_t)j * a.kcap;
      double acc = 0.0;
      for (int i = lane; i < k; i += 32) acc += col[i] * __ldcg(a.rk + i);
      acc = warp_sum(acc);
      if (lane == 0) a.vvec[a.Rp[j]] = acc;
    }
    FZ_BAR();
  }

  // ---- ordered non-zero list of v, nnz / sums of squares ---------------------------------------------------------
  const int nseg = (m + FZ_SEG - 1) / FZ_SEG;
  // 10. per-segment counts
  for (int sg = blockIdx.x; sg < nseg; sg += gridDim.x) {
    const int i0 = sg * FZ_SEG + 2 * tid;
    const double v0 = i0 < m ? __ldcg(a.vvec + i0) : 0.0;
    const double v1 = i0 + 1 < m ? __ldcg(a.vvec + i0 + 1) : 0.0;
    const int c = fz_block_sum_int((v0 != 0.0) + (v1 != 0.0), smi);
    const double s = block_sum(v0 * v0 + v1 * v1, smd);
    if (tid == 0) { a.seg_cnt[sg] = c; a.seg_ss[sg] = s; }
  }
  if (blockIdx.x == gridDim.x - 1 && warp == 0) {  // statistics of alpha: CTA partials in CTA order
    int c = 0;
    double t = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) { c += __ldcg(a.cta_cnt + b); t += __ldcg(a.cta_ss + b); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULLMASK, c, o);
    t = warp_sum(t);
    if (lane == 0) { a.icnt[1] = c; a.scal[2] = t; }
  }
  FZ_BAR();
  // 11. write the list
  for (int sg = blockIdx.x; sg < nseg; sg += gridDim.x) {
    __syncthreads();
    if (warp == 0) {
      int acc = 0;
      for (int b = lane; b < sg; b += 32) acc += __ldcg(a.seg_cnt + b);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULLMASK, acc, o);
      if (lane == 0) s_base = acc;
    }
    const int i0 = sg * FZ_SEG + 2 * tid;
    const double v0 = i0 < m ? __ldcg(a.vvec + i0) : 0.0;
    const double v1 = i0 + 1 < m ? __ldcg(a.vvec + i0 + 1) : 0.0;
    const int c = (v0 != 0.0) + (v1 != 0.0);
    int incl = c;  // inclusive scan within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULLMASK, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) smi[warp] = incl;
    __syncthreads();
    int off = s_base + incl - c;
    for (int w2 = 0; w2 < warp; ++w2) off += smi[w2];
    if (v0 != 0.0) { a.vidx[off] = i0; a.vval[off] = v0; ++off; }
    if (v1 != 0.0) { a.vidx[off] = i0 + 1; a.vval[off] = v1; }
  }
  if (blockIdx.x == 0 && warp == 0) {
    int c = 0;
    double t = 0.0;
    for (int b = lane; b < nseg; b += 32) { c += __ldcg(a.seg_cnt + b); t += __ldcg(a.seg_ss + b); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULLMASK, c, o);
    t = warp_sum(t);
    if (lane == 0) { a.icnt[2] = c; a.scal[3] = t; }
  }
}
