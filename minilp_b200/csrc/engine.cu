// minilp_b200 engine: device-resident revised-simplex state and the sm_100a kernels of the pivot path.
// Reference items replaced (file:line under /root/reference/src) are cited at each kernel / entry point.
// Compiled with -fmad=false: the reference (Rust) never contracts a*b+c, and every element-wise update
// here reproduces the reference's operation order exactly; only reductions (dot products, sums of
// squares, arg-min/max scans) are evaluated in a different — tree / chunked — order.
#include "minilp_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }
extern "C" const char* mlp_last_error(void) { return g_err.c_str(); }
extern "C" const char* mlp_version(void) { return "minilp_b200 0.1 (sm_100a)"; }
extern "C" int mlp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

#define CU(x)                                                                                   \
  do {                                                                                          \
    cudaError_t err__ = (x);                                                                    \
    if (err__ != cudaSuccess) {                                                                 \
      set_err(std::string(#x) + ": " + cudaGetErrorString(err__) + " @" + std::to_string(__LINE__)); \
      return MLP_CUDA_ERROR;                                                                    \
    }                                                                                           \
  } while (0)
#define ST(x)                          \
  do {                                 \
    mlp_status st__ = (x);             \
    if (st__ != MLP_OK) return st__;   \
  } while (0)

static constexpr double EPS = 1e-8;  // solver.rs:12
#define FULLMASK 0xffffffffu

// ------------------------------------------------------------------------------------------------ engine
struct DevRes {  // small result block, mirrored in pinned host memory
  double f[8];
  long long i[6];
  int flags[4];  // [0] nonfinite, [1] singular
};

struct mlp_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t m = 0, n = 0, nt = 0, lda = 0, ldv = 0;  // ldv: padded length of var-indexed work arrays
  int sm_count = 148;
  bool initialized = false;
  int enable_pse = 0, enable_dse = 0;

  double* A = nullptr;                                    // m x lda row-major
  double *lo = nullptr, *hi = nullptr, *cobj = nullptr;  // n+m
  double *d = nullptr, *gam = nullptr, *xnb = nullptr;   // n+m
  uint8_t* vflag = nullptr;                               // n+m
  int32_t* vpos = nullptr;                                // n+m
  int32_t* bvar = nullptr;                                // m
  double *xB = nullptr, *loB = nullptr, *hiB = nullptr, *w = nullptr, *rhs = nullptr;  // m
  double *alpha = nullptr, *rho = nullptr, *tau = nullptr, *vvec = nullptr;             // m
  double *work_m = nullptr, *work_m2 = nullptr;                                          // m
  double *rc = nullptr, *helper = nullptr;                                               // ldv (n+m padded)
  int32_t* list_idx = nullptr;  // m
  double* list_val = nullptr;   // m
  double* partial = nullptr;    // price partial sums: max_chunks x lda
  int max_chunks = 64;
  // reduction scratch
  double* red_f = nullptr;      // 4096 doubles
  long long* red_i = nullptr;   // 4096
  unsigned* red_counter = nullptr;
  double* scal = nullptr;       // device scalars: [0] max_step [1] rho sumsq [2] alpha sumsq [3] v sumsq ...
  int32_t* icnt = nullptr;      // device ints: [0] rho nnz [1] alpha nnz [2] v nnz
  int32_t* seg_cnt = nullptr;   // compaction: per-segment counts / sums of squares
  double* seg_ss = nullptr;
  double *gt_part_k = nullptr, *gt_part_K = nullptr;  // k_gemv_t partials: GT_MAXSPLIT x kcap / Kcap
  DevRes* d_res = nullptr;
  DevRes* h_res = nullptr;  // pinned

  // dense LU of the basis
  int64_t k = 0, kcap = 0;
  int32_t *Jpos = nullptr, *Jvar = nullptr, *Rp = nullptr;  // kcap
  int32_t* rowcover = nullptr;                                // m
  double *Bcols = nullptr, *LUc = nullptr;                    // m x kcap, kcap x kcap
  double *xk = nullptr;                                       // kcap
  // eta file
  int64_t K = 0, Kcap = 0;
  double *E = nullptr, *G = nullptr;  // m x Kcap, Kcap x Kcap (col-major, unit lower coupling matrix)
  int32_t *etaR = nullptr, *etaPrev = nullptr, *etaHead = nullptr;  // Kcap
  double *tK = nullptr;               // Kcap
  int64_t lu_nnz = 0;

  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t pev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [slot][begin/end]; slot 0 rho, 1 v
  bool ppending[2] = {false, false};
  int prof_on = 0;
  mlp_profile prof{};

  std::vector<int32_t> h_bvar;
  std::vector<int32_t> h_last_eta_of_row;  // m, -1 if none
  mlp_counters cnt{};
};

#define LAUNCH(e, kern, grid, block, smem, ...)              \
  do {                                                       \
    kern<<<(grid), (block), (smem), (e)->stream>>>(__VA_ARGS__); \
    (e)->cnt.kernel_launches += 1;                           \
  } while (0)

template <class T> static mlp_status dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t err = cudaMalloc((void**)p, count * sizeof(T));
  if (err != cudaSuccess) {
    set_err(std::string("cudaMalloc: ") + cudaGetErrorString(err));
    return err == cudaErrorMemoryAllocation ? MLP_NOMEM : MLP_CUDA_ERROR;
  }
  return MLP_OK;
}
template <class T> static void dev_free(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}
static mlp_status h2d(mlp_engine* e, void* dst, const void* src, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream));
  e->cnt.h2d_bytes += (int64_t)bytes;
  return MLP_OK;
}
static mlp_status d2h(mlp_engine* e, void* dst, const void* src, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->cnt.d2h_bytes += (int64_t)bytes;
  return MLP_OK;
}
static mlp_status fetch_res(mlp_engine* e) { return d2h(e, e->h_res, e->d_res, sizeof(DevRes)); }

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

// ------------------------------------------------------------------------------------------------ K5 price-out
// calc_row_coeffs price-out (solver.rs:685-692), the N^T v product of update_primal_sq_norms (1117-1132)
// and the column norms of try_new (297-299).  Row-gather GEMV^T over row-major A: a CTA owns a 512-column
// tile and one chunk of the multiplier's support; each thread accumulates two adjacent columns with
// 128-bit loads, rows of the chunk are taken in list order (the reference's order of accumulation within
// the chunk); (row, weight) pairs are staged through shared memory.  Partial sums per chunk are reduced
// in chunk order by k_price_finish — no atomics, bitwise reproducible.
constexpr int PR_THREADS = 256;
constexpr int PR_TILE = PR_THREADS * 2;
constexpr int PR_BATCH = 256;
constexpr int PR_UNROLL = 8;

template <int MODE>  // 0: sum_r w_r * A[r,j]   1: sum_r A[r,j]^2
__global__ void __launch_bounds__(PR_THREADS)
k_price_partial(const double* __restrict__ A, int64_t lda, const int32_t* __restrict__ rows,
                const double* __restrict__ wts, const int32_t* __restrict__ count_ptr, int32_t fixed_count,
                double* __restrict__ partial) {
  __shared__ int32_t srow[PR_BATCH];
  __shared__ double sw[PR_BATCH];
  const int s = count_ptr ? *count_ptr : fixed_count;
  const int C = gridDim.y;
  const int L = (s + C - 1) / C;
  const int k0 = blockIdx.y * L;
  const int k1 = min(s, k0 + L);
  const int64_t col = ((int64_t)blockIdx.x * PR_THREADS + threadIdx.x) * 2;
  const bool active = col < lda;
  double acc0 = 0.0, acc1 = 0.0;
  const double* base = A + col;
  for (int kb = k0; kb < k1; kb += PR_BATCH) {
    const int nb = min(PR_BATCH, k1 - kb);
    __syncthreads();
    for (int t = threadIdx.x; t < nb; t += PR_THREADS) {
      srow[t] = rows ? rows[kb + t] : kb + t;
      sw[t] = (MODE == 0) ? wts[kb + t] : 1.0;
    }
    __syncthreads();
    if (active) {
      int i = 0;
      for (; i + PR_UNROLL <= nb; i += PR_UNROLL) {
        double2 v[PR_UNROLL];
#pragma unroll
        for (int u = 0; u < PR_UNROLL; ++u)
          v[u] = __ldcs(reinterpret_cast<const double2*>(base + (int64_t)srow[i + u] * lda));
#pragma unroll
        for (int u = 0; u < PR_UNROLL; ++u) {
          if (MODE == 0) {
            const double wv = sw[i + u];
            acc0 += wv * v[u].x;
            acc1 += wv * v[u].y;
          } else {
            acc0 += v[u].x * v[u].x;
            acc1 += v[u].y * v[u].y;
          }
        }
      }
      for (; i < nb; ++i) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(base + (int64_t)srow[i] * lda));
        if (MODE == 0) {
          const double wv = sw[i];
          acc0 += wv * v.x;
          acc1 += wv * v.y;
        } else {
          acc0 += v.x * v.x;
          acc1 += v.y * v.y;
        }
      }
    }
  }
  if (active) {
    double2 o;
    o.x = acc0;
    o.y = acc1;
    *reinterpret_cast<double2*>(partial + (int64_t)blockIdx.y * lda + col) = o;
  }
}

// Reduce chunk partials in chunk order; slack columns of [A|I] contribute rho_i (the `I` part of the CSR
// row, solver.rs:250); basic variables are not part of row_coeffs (solver.rs:688).
// mode 0: out = sum   mode 1: out = sum + 1 (primal edge norms, solver.rs:298)
__global__ void k_price_finish(const double* __restrict__ partial, int C, int64_t lda, int64_t n, int64_t m,
                               const double* __restrict__ slack_vals, const uint8_t* __restrict__ vflag,
                               double* __restrict__ out, int mode) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n + m) return;
  double r;
  if (v < n) {
    r = 0.0;
    for (int c = 0; c < C; ++c) r += partial[(int64_t)c * lda + v];
    if (mode == 1) r += 1.0;
  } else {
    r = (mode == 1) ? 2.0 : slack_vals[v - n];  // |e_i|^2 + 1
  }
  if (mode == 0 && (vflag[v] & MLP_BASIC)) r = 0.0;
  out[v] = r;
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
// rhs.set(column of var) (solver.rs:672-675, sparse.rs:103): column of [A|I] into a dense m-vector
__global__ void k_load_col(const double* __restrict__ A, int64_t lda, int64_t n, int m, int64_t var, double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  dst[i] = var < n ? A[(int64_t)i * lda + var] : ((int64_t)i == var - n ? 1.0 : 0.0);
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

// ------------------------------------------------------------------------------------------------ basis solve pieces
// FTRAN tail: alpha[pos] for slack positions = a_i - (D1 x)_i ; alpha[Jpos[t]] = x[t]   (U-solve of the
// identity-bordered basis, lu.rs:93 with B = [D | E_S]).
__global__ void __launch_bounds__(256) k_ftran_finish(const double* __restrict__ Bcols, int64_t ldb, int m, int k,
                                                       const double* __restrict__ xk, const double* __restrict__ rhs0,
                                                       const int32_t* __restrict__ rowcover, const int32_t* __restrict__ Jpos,
                                                       double* __restrict__ out) {
  __shared__ double ts[512];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = i < m ? rhs0[i] : 0.0;
  const int cov = i < m ? rowcover[i] : -1;
  for (int j0 = 0; j0 < k; j0 += 512) {
    const int nj = min(512, k - j0);
    __syncthreads();
    for (int q = threadIdx.x; q < nj; q += blockDim.x) ts[q] = xk[j0 + q];
    __syncthreads();
    if (cov >= 0) {
      const double* p = Bcols + (int64_t)j0 * ldb + i;
      int j = 0;
      for (; j + 8 <= nj; j += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[(int64_t)(j + u) * ldb];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc -= ts[j + u] * v[u];
      }
      for (; j < nj; ++j) acc -= ts[j] * p[(int64_t)j * ldb];
    }
  }
  if (cov >= 0) out[cov] = acc;
  if (i < k) out[Jpos[i]] = xk[i];
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
__global__ void k_gather_bcols(const double* __restrict__ A, int64_t lda, int m, int k, const int32_t* __restrict__ Jvar,
                               double* __restrict__ Bcols, int64_t ldb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (i < m && t < k) Bcols[(int64_t)t * ldb + i] = A[(int64_t)i * lda + Jvar[t]];
}
__global__ void k_extract_core(const double* __restrict__ Bcols, int64_t ldb, int k, const int32_t* __restrict__ Rp,
                               double* __restrict__ C, int64_t ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (i < k && t < k) C[(int64_t)t * ld + i] = Bcols[(int64_t)t * ldb + Rp[i]];
}

// ------------------------------------------------------------------------------------------------ K1 pricing scan
// choose_pivot, solver.rs:696-739: arg-max of d^2/gamma (or |d|) over eligible non-basic variables,
// strict '>' in ascending position order => lowest position wins ties.
__global__ void __launch_bounds__(256) k_select_primal(const double* __restrict__ d, const double* __restrict__ gam,
                                                        const uint8_t* __restrict__ vflag, const int32_t* __restrict__ vpos,
                                                        int64_t nt, int use_se, double* __restrict__ red_f,
                                                        long long* __restrict__ red_i, unsigned* counter,
                                                        const double* __restrict__ xnb, const double* __restrict__ lo,
                                                        const double* __restrict__ hi, DevRes* res) {
  __shared__ double smk[32];
  __shared__ long long smi[32];
  KeyIdx best{-INFINITY, LLONG_MAX};
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nt; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned f = vflag[v];
    if (f & MLP_BASIC) continue;
    const double dv = d[v];
    if (((f & MLP_AT_MIN) && dv > -EPS) || ((f & MLP_AT_MAX) && dv < EPS)) continue;  // 705-708
    const double score = use_se ? dv * dv / gam[v] : fabs(dv);
    // idx packs (pos, var): pos decides ties
    const long long key2 = ((long long)vpos[v] << 32) | (long long)v;
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
    if (b.idx == LLONG_MAX) { res->i[0] = -1; res->i[1] = -1; }
    else {
      const long long v = b.idx & 0xffffffffLL;
      res->i[0] = v;
      res->i[1] = b.idx >> 32;
      res->f[0] = d[v];
      res->f[1] = b.key;
      res->f[2] = xnb[v];
      res->f[3] = lo[v];
      res->f[4] = hi[v];
    }
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
                                                       double* __restrict__ red_f, unsigned* counter, double* __restrict__ scal) {
  __shared__ double sm[32];
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
__global__ void __launch_bounds__(256) k_ratio_dual_2(const double* __restrict__ rc, const double* __restrict__ d,
                                                       const uint8_t* __restrict__ vflag, const int32_t* __restrict__ vpos,
                                                       const double* __restrict__ xnb, int64_t nt, int lds,
                                                       const double* __restrict__ scal, double* __restrict__ red_f,
                                                       long long* __restrict__ red_i, unsigned* counter, DevRes* res) {
  __shared__ double smk[32];
  __shared__ long long smi[32];
  const double max_step = scal[0];
  KeyIdx best{-INFINITY, LLONG_MAX};
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nt; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned f = vflag[v];
    if (f & MLP_BASIC) continue;
    const double coeff = rc[v];
    if (!dual_eligible(coeff, f, lds)) continue;
    const double oc = clamp_obj(d[v], f);
    const double cur = fabs(oc) / fabs(coeff);  // 993
    if (cur <= max_step && better_max(fabs(coeff), v, best.key, best.idx)) { best.key = fabs(coeff); best.idx = v; }
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
      const long long v = b.idx;
      res->i[0] = v;
      res->i[1] = vpos[v];
      res->f[0] = rc[v];
      res->f[1] = d[v];
      res->f[2] = xnb[v];
    }
  }
}

// ------------------------------------------------------------------------------------------------ pivot updates
// Row half of Solver::pivot: basic values (solver.rs:1049-1055), dual steepest-edge norms
// (update_dual_sq_norms 1163-1173) and the new eta column (push_eta_matrix 1274-1284).
__global__ void __launch_bounds__(256) k_pivot_rows(const double* __restrict__ alpha, const double* __restrict__ tau,
                                                     double* __restrict__ xB, double* __restrict__ w, int m, int row,
                                                     double entering_new_val, double entering_diff, double coeff, int has_elem,
                                                     int dse, const double* __restrict__ scal, double* __restrict__ eta_col,
                                                     int* __restrict__ flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
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
// Variable half: reduced costs (solver.rs:1073-1080) and primal steepest-edge norms (1139-1150).
__global__ void __launch_bounds__(256) k_pivot_vars(double* __restrict__ d, double* __restrict__ gam,
                                                     const double* __restrict__ rc, const double* __restrict__ helper,
                                                     const uint8_t* __restrict__ vflag, int64_t nt, int64_t q, double coeff,
                                                     int pse, const double* __restrict__ scal, int* __restrict__ flags) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nt || v == q) return;
  if (vflag[v] & MLP_BASIC) return;
  const double c = rc[v];
  if (c == 0.0) return;
  const double pivot_obj = d[q] / coeff;  // 1073
  d[v] -= pivot_obj * c;
  if (pse) {
    const double psn = scal[2] + 1.0;  // 1136
    const double pcs = coeff * coeff;
    const double g = gam[v] + (-2.0 * c * helper[v] / coeff + psn * c * c / pcs);  // 1144-1146
    gam[v] = g;
    if (!isfinite(g)) flags[0] = 1;
  }
}
// Bookkeeping of Solver::pivot done by one thread: 1057-1058, 1066-1071, 1076, 1142, 1088-1091.
__global__ void k_pivot_swap(double* d, double* gam, double* xnb, uint8_t* vflag, int32_t* vpos, int32_t* bvar, double* loB,
                             double* hiB, const double* lo, const double* hi, int64_t q, int col, int row, double coeff,
                             double leaving_new_val, int pse, const double* scal, int* flags, DevRes* res) {
  const int lv = bvar[row];
  const double pivot_obj = d[q] / coeff;
  loB[row] = lo[q];
  hiB[row] = hi[q];
  xnb[lv] = leaving_new_val;
  unsigned f = 0;  // nb_var_is_fixed stays with the non-basic position in the reference and is false on this path
  if (leaving_new_val == lo[lv]) f |= MLP_AT_MIN;
  if (leaving_new_val == hi[lv]) f |= MLP_AT_MAX;
  vflag[lv] = (uint8_t)f;
  vpos[lv] = col;
  d[lv] = -pivot_obj;
  if (pse) {
    const double g = (scal[2] + 1.0) / (coeff * coeff);
    gam[lv] = g;
    if (!isfinite(g)) flags[0] = 1;
  }
  bvar[row] = (int32_t)q;
  vflag[q] = MLP_BASIC;
  vpos[q] = row;
  res->i[0] = lv;
  res->flags[0] = flags[0];
  res->flags[1] = flags[1];
}
__global__ void k_flip_var(double* xnb, uint8_t* vflag, const double* lo, const double* hi, int64_t q, double new_val) {
  xnb[q] = new_val;
  unsigned f = vflag[q] & MLP_FIXED;
  if (new_val == lo[q]) f |= MLP_AT_MIN;
  if (new_val == hi[q]) f |= MLP_AT_MAX;
  vflag[q] = (uint8_t)f;  // solver.rs:1038-1040
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

// ------------------------------------------------------------------------------------------------ init kernels
// basic_var_vals = rhs - A x_N at the initial point (solver.rs:234-238). One CTA per row.
__global__ void __launch_bounds__(256) k_init_basic_vals(const double* __restrict__ A, int64_t lda, int64_t n,
                                                          const double* __restrict__ xnb, const double* __restrict__ rhs,
                                                          double* __restrict__ xB) {
  __shared__ double sm[32];
  const int r = blockIdx.x;
  const double* row = A + (int64_t)r * lda;
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) acc += row[j] * xnb[j];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) xB[r] = rhs[r] - tot;
}
// d_N = c_N - N^T y (recalc_obj_coeffs, solver.rs:1216-1222)
__global__ void k_recalc_d(const double* __restrict__ cobj, const double* __restrict__ rc, const uint8_t* __restrict__ vflag,
                           int64_t nt, double* __restrict__ d) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nt || (vflag[v] & MLP_BASIC)) return;
  d[v] = cobj[v] - rc[v];
}
// objective from scratch (solver.rs:1224-1230): basic part in row order, then non-basic part; single CTA
__global__ void __launch_bounds__(1024) k_recalc_obj(const double* __restrict__ cobj, const int32_t* __restrict__ bvar,
                                                      const double* __restrict__ xB, int m, const double* __restrict__ xnb,
                                                      const uint8_t* __restrict__ vflag, int64_t nt, DevRes* res) {
  __shared__ double sm[32];
  double acc = 0.0;
  for (int r = threadIdx.x; r < m; r += blockDim.x) acc += cobj[bvar[r]] * xB[r];
  for (int64_t v = threadIdx.x; v < nt; v += blockDim.x)
    if (!(vflag[v] & MLP_BASIC)) acc += cobj[v] * xnb[v];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) res->f[0] = tot;
}
__global__ void k_gather_cB(const double* __restrict__ cobj, const int32_t* __restrict__ bvar, int m, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m) out[r] = cobj[bvar[r]];
}

// ================================================================================================ host side
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

static int price_chunks(const mlp_engine* e) {
  const int tiles = cdiv(e->lda, PR_TILE);
  int c = cdiv((int64_t)e->sm_count * 8, tiles);
  return std::max(1, std::min(c, e->max_chunks));
}

// out (var-indexed) = N^T w over the listed rows (+ slack part), basic entries zeroed
static mlp_status price_list(mlp_engine* e, const int32_t* rows, const double* wts, const int32_t* count_ptr, int fixed_count,
                             const double* slack_vals, double* out, int prof_slot = -1) {
  const int C = price_chunks(e);
  dim3 grid(cdiv(e->lda, PR_TILE), C);
  const bool prof = e->prof_on && prof_slot >= 0;
  if (prof) CU(cudaEventRecord(e->pev[prof_slot][0], e->stream));
  LAUNCH(e, k_price_partial<0>, grid, PR_THREADS, 0, e->A, e->lda, rows, wts, count_ptr, fixed_count, e->partial);
  LAUNCH(e, k_price_finish, cdiv(e->nt, 256), 256, 0, e->partial, C, e->lda, e->n, e->m, slack_vals, e->vflag, out, 0);
  if (prof) {
    CU(cudaEventRecord(e->pev[prof_slot][1], e->stream));
    e->ppending[prof_slot] = true;
  }
  return MLP_OK;
}
// after a stream sync: fold pending price-out timings into the profile. support sizes come from h_res->i[2..3].
static mlp_status collect_profile(mlp_engine* e, int64_t s_rho, int64_t s_v) {
  for (int slot = 0; slot < 2; ++slot) {
    if (!e->ppending[slot]) continue;
    e->ppending[slot] = false;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e->pev[slot][0], e->pev[slot][1]));
    const int64_t sz = slot == 0 ? s_rho : s_v;
    const int64_t bytes = 8 * e->n * sz + 8 * sz + 8 * e->n;
    if (slot == 0) { e->prof.price_rho_ms += ms; e->prof.price_rho_launches += 1; e->prof.price_rho_bytes += bytes; }
    else { e->prof.price_v_ms += ms; e->prof.price_v_launches += 1; e->prof.price_v_bytes += bytes; }
  }
  return MLP_OK;
}

// list / stats of a dense m-vector (see k_compact_count). idx == nullptr: stats only.
static void compact(mlp_engine* e, const double* x, int32_t* idx, double* val, int32_t* count, double* sumsq) {
  const int m = (int)e->m, nseg = cdiv(m, CP_SEG);
  LAUNCH(e, k_compact_count, nseg, CP_SEG, 0, x, m, e->seg_cnt, e->seg_ss, e->red_counter, count, sumsq);
  if (idx) LAUNCH(e, k_compact_write, nseg, CP_SEG, 0, x, m, e->seg_cnt, idx, val);
}
static void gemv_t(mlp_engine* e, const double* M, int64_t ld, int rows, int cols, const double* x, double* part,
                   const double* base, const int32_t* base_idx, double* out, int negate) {
  if (cols <= 0) return;
  int S = std::max(1, std::min(GT_MAXSPLIT, cdiv(2 * (int64_t)e->sm_count, cols)));
  S = std::min(S, std::max(1, rows / 2048));
  LAUNCH(e, k_gemv_t_part, dim3((unsigned)cols, (unsigned)S), 256, 0, M, ld, rows, cols, x, part);
  LAUNCH(e, k_gemv_t_fin, cdiv(cols, 256), 256, 0, part, S, cols, base, base_idx, out, negate);
}

template <bool FWD, bool AXPY, bool UNIT> static void trsv(mlp_engine* e, const double* M, int64_t ld, int n, double* x) {
  if (n <= 0) return;
  auto kern = k_trsv<FWD, AXPY, UNIT>;
  LAUNCH(e, kern, 1, 1024, 0, M, ld, n, x);
}

// BasisSolver::solve (solver.rs:1305-1319). rhs0: dense m-vector by constraint row (device). out: by basis position.
static mlp_status ftran(mlp_engine* e, const double* rhs0, double* out) {
  const int m = (int)e->m, k = (int)e->k, K = (int)e->K;
  if (k > 0) {
    LAUNCH(e, k_gather_idx, cdiv(k, 256), 256, 0, rhs0, e->Rp, k, e->xk);
    trsv<true, true, true>(e, e->LUc, e->kcap, k, e->xk);    // L y = P a_R   (lu.rs:92)
    trsv<false, true, false>(e, e->LUc, e->kcap, k, e->xk);  // U x = y      (lu.rs:93)
  }
  LAUNCH(e, k_ftran_finish, cdiv(std::max(m, k), 256), 256, 0, e->Bcols, e->m, m, k, e->xk, rhs0, e->rowcover, e->Jpos, out);
  if (K > 0) {  // eta file, solver.rs:1310-1316 in closed form
    LAUNCH(e, k_gather_idx, cdiv(K, 256), 256, 0, out, e->etaR, K, e->tK);
    trsv<true, true, true>(e, e->G, e->Kcap, K, e->tK);
    LAUNCH(e, k_gemv_n_sub, cdiv(m, 256), 256, 0, e->E, e->m, m, K, e->tK, out);
  }
  return MLP_OK;
}

// BasisSolver::solve_transp (solver.rs:1322-1338). c: dense m-vector by basis position (device, DESTROYED).
// unit_row >= 0 tells that c == e_unit_row (the eta dot products degenerate to a row gather). out: by constraint row.
static mlp_status btran(mlp_engine* e, double* c, int unit_row, double* out) {
  const int m = (int)e->m, k = (int)e->k, K = (int)e->K;
  if (K > 0) {  // etas in reverse, 1325-1333
    if (unit_row >= 0) LAUNCH(e, k_gather_row, cdiv(K, 256), 256, 0, e->E, e->m, unit_row, K, e->tK);
    else gemv_t(e, e->E, e->m, m, K, c, e->gt_part_K, nullptr, nullptr, e->tK, 0);
    trsv<false, false, true>(e, e->G, e->Kcap, K, e->tK);
    LAUNCH(e, k_eta_scatter, cdiv(K, 256), 256, 0, e->tK, e->etaR, e->etaPrev, e->etaHead, K, c);
  }
  LAUNCH(e, k_btran_start, cdiv(m, 256), 256, 0, c, e->rowcover, m, out, e->work_m2);
  if (k > 0) {
    gemv_t(e, e->Bcols, e->m, m, k, e->work_m2, e->gt_part_k, c, e->Jpos, e->xk, 1);
    trsv<true, false, false>(e, e->LUc, e->kcap, k, e->xk);  // U^T z = rhs   (lu_factors_transp.lower = U^T, lu.rs:110)
    trsv<false, false, true>(e, e->LUc, e->kcap, k, e->xk);  // L^T y = z
    LAUNCH(e, k_scatter_idx, cdiv(k, 256), 256, 0, e->xk, e->Rp, k, out);
  }
  return MLP_OK;
}

static mlp_status ensure_lu_capacity(mlp_engine* e, int64_t k) {
  if (k <= e->kcap && e->Bcols) return MLP_OK;
  // first allocation is generous (up to ~1 GB for the basis columns): cudaFree/cudaMalloc of the big arenas
  // costs tens of milliseconds, so capacity grows by doubling and rarely
  int64_t cap = std::max<int64_t>(e->kcap, std::min<int64_t>(e->m, std::max<int64_t>(64, std::min<int64_t>(1024, (1ll << 30) / (8 * e->m)))));
  while (cap < k) cap *= 2;
  cap = std::min<int64_t>(cap, e->m);
  dev_free(e->Jpos); dev_free(e->Jvar); dev_free(e->Rp); dev_free(e->Bcols); dev_free(e->LUc); dev_free(e->xk);
  dev_free(e->gt_part_k);
  e->kcap = cap;
  ST(dev_alloc(&e->gt_part_k, (size_t)GT_MAXSPLIT * cap));
  ST(dev_alloc(&e->Jpos, cap)); ST(dev_alloc(&e->Jvar, cap)); ST(dev_alloc(&e->Rp, cap));
  ST(dev_alloc(&e->Bcols, (size_t)e->m * cap)); ST(dev_alloc(&e->LUc, (size_t)cap * cap)); ST(dev_alloc(&e->xk, cap));
  return MLP_OK;
}
static mlp_status ensure_eta_capacity(mlp_engine* e, int64_t K) {
  if (K <= e->Kcap && e->E) return MLP_OK;
  int64_t cap = std::max<int64_t>(e->Kcap, std::max<int64_t>(96, std::min<int64_t>(2080, (2ll << 30) / (8 * e->m))));
  while (cap < K) cap *= 2;
  dev_free(e->E); dev_free(e->G); dev_free(e->etaR); dev_free(e->etaPrev); dev_free(e->etaHead); dev_free(e->tK);
  dev_free(e->gt_part_K);
  e->Kcap = cap;
  ST(dev_alloc(&e->gt_part_K, (size_t)GT_MAXSPLIT * cap));
  ST(dev_alloc(&e->E, (size_t)e->m * cap)); ST(dev_alloc(&e->G, (size_t)cap * cap));
  ST(dev_alloc(&e->etaR, cap)); ST(dev_alloc(&e->etaPrev, cap)); ST(dev_alloc(&e->etaHead, cap)); ST(dev_alloc(&e->tK, cap));
  return MLP_OK;
}

static mlp_status refactor_impl(mlp_engine* e) {
  const int64_t m = e->m, n = e->n;
  std::vector<int32_t> jpos, jvar, rowcover(m, -1), R;
  for (int64_t p = 0; p < m; ++p) {
    const int32_t v = e->h_bvar[p];
    if (v < n) { jpos.push_back((int32_t)p); jvar.push_back(v); }
    else rowcover[v - n] = (int32_t)p;
  }
  for (int64_t i = 0; i < m; ++i) if (rowcover[i] < 0) R.push_back((int32_t)i);
  const int64_t k = (int64_t)jpos.size();
  if ((int64_t)R.size() != k) { set_err("refactor: basis bookkeeping inconsistent"); return MLP_INVALID; }
  ST(ensure_lu_capacity(e, k));
  // eta arena: the reference allows eta nnz up to lu nnz (solver.rs:1096-1097) ~ (k+1) dense columns
  ST(ensure_eta_capacity(e, 2 * k + 32));
  e->k = k;
  e->K = 0;
  CU(cudaMemsetAsync(e->d_res->flags, 0, 4 * sizeof(int), e->stream));
  std::fill(e->h_last_eta_of_row.begin(), e->h_last_eta_of_row.end(), -1);
  ST(h2d(e, e->rowcover, rowcover.data(), m * sizeof(int32_t)));
  if (k > 0) {
    ST(h2d(e, e->Jpos, jpos.data(), k * sizeof(int32_t)));
    ST(h2d(e, e->Jvar, jvar.data(), k * sizeof(int32_t)));
    ST(h2d(e, e->Rp, R.data(), k * sizeof(int32_t)));
    CU(cudaStreamSynchronize(e->stream));  // host vectors go out of scope
    LAUNCH(e, k_gather_bcols, dim3(cdiv(m, 256), (unsigned)k), 256, 0, e->A, e->lda, (int)m, (int)k, e->Jvar, e->Bcols, e->m);
    LAUNCH(e, k_extract_core, dim3(cdiv(k, 256), (unsigned)k), 256, 0, e->Bcols, e->m, (int)k, e->Rp, e->LUc, e->kcap);
    int* flags = e->d_res->flags;
    for (int t = 0; t < (int)k; ++t) {
      LAUNCH(e, k_lu_pivot, 1, 1024, 0, e->LUc, e->kcap, (int)k, t, e->Rp, flags);
      const int rem = (int)k - t - 1;
      if (rem > 0) LAUNCH(e, k_lu_update, dim3(cdiv(rem, 32), cdiv(rem, 8)), dim3(32, 8), 0, e->LUc, e->kcap, (int)k, t, flags);
    }
    ST(fetch_res(e));
    if (e->h_res->flags[1]) { set_err("singular basis"); return MLP_SINGULAR; }
  } else {
    CU(cudaStreamSynchronize(e->stream));
  }
  // LUFactors::nnz (lu.rs:52-54) of the reference's factors of this basis when A is fully dense:
  // L: k(k-1)/2, U: (m-k)k + k(k-1)/2, plus m.
  e->lu_nnz = k * (k - 1) + (m - k) * k + m;
  e->cnt.refactors += 1;
  e->cnt.k_structural = k;
  return MLP_OK;
}

extern "C" {

mlp_status mlp_engine_create_dense(int device, int64_t m, int64_t n, mlp_engine** out) {
  *out = nullptr;
  if (m <= 0 || n <= 0 || m > 0x7fffffff || n + m > 0x7fffffff) { set_err("bad dimensions"); return MLP_INVALID; }
  if (mlp_device_count() <= device) { set_err("no CUDA device: the engine has no CPU fallback"); return MLP_NO_DEVICE; }
  CU(cudaSetDevice(device));
  mlp_engine* e = new mlp_engine();
  e->device = device;
  e->m = m; e->n = n; e->nt = n + m;
  e->lda = (n + 15) / 16 * 16;
  e->ldv = e->nt;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  e->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  mlp_status st = MLP_OK;
  auto A = [&](mlp_status s) { if (st == MLP_OK) st = s; };
  A(dev_alloc(&e->A, (size_t)m * e->lda));
  A(dev_alloc(&e->lo, e->nt)); A(dev_alloc(&e->hi, e->nt)); A(dev_alloc(&e->cobj, e->nt));
  A(dev_alloc(&e->d, e->nt)); A(dev_alloc(&e->gam, e->nt)); A(dev_alloc(&e->xnb, e->nt));
  A(dev_alloc(&e->vflag, e->nt)); A(dev_alloc(&e->vpos, e->nt)); A(dev_alloc(&e->bvar, m));
  A(dev_alloc(&e->xB, m)); A(dev_alloc(&e->loB, m)); A(dev_alloc(&e->hiB, m)); A(dev_alloc(&e->w, m)); A(dev_alloc(&e->rhs, m));
  A(dev_alloc(&e->alpha, m)); A(dev_alloc(&e->rho, m)); A(dev_alloc(&e->tau, m)); A(dev_alloc(&e->vvec, m));
  A(dev_alloc(&e->work_m, m)); A(dev_alloc(&e->work_m2, m));
  A(dev_alloc(&e->rc, e->nt)); A(dev_alloc(&e->helper, e->nt));
  A(dev_alloc(&e->list_idx, m)); A(dev_alloc(&e->list_val, m));
  A(dev_alloc(&e->partial, (size_t)e->max_chunks * e->lda));
  A(dev_alloc(&e->red_f, 4096)); A(dev_alloc(&e->red_i, 4096)); A(dev_alloc(&e->red_counter, 4));
  A(dev_alloc(&e->scal, 16)); A(dev_alloc(&e->icnt, 16));
  A(dev_alloc(&e->seg_cnt, (size_t)cdiv(m, CP_SEG) + 1)); A(dev_alloc(&e->seg_ss, (size_t)cdiv(m, CP_SEG) + 1)); A(dev_alloc(&e->d_res, 1)); A(dev_alloc(&e->rowcover, m));
  if (st != MLP_OK) { mlp_engine_destroy(e); return st; }
  CU(cudaHostAlloc((void**)&e->h_res, sizeof(DevRes), cudaHostAllocDefault));
  for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&e->ev[i]));
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) CU(cudaEventCreate(&e->pev[i][j]));
  CU(cudaMemsetAsync(e->A, 0, (size_t)m * e->lda * sizeof(double), e->stream));
  CU(cudaMemsetAsync(e->red_counter, 0, 4 * sizeof(unsigned), e->stream));
  CU(cudaMemsetAsync(e->d_res, 0, sizeof(DevRes), e->stream));
  CU(cudaMemsetAsync(e->gam, 0, e->nt * sizeof(double), e->stream));
  CU(cudaMemsetAsync(e->helper, 0, e->nt * sizeof(double), e->stream));
  CU(cudaMemsetAsync(e->rc, 0, e->nt * sizeof(double), e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->h_bvar.assign(m, 0);
  e->h_last_eta_of_row.assign(m, -1);
  *out = e;
  return MLP_OK;
}

void mlp_engine_destroy(mlp_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  dev_free(e->A); dev_free(e->lo); dev_free(e->hi); dev_free(e->cobj); dev_free(e->d); dev_free(e->gam); dev_free(e->xnb);
  dev_free(e->vflag); dev_free(e->vpos); dev_free(e->bvar); dev_free(e->xB); dev_free(e->loB); dev_free(e->hiB); dev_free(e->w);
  dev_free(e->rhs); dev_free(e->alpha); dev_free(e->rho); dev_free(e->tau); dev_free(e->vvec); dev_free(e->work_m);
  dev_free(e->work_m2); dev_free(e->rc); dev_free(e->helper); dev_free(e->list_idx); dev_free(e->list_val); dev_free(e->partial);
  dev_free(e->red_f); dev_free(e->red_i); dev_free(e->red_counter); dev_free(e->scal); dev_free(e->icnt); dev_free(e->d_res);
  dev_free(e->rowcover); dev_free(e->Jpos); dev_free(e->Jvar); dev_free(e->Rp); dev_free(e->Bcols); dev_free(e->LUc); dev_free(e->xk);
  dev_free(e->E); dev_free(e->G); dev_free(e->etaR); dev_free(e->etaPrev); dev_free(e->etaHead); dev_free(e->tK);
  dev_free(e->seg_cnt); dev_free(e->seg_ss); dev_free(e->gt_part_k); dev_free(e->gt_part_K);
  if (e->h_res) cudaFreeHost(e->h_res);
  for (int i = 0; i < 4; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) if (e->pev[i][j]) cudaEventDestroy(e->pev[i][j]);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

mlp_status mlp_engine_upload_rows(mlp_engine* e, int64_t row0, int64_t nrows, const double* rows_host) {
  if (!e || row0 < 0 || nrows < 0 || row0 + nrows > e->m) { set_err("upload_rows: range"); return MLP_INVALID; }
  CU(cudaSetDevice(e->device));
  CU(cudaMemcpy2DAsync(e->A + row0 * e->lda, e->lda * sizeof(double), rows_host, e->n * sizeof(double), e->n * sizeof(double),
                       (size_t)nrows, cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->cnt.h2d_bytes += nrows * e->n * (int64_t)sizeof(double);
  return MLP_OK;
}

mlp_status mlp_engine_init_state(mlp_engine* e, const mlp_init_state* st) {
  if (!e || !st) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int64_t m = e->m, n = e->n, nt = e->nt;
  ST(h2d(e, e->lo, st->orig_var_mins, nt * 8));
  ST(h2d(e, e->hi, st->orig_var_maxs, nt * 8));
  ST(h2d(e, e->cobj, st->orig_obj_coeffs, nt * 8));
  ST(h2d(e, e->rhs, st->orig_rhs, m * 8));
  std::vector<double> d(nt, 0.0), xnb(nt, 0.0), gam(nt, 0.0);
  std::vector<uint8_t> fl(nt, 0);
  std::vector<int32_t> pos(nt, 0), bvar(m);
  for (int64_t c = 0; c < n; ++c) {
    const int64_t v = st->nb_vars[c];
    if (v < 0 || v >= nt) { set_err("init_state: nb_vars"); return MLP_INVALID; }
    d[v] = st->nb_var_obj_coeffs[c];
    xnb[v] = st->nb_var_vals[c];
    fl[v] = st->nb_var_states[c] & (MLP_AT_MIN | MLP_AT_MAX | MLP_FIXED);
    pos[v] = (int32_t)c;
    if (st->primal_edge_sq_norms) gam[v] = st->primal_edge_sq_norms[c];
  }
  for (int64_t r = 0; r < m; ++r) {
    const int64_t v = st->basic_vars[r];
    if (v < 0 || v >= nt) { set_err("init_state: basic_vars"); return MLP_INVALID; }
    fl[v] = MLP_BASIC;
    pos[v] = (int32_t)r;
    bvar[r] = (int32_t)v;
  }
  e->h_bvar = bvar;
  e->enable_pse = st->enable_primal_steepest_edge;
  e->enable_dse = st->enable_dual_steepest_edge;
  ST(h2d(e, e->d, d.data(), nt * 8));
  ST(h2d(e, e->xnb, xnb.data(), nt * 8));
  ST(h2d(e, e->gam, gam.data(), nt * 8));
  ST(h2d(e, e->vflag, fl.data(), nt));
  ST(h2d(e, e->vpos, pos.data(), nt * 4));
  ST(h2d(e, e->bvar, bvar.data(), m * 4));
  ST(h2d(e, e->loB, st->basic_var_mins, m * 8));
  ST(h2d(e, e->hiB, st->basic_var_maxs, m * 8));
  if (st->basic_var_vals) ST(h2d(e, e->xB, st->basic_var_vals, m * 8));
  if (st->dual_edge_sq_norms) ST(h2d(e, e->w, st->dual_edge_sq_norms, m * 8));
  CU(cudaStreamSynchronize(e->stream));
  if (!st->basic_var_vals) LAUNCH(e, k_init_basic_vals, (unsigned)m, 256, 0, e->A, e->lda, n, e->xnb, e->rhs, e->xB);
  if (!st->dual_edge_sq_norms) LAUNCH(e, k_fill, cdiv(m, 256), 256, 0, e->w, m, 1.0);
  if (e->enable_pse && !st->primal_edge_sq_norms) {
    // |a_j|^2 + 1 (solver.rs:297-299): all m rows, unit weights
    const int C = price_chunks(e);
    dim3 grid(cdiv(e->lda, PR_TILE), C);
    LAUNCH(e, k_price_partial<1>, grid, PR_THREADS, 0, e->A, e->lda, (const int32_t*)nullptr, (const double*)nullptr,
           (const int32_t*)nullptr, (int32_t)m, e->partial);
    LAUNCH(e, k_price_finish, cdiv(nt, 256), 256, 0, e->partial, C, e->lda, n, m, (const double*)nullptr, e->vflag, e->gam, 1);
  }
  e->initialized = true;
  ST(refactor_impl(e));
  return MLP_OK;
}

mlp_status mlp_engine_set_primal_steepest_edge(mlp_engine* e, int32_t enable) {
  if (!e) return MLP_INVALID;
  e->enable_pse = enable;
  return MLP_OK;
}

mlp_status mlp_refactor(mlp_engine* e, int64_t* lu_nnz) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  ST(refactor_impl(e));
  if (lu_nnz) *lu_nnz = e->lu_nnz;
  return MLP_OK;
}

mlp_status mlp_select_entering_primal(mlp_engine* e, mlp_entering* out) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int grid = std::min(cdiv(e->nt, 256), 1024);
  LAUNCH(e, k_select_primal, grid, 256, 0, e->d, e->gam, e->vflag, e->vpos, e->nt, e->enable_pse, e->red_f, e->red_i,
         e->red_counter, e->xnb, e->lo, e->hi, e->d_res);
  ST(fetch_res(e));
  out->var = e->h_res->i[0];
  out->pos = e->h_res->i[1];
  out->obj_coeff = e->h_res->f[0];
  out->score = e->h_res->f[1];
  out->cur_val = e->h_res->f[2];
  out->var_min = e->h_res->f[3];
  out->var_max = e->h_res->f[4];
  return MLP_OK;
}

mlp_status mlp_ftran_col(mlp_engine* e, int64_t var) {
  if (!e || !e->initialized || var < 0 || var >= e->nt) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  LAUNCH(e, k_load_col, cdiv(e->m, 256), 256, 0, e->A, e->lda, e->n, (int)e->m, var, e->work_m);
  ST(ftran(e, e->work_m, e->alpha));
  // |alpha|^2 and nnz(alpha) for update_primal_sq_norms (1136) and the eta bookkeeping
  compact(e, e->alpha, nullptr, nullptr, e->icnt + 1, e->scal + 2);
  return MLP_OK;
}

mlp_status mlp_ratio_primal(mlp_engine* e, int32_t sign, double max_step0, mlp_leaving* out) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int grid = std::min(cdiv(e->m, 256), 1024);
  LAUNCH(e, k_ratio_primal_1, grid, 256, 0, e->alpha, e->xB, e->loB, e->hiB, (int)e->m, sign, max_step0, e->red_f,
         e->red_counter, e->scal);
  LAUNCH(e, k_ratio_primal_2, grid, 256, 0, e->alpha, e->xB, e->loB, e->hiB, (int)e->m, sign, e->scal, e->red_f, e->red_i,
         e->red_counter, e->d_res);
  ST(fetch_res(e));
  out->row = e->h_res->i[0];
  out->coeff = e->h_res->f[0];
  out->leaving_new_val = e->h_res->f[1];
  out->basic_val = e->h_res->f[2];
  return MLP_OK;
}

mlp_status mlp_btran_unit(mlp_engine* e, int64_t row) {
  if (!e || !e->initialized || row < 0 || row >= e->m) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  LAUNCH(e, k_set_unit, cdiv(e->m, 256), 256, 0, e->work_m, e->m, row);
  ST(btran(e, e->work_m, (int)row, e->rho));
  // inv_basis_row_coeffs as a sparse list + |rho|^2 (solver.rs:683, 1160)
  compact(e, e->rho, e->list_idx, e->list_val, e->icnt, e->scal + 1);
  return MLP_OK;
}

mlp_status mlp_price_row(mlp_engine* e) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  return price_list(e, e->list_idx, e->list_val, e->icnt, 0, e->rho, e->rc, 0);
}

mlp_status mlp_calc_row_coeffs(mlp_engine* e, int64_t row) {
  ST(mlp_btran_unit(e, row));
  return mlp_price_row(e);
}

mlp_status mlp_select_row_dual(mlp_engine* e, mlp_dual_row* out) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int grid = std::min(cdiv(e->m, 256), 1024);
  LAUNCH(e, k_select_row_dual, grid, 256, 0, e->xB, e->loB, e->hiB, e->w, (int)e->m, e->enable_dse, e->red_f, e->red_i,
         e->red_counter, e->d_res);
  ST(fetch_res(e));
  out->row = e->h_res->i[0];
  out->val = e->h_res->f[0];
  out->min = e->h_res->f[1];
  out->max = e->h_res->f[2];
  return MLP_OK;
}

mlp_status mlp_ratio_dual(mlp_engine* e, int64_t row, double leaving_new_val, mlp_dual_entering* out) {
  if (!e || !e->initialized || row < 0 || row >= e->m) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  // leaving_diff_sign = leaving_new_val > basic_var_vals[row] (solver.rs:925): the host knows basic_val from
  // select_row_dual, but fix_var-style callers may not; read it back (8 bytes).
  double bv = 0.0;
  ST(d2h(e, &bv, e->xB + row, sizeof(double)));
  const int lds = leaving_new_val > bv ? 1 : 0;
  const int grid = std::min(cdiv(e->nt, 256), 1024);
  LAUNCH(e, k_ratio_dual_1, grid, 256, 0, e->rc, e->d, e->vflag, e->nt, lds, e->red_f, e->red_counter, e->scal);
  LAUNCH(e, k_ratio_dual_2, grid, 256, 0, e->rc, e->d, e->vflag, e->vpos, e->xnb, e->nt, lds, e->scal, e->red_f, e->red_i,
         e->red_counter, e->d_res);
  ST(fetch_res(e));
  out->var = e->h_res->i[0];
  out->pos = e->h_res->i[1];
  out->coeff = e->h_res->f[0];
  out->obj_coeff = e->h_res->f[1];
  out->cur_val = e->h_res->f[2];
  return MLP_OK;
}

mlp_status mlp_pivot(mlp_engine* e, const mlp_pivot_info* pi, mlp_pivot_result* out) {
  if (!e || !e->initialized || !pi || !out) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int m = (int)e->m;
  const int64_t q = pi->entering_var;
  out->leaving_var = -1;
  out->col_nnz = 0;
  out->refactored = 0;
  out->lu_nnz = e->lu_nnz;
  CU(cudaMemsetAsync(e->d_res->flags, 0, 4 * sizeof(int), e->stream));
  if (!pi->has_elem) {  // solver.rs:1031-1042
    LAUNCH(e, k_pivot_rows, cdiv(m, 256), 256, 0, e->alpha, e->tau, e->xB, e->w, m, -1, pi->entering_new_val, pi->entering_diff,
           1.0, 0, 0, e->scal, (double*)nullptr, e->d_res->flags);
    LAUNCH(e, k_flip_var, 1, 1, 0, e->xnb, e->vflag, e->lo, e->hi, q, pi->entering_new_val);
    out->eta_count = e->K;
    return MLP_OK;
  }
  const int row = (int)pi->row;
  bool do_refactor = pi->refactor != 0;
  if (!do_refactor && e->K >= e->Kcap) do_refactor = true;  // arena full
  if (e->enable_dse) {
    // tau = B^-1 rho (solver.rs:1157). rho is by constraint row = the layout FTRAN takes.
    ST(ftran(e, e->rho, e->tau));
  }
  double* eta_col = do_refactor ? nullptr : e->E + (size_t)e->K * e->m;
  LAUNCH(e, k_pivot_rows, cdiv(m, 256), 256, 0, e->alpha, e->tau, e->xB, e->w, m, row, pi->entering_new_val, pi->entering_diff,
         pi->coeff, 1, e->enable_dse, e->scal, eta_col, e->d_res->flags);
  if (e->enable_pse) {
    // v = B^-T alpha_q (1114), helper = N^T v (1117-1132)
    CU(cudaMemcpyAsync(e->work_m, e->alpha, (size_t)m * 8, cudaMemcpyDeviceToDevice, e->stream));
    ST(btran(e, e->work_m, -1, e->vvec));
    compact(e, e->vvec, e->list_idx, e->list_val, e->icnt + 2, e->scal + 3);
    ST(price_list(e, e->list_idx, e->list_val, e->icnt + 2, 0, e->vvec, e->helper, 1));
  }
  LAUNCH(e, k_pivot_vars, cdiv(e->nt, 256), 256, 0, e->d, e->gam, e->rc, e->helper, e->vflag, e->nt, q, pi->coeff, e->enable_pse,
         e->scal, e->d_res->flags);
  LAUNCH(e, k_pivot_swap, 1, 1, 0, e->d, e->gam, e->xnb, e->vflag, e->vpos, e->bvar, e->loB, e->hiB, e->lo, e->hi, q, (int)pi->col,
         row, pi->coeff, pi->leaving_new_val, e->enable_pse, e->scal, e->d_res->flags, e->d_res);
  const int32_t lv_host = e->h_bvar[row];
  e->h_bvar[row] = (int32_t)q;
  if (!do_refactor) {
    const int prev = e->h_last_eta_of_row[row];
    LAUNCH(e, k_eta_grow, cdiv(std::max<int64_t>(e->K, 1), 256), 256, 0, e->E, e->m, (int)e->K, row, e->G, e->Kcap, e->etaR,
           e->etaPrev, e->etaHead, prev);
    e->h_last_eta_of_row[row] = (int)e->K;
    e->K += 1;
    e->cnt.etas_pushed += 1;
  }
  // one device->host read per pivot: status flags, leaving var, nnz(alpha)
  CU(cudaMemcpyAsync(&e->d_res->i[1], e->icnt + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
  if (e->prof_on) {
    CU(cudaMemcpyAsync(&e->d_res->i[2], e->icnt + 0, sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(&e->d_res->i[3], e->icnt + 2, sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
  }
  ST(fetch_res(e));
  if (e->prof_on) ST(collect_profile(e, e->h_res->i[2] & 0xffffffffLL, e->h_res->i[3] & 0xffffffffLL));
  out->leaving_var = e->h_res->i[0];
  out->col_nnz = (int64_t)(int32_t)(e->h_res->i[1] & 0xffffffffLL);
  if (out->leaving_var != lv_host) { set_err("pivot: host/device basis mirrors diverged"); return MLP_INVALID; }
  if (e->h_res->flags[0]) { set_err("non-finite steepest-edge norm"); return MLP_NONFINITE; }
  if (do_refactor) {
    ST(refactor_impl(e));
    out->refactored = 1;
    out->lu_nnz = e->lu_nnz;
  }
  out->eta_count = e->K;
  return MLP_OK;
}

mlp_status mlp_recalc_obj_coeffs(mlp_engine* e, double* cur_obj_val) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int m = (int)e->m;
  if (e->K > 0) ST(refactor_impl(e));  // solver.rs:1200-1203
  LAUNCH(e, k_gather_cB, cdiv(m, 256), 256, 0, e->cobj, e->bvar, m, e->work_m);
  ST(btran(e, e->work_m, -1, e->vvec));  // multipliers y (1205-1214)
  compact(e, e->vvec, e->list_idx, e->list_val, e->icnt + 2, e->scal + 3);
  ST(price_list(e, e->list_idx, e->list_val, e->icnt + 2, 0, e->vvec, e->helper));
  LAUNCH(e, k_recalc_d, cdiv(e->nt, 256), 256, 0, e->cobj, e->helper, e->vflag, e->nt, e->d);
  LAUNCH(e, k_recalc_obj, 1, 1024, 0, e->cobj, e->bvar, e->xB, m, e->xnb, e->vflag, e->nt, e->d_res);
  ST(fetch_res(e));
  *cur_obj_val = e->h_res->f[0];
  return MLP_OK;
}

mlp_status mlp_download_f64(mlp_engine* e, int32_t which, double* out, int64_t count) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const double* src = nullptr;
  int64_t len = 0;
  switch (which) {
    case MLP_ARR_OBJ_COEFFS: src = e->d; len = e->nt; break;
    case MLP_ARR_PRIMAL_NORMS: src = e->gam; len = e->nt; break;
    case MLP_ARR_NB_VALS: src = e->xnb; len = e->nt; break;
    case MLP_ARR_BASIC_VALS: src = e->xB; len = e->m; break;
    case MLP_ARR_DUAL_NORMS: src = e->w; len = e->m; break;
    case MLP_ARR_COL_COEFFS: src = e->alpha; len = e->m; break;
    case MLP_ARR_INV_BASIS_ROW: src = e->rho; len = e->m; break;
    case MLP_ARR_ROW_COEFFS: src = e->rc; len = e->nt; break;
    case MLP_ARR_BASIC_MINS: src = e->loB; len = e->m; break;
    case MLP_ARR_BASIC_MAXS: src = e->hiB; len = e->m; break;
    case MLP_ARR_SE_HELPER: src = e->helper; len = e->nt; break;
    default: return MLP_INVALID;
  }
  if (count != len) { set_err("download: count mismatch"); return MLP_INVALID; }
  return d2h(e, out, src, len * 8);
}
mlp_status mlp_download_basic_vars(mlp_engine* e, int64_t* out) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  std::vector<int32_t> tmp(e->m);
  ST(d2h(e, tmp.data(), e->bvar, e->m * 4));
  for (int64_t i = 0; i < e->m; ++i) out[i] = tmp[i];
  return MLP_OK;
}
mlp_status mlp_download_var_state(mlp_engine* e, uint8_t* flags, int32_t* pos) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  ST(d2h(e, flags, e->vflag, e->nt));
  return d2h(e, pos, e->vpos, e->nt * 4);
}
mlp_status mlp_get_counters(mlp_engine* e, mlp_counters* out) {
  if (!e) return MLP_INVALID;
  *out = e->cnt;
  out->lu_nnz = e->lu_nnz;
  out->eta_count = e->K;
  return MLP_OK;
}
void* mlp_engine_stream(mlp_engine* e) { return e ? (void*)e->stream : nullptr; }
mlp_status mlp_engine_sync(mlp_engine* e) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaGetLastError());
  return MLP_OK;
}

mlp_status mlp_event_mark(mlp_engine* e, int32_t slot) {
  if (!e || slot < 0 || slot > 3) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev[slot], e->stream));
  return MLP_OK;
}
mlp_status mlp_event_elapsed_ms(mlp_engine* e, int32_t a, int32_t b, double* ms) {
  if (!e || a < 0 || a > 3 || b < 0 || b > 3) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  CU(cudaEventSynchronize(e->ev[b]));
  float f = 0.f;
  CU(cudaEventElapsedTime(&f, e->ev[a], e->ev[b]));
  *ms = f;
  return MLP_OK;
}
mlp_status mlp_profile_enable(mlp_engine* e, int32_t on) {
  if (!e) return MLP_INVALID;
  e->prof_on = on;
  e->ppending[0] = e->ppending[1] = false;
  if (on) e->prof = mlp_profile{};
  return MLP_OK;
}
mlp_status mlp_profile_get(mlp_engine* e, mlp_profile* out) {
  if (!e || !out) return MLP_INVALID;
  *out = e->prof;
  return MLP_OK;
}

mlp_status mlp_bench_price_dense(mlp_engine* e, int32_t iters, double* ms_per_launch, int64_t* bytes_per_launch) {
  if (!e || iters <= 0) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int m = (int)e->m;
  LAUNCH(e, k_fill, cdiv(m, 256), 256, 0, e->vvec, (int64_t)m, 0.5);
  compact(e, e->vvec, e->list_idx, e->list_val, e->icnt + 2, e->scal + 3);
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  ST(price_list(e, e->list_idx, e->list_val, e->icnt + 2, 0, e->vvec, e->helper));  // warm-up
  CU(cudaEventRecord(a, e->stream));
  for (int i = 0; i < iters; ++i) ST(price_list(e, e->list_idx, e->list_val, e->icnt + 2, 0, e->vvec, e->helper));
  CU(cudaEventRecord(b, e->stream));
  CU(cudaEventSynchronize(b));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *ms_per_launch = (double)ms / iters;
  // algorithmic bytes: 8 n s + 8 s + 8 n  with s = m
  *bytes_per_launch = 8 * e->n * (int64_t)m + 8 * (int64_t)m + 8 * e->n;
  return MLP_OK;
}

}  // extern "C"
