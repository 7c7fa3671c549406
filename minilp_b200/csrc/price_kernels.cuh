// Dense price-out kernels: the LDG form and the bulk-copy (TMA) ring.  Included by engine.cu only, after the
// engine types and kernels_common.cuh.
#pragma once

// ------------------------------------------------------------------------------------------------ K5 price-out
// calc_row_coeffs price-out (solver.rs:685-692), the N^T v product of update_primal_sq_norms (1117-1132), the full
// c_N - N^T y of recalc_obj_coeffs (1216-1222) and the column norms of try_new (297-299).  Row-gather GEMV^T over
// row-major A: a CTA owns a 512-column tile and one chunk of the multiplier's support; each thread accumulates two
// adjacent columns with 128-bit streaming loads, 8 rows in flight; rows of a chunk are taken in list order;
// (row, weight) pairs are staged through shared memory.  The chunking depends ONLY on the support size s
// (C = clamp(ceil(s/128), 1, 64)), never on the grid or the shard width, so a column's sum is bit-identical however
// the columns are sharded.  Chunk partials are reduced in chunk order by k_price_finish — no atomics.
constexpr int PR_THREADS = 256;
constexpr int PR_TILE = PR_THREADS * 2;
constexpr int PR_BATCH = 256;
constexpr int PR_UNROLL = 8;
constexpr int PR_MAXC = 64;
__host__ __device__ __forceinline__ int price_chunks_for(int s) {
  int c = (s + 127) / 128;
  return c < 1 ? 1 : (c > PR_MAXC ? PR_MAXC : c);
}

template <int MODE>  // 0: sum_r w_r * A[r,j]   1: sum_r A[r,j]^2
__global__ void __launch_bounds__(PR_THREADS)
k_price_partial(const double* __restrict__ A, int64_t lda, const int32_t* __restrict__ rows,
                const double* __restrict__ wts, const int32_t* __restrict__ count_ptr, int32_t fixed_count,
                double* __restrict__ partial) {
  pdl_wait();
  __shared__ int32_t srow[PR_BATCH];
  __shared__ double sw[PR_BATCH];
  const int s = count_ptr ? *count_ptr : fixed_count;
  const int C = price_chunks_for(s);
  const int L = (s + C - 1) / C;
  const int tiles = (int)((lda + PR_TILE - 1) / PR_TILE);
  // Persistent CTAs: the grid is sized to a fixed number of CTAs per SM (not to the work), so that the rest of each SM stays
  // free for the latency-bound kernels of the other lane; work item = (column tile, support chunk).
  for (int item = blockIdx.x; item < tiles * C; item += gridDim.x) {
    const int tile = item % tiles, chunk = item / tiles;
    const int k0 = chunk * L;
    const int k1 = min(s, k0 + L);
    const int64_t col = ((int64_t)tile * PR_THREADS + threadIdx.x) * 2;
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
      *reinterpret_cast<double2*>(partial + (int64_t)chunk * lda + col) = o;
    }
  }
}

// ---- bulk-copy (TMA) form of the same price-out --------------------------------------------------------------
// The LDG kernel above needs 6 resident CTAs per SM (all registers) to keep enough bytes in flight; that starves the
// latency-bound kernels of lane 1 that should run beside it.  Here the bytes in flight live in SHARED memory instead:
// one CTA per SM, a producer warp gathers the listed rows with 1-D bulk copies (cp.async.bulk, 4 KB row segments,
// L2 evict-first) into a TP_STAGES-deep ring guarded by mbarriers, eight consumer warps accumulate.  Work items,
// chunking and the per-column accumulation order (list order within a chunk, thread t owns columns 2t, 2t+1 of the
// tile) are those of k_price_partial, so the partial sums are bit-identical.
constexpr int TP_STAGE_BYTES = 32768;             // one stage: R rows x tile_cols x 8 B, R = 32768 / (tile_cols * 8) <= 32
constexpr int TP_STAGES = 6;                      // ring depth: 6 x 32 KB = 192 KB in flight per SM
constexpr int TP_MAXROWS = 32;                    // one row per producer lane
constexpr int TP_CONSUMERS = PR_THREADS;          // 8 warps
constexpr int TP_THREADS = TP_CONSUMERS + 32;     // + producer warp
constexpr size_t TP_SMEM = (size_t)TP_STAGES * TP_STAGE_BYTES + TP_STAGES * TP_MAXROWS * 8 + 2 * TP_STAGES * 8 + 16;
// Tile width: 512 columns (4 KB row segments).  Narrower tiles were measured and lose: 6.77 TB/s at 512, 6.19 at 256,
// 4.30 at 128 columns (more, smaller bulk copies per byte); a narrow column block of a sharded engine has fewer work
// items per SM, but the tail still has enough SMs active to saturate HBM.  MLP_PRICE_TILE overrides for experiments.
// Choice of the tiling (host), from the measured sweeps in profiles/r01d_price_sweep.md (B200, isolated dense N^T v,
// m = 50k; GB/s at 512-column tiles -> at the width chosen here):  n_loc 50 000: 6760 -> 7068 (1280);  25 000: 6735 -> 7070
// (1280);  12 500: 6668 -> 7102 (4096);  6 250: 6125 -> 6998 (4096).  Wider row segments mean fewer, larger bulk copies and
// longer contiguous DRAM bursts, and that matters more the shorter the rows of the local block are; widths whose rows fill
// a 32 KB stage badly (2560 columns = 20 KB: one row per stage) lose the bytes in flight again.  The kernel is HBM-bound
// with ~28 MB in flight, so a partly filled last round of work items costs little (a round model that predicted gains
// from balancing it did not survive the measurement; the tail split stays available as a knob, default 1).
static void choose_price_tiling(int64_t lda, int64_t /*m*/, int /*G*/, int* tile, int* split) {
  *split = 1;
  const int w = lda < 20000 ? 4096 : 1280;
  const int need = (int)std::min<int64_t>(4096, (lda + 63) / 64 * 64);  // never wider than the block itself
  *tile = std::max(128, std::min(w, need));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(0x12F0000000000000ull)  // L2 evict-first: A is streamed once per pivot
      : "memory");
}

// NV = column pairs per consumer thread: a tile is up to NV * 512 columns wide (NV * 4 KB row segments).  Thread t owns
// column pairs t, t + 256, ... of the tile; every column still accumulates its rows in list order, so the partial sums
// do not depend on the tiling.
//
// Work items and load balance.  An item is (column tile, support chunk); CTA b takes items b, b + G, ... (G CTAs).  With
// tiles * C items the last round is only partly filled — at 50k columns and 1024-column tiles 3136 items are 21.2
// rounds, i.e. 4 % of the kernel runs with 80 % of the SMs idle.  So only the items of the FULL rounds keep the whole tile
// width; the items of the last, partial round are cut into `split` column slices each, which spreads that round over
// all CTAs again (narrower row segments are less efficient, but only the tail pays that).
struct PriceItem {
  int chunk;
  int64_t col0;
  int cols;   // columns actually present (0: the slice lies beyond the matrix)
  int width;  // nominal width: row stride of the stage in shared memory
};
__device__ __forceinline__ PriceItem price_item(int idx, int main_items, int tiles, int tile_cols, int split, int64_t lda) {
  PriceItem it;
  int item, sub = 0;
  it.width = tile_cols;
  if (idx < main_items) item = idx;
  else {
    const int j = idx - main_items;
    item = main_items + j / split;
    sub = j % split;
    it.width = tile_cols / split;
  }
  it.chunk = item / tiles;
  it.col0 = (int64_t)(item % tiles) * tile_cols + (int64_t)sub * it.width;
  const int64_t tile_end = min(lda, (int64_t)(item % tiles + 1) * tile_cols);
  const int64_t c = min((int64_t)it.width, tile_end - it.col0);
  it.cols = c > 0 ? (int)c : 0;
  return it;
}
template <int NV>
__global__ void __launch_bounds__(TP_THREADS, 1)
k_price_partial_tma(const double* __restrict__ A, int64_t lda, const int32_t* __restrict__ rows,
                    const double* __restrict__ wts, const int32_t* __restrict__ count_ptr, int32_t fixed_count,
                    double* __restrict__ partial, int tile_cols, int split) {
  pdl_wait();
  extern __shared__ __align__(128) unsigned char tp_smem[];
  double* sw = reinterpret_cast<double*>(tp_smem + (size_t)TP_STAGES * TP_STAGE_BYTES);   // [stage][row] weights
  uint64_t* full = reinterpret_cast<uint64_t*>(sw + TP_STAGES * TP_MAXROWS);
  uint64_t* empty = full + TP_STAGES;
  const int s = count_ptr ? *count_ptr : fixed_count;
  const int C = price_chunks_for(s);
  const int L = (s + C - 1) / C;
  const int tiles = (int)((lda + tile_cols - 1) / tile_cols);
  const int items = tiles * C;
  const int main_items = items / (int)gridDim.x * (int)gridDim.x;
  const int total = main_items + (items - main_items) * split;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int q = 0; q < TP_STAGES; ++q) { mbar_init(full + q, 1); mbar_init(empty + q, TP_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  uint32_t it = 0;  // stages handled so far by this role: slot = it % TP_STAGES, phase = (it / TP_STAGES) & 1
  if (warp == TP_CONSUMERS / 32) {
    // ---------------- producer warp: lane r < R fetches row r of the stage
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
      const PriceItem pi = price_item(idx, main_items, tiles, tile_cols, split, lda);
      if (pi.cols == 0) continue;
      const int k0 = pi.chunk * L, k1 = min(s, k0 + L);
      const int tile_bytes = pi.width * 8;
      const int R = min(TP_MAXROWS, TP_STAGE_BYTES / tile_bytes);  // rows per stage
      const uint32_t tbytes = (uint32_t)pi.cols * 8u;
      for (int kb = k0; kb < k1; kb += R, ++it) {
        const int nr = min(R, k1 - kb);
        const int slot = it % TP_STAGES;
        int32_t r = 0;
        double wv = 0.0;
        if (lane < nr) { r = rows[kb + lane]; wv = wts[kb + lane]; }  // issued before the wait: latency overlaps
        mbar_wait(empty + slot, ((it / TP_STAGES) & 1) ^ 1);
        if (lane < nr) sw[slot * TP_MAXROWS + lane] = wv;
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(full + slot, tbytes * nr);
        __syncwarp();
        if (lane < nr)
          bulk_g2s(tp_smem + (size_t)slot * TP_STAGE_BYTES + (size_t)lane * tile_bytes, A + (int64_t)r * lda + pi.col0, tbytes,
                   full + slot);
      }
    }
  } else {
    // ---------------- consumers: thread t owns column pairs t + 256 v (v < NV) of the tile
    const int t = threadIdx.x;
    constexpr int RU = 8 / NV;            // rows per unrolled batch: 8 loads in flight per thread
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
      const PriceItem pi = price_item(idx, main_items, tiles, tile_cols, split, lda);
      if (pi.cols == 0) continue;
      const int k0 = pi.chunk * L, k1 = min(s, k0 + L);
      const int tile_bytes = pi.width * 8;
      const int R = min(TP_MAXROWS, TP_STAGE_BYTES / tile_bytes);
      const int rstride = tile_bytes / 16;  // double2 per row
      const int64_t col = pi.col0 + 2 * t;
      bool active[NV];
      double acc0[NV], acc1[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        active[v] = 2 * (t + TP_CONSUMERS * v) < pi.cols;
        acc0[v] = 0.0;
        acc1[v] = 0.0;
      }
      for (int kb = k0; kb < k1; kb += R, ++it) {
        const int nr = min(R, k1 - kb);
        const int slot = it % TP_STAGES;
        mbar_wait(full + slot, (it / TP_STAGES) & 1);
        if (active[0]) {
          const double2* p = reinterpret_cast<const double2*>(tp_smem + (size_t)slot * TP_STAGE_BYTES) + t;
          const double* w = sw + slot * TP_MAXROWS;
          int u0 = 0;
          for (; u0 + RU <= nr; u0 += RU) {
            double2 x[RU][NV];
#pragma unroll
            for (int u = 0; u < RU; ++u)
#pragma unroll
              for (int v = 0; v < NV; ++v)
                if (NV == 1 || active[v]) x[u][v] = p[(u0 + u) * rstride + TP_CONSUMERS * v];
#pragma unroll
            for (int u = 0; u < RU; ++u) {
              const double wv = w[u0 + u];
#pragma unroll
              for (int v = 0; v < NV; ++v)
                if (NV == 1 || active[v]) {
                  acc0[v] += wv * x[u][v].x;
                  acc1[v] += wv * x[u][v].y;
                }
            }
          }
          for (; u0 < nr; ++u0) {
            const double wv = w[u0];
#pragma unroll
            for (int v = 0; v < NV; ++v)
              if (NV == 1 || active[v]) {
                const double2 x = p[u0 * rstride + TP_CONSUMERS * v];
                acc0[v] += wv * x.x;
                acc1[v] += wv * x.y;
              }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (active[v]) {
          double2 o;
          o.x = acc0[v];
          o.y = acc1[v];
          *reinterpret_cast<double2*>(partial + (int64_t)pi.chunk * lda + col + 2 * TP_CONSUMERS * v) = o;
        }
    }
  }
}

// Reduce chunk partials in chunk order; slack columns of [A|I] contribute rho_i (the `I` part of the CSR row,
// solver.rs:250); basic variables are not part of row_coeffs (solver.rs:688).
// mode 0: out = sum   mode 1: out = sum + 1 (primal edge norms, solver.rs:298)
__global__ void k_price_finish(const double* __restrict__ partial, const int32_t* __restrict__ count_ptr, int32_t fixed_count,
                               int64_t lda, int64_t n, int64_t m, const double* __restrict__ slack_vals,
                               const uint8_t* __restrict__ vflag, double* __restrict__ out, int mode) {
  pdl_wait();
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n + m) return;
  const int C = price_chunks_for(count_ptr ? *count_ptr : fixed_count);
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

