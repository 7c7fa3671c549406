// Micro-benchmark: cost of one grid-wide barrier of a cooperative one-CTA-per-SM grid (the barrier of chain_fused.cuh)
// in three forms.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cuda/atomic>

__device__ __forceinline__ void bar_fence(unsigned* bar) {  // A: __threadfence + atomicAdd + volatile spin (cooperative-groups scheme)
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned nb = 1;
    if (blockIdx.x == 0) nb = 0x80000000u - (gridDim.x - 1);
    __threadfence();
    const unsigned old = atomicAdd(bar, nb);
    while (((old ^ *((volatile unsigned*)bar)) & 0x80000000u) == 0) {}
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ void bar_acqrel(unsigned* bar) {  // B: release add + acquire load
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned nb = 1;
    if (blockIdx.x == 0) nb = 0x80000000u - (gridDim.x - 1);
    cuda::atomic_ref<unsigned, cuda::thread_scope_device> a(*bar);
    const unsigned old = a.fetch_add(nb, cuda::memory_order_release);
    while (((old ^ a.load(cuda::memory_order_acquire)) & 0x80000000u) == 0) {}
  }
  __syncthreads();
}
__device__ __forceinline__ void bar_flags(unsigned* flags, unsigned gen) {  // C: per-CTA flag slots, every CTA polls all slots with one warp
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    *((volatile unsigned*)(flags + blockIdx.x * 32)) = gen;  // 128-byte slots
  }
  if (threadIdx.x < 32) {
    for (unsigned b = threadIdx.x; b < gridDim.x; b += 32)
      while (*((volatile unsigned*)(flags + b * 32)) < gen) {}
    __threadfence();
  }
  __syncthreads();
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_bars(unsigned* bar, unsigned* flags, int iters, unsigned gen0, double* sink) {
  double acc = 0.0;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) bar_fence(bar);
    else if (MODE == 1) bar_acqrel(bar);
    else bar_flags(flags, gen0 + i + 1);
    acc += i;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}
int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned *bar, *flags;
  double* sink;
  cudaMalloc(&bar, 256); cudaMemset(bar, 0, 256);
  cudaMalloc(&flags, sms * 128); cudaMemset(flags, 0, sms * 128);
  cudaMalloc(&sink, 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 1000;
  unsigned gen = 0;
  for (int mode = 0; mode < 3; ++mode)
    for (int rep = 0; rep < 3; ++rep) {
      int it = iters;
      unsigned g0 = gen;
      void* args[] = {&bar, &flags, &it, &g0, &sink};
      const void* fn = mode == 0 ? (const void*)k_bars<0> : mode == 1 ? (const void*)k_bars<1> : (const void*)k_bars<2>;
      cudaEventRecord(a);
      cudaError_t err = cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(512), args, 0, 0);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      if (mode == 2) gen += iters;
      printf("{\"mode\": \"%s\", \"sms\": %d, \"us_per_barrier\": %.3f, \"err\": \"%s\"}\n",
             mode == 0 ? "threadfence+atomicAdd+volatile" : mode == 1 ? "release-add+acquire-load" : "flag-slots", sms,
             1000.0 * ms / iters, cudaGetErrorString(err));
    }
  return 0;
}
