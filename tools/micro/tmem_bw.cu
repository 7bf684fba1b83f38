// Microbenchmark: tcgen05.ld (32x32b.x32) throughput per SM as a function of the number of reading warps.
#include <cstdio>
#include "../../fast3r_b200/csrc/common.cuh"
using namespace f3r;

__global__ void k(int iters, long long* cycles, float* sink, int nwarps) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tptr;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(base + lane_base + ((c * 32 + (warp >> 2) * 128) & 511), r);
        tmem_ld_wait();
        acc += __uint_as_float(r[0] & 1u);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0 && warp < nwarps) cycles[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(base); }
}

// variant: issue all 4 loads, then one wait (what the attention kernel does)
__global__ void k2(int iters, long long* cycles, float* sink, int nwarps) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tptr;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(base + lane_base + ((c * 32 + (warp >> 2) * 128) & 511), r + 32 * c);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 128; c += 8) acc += __uint_as_float(r[c] & 1u);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0 && warp < nwarps) cycles[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(base); }
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 32 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2000;
  for (int variant = 0; variant < 2; ++variant)
    for (int nw : {1, 2, 4, 8, 16}) {
      cudaMemset(cyc, 0, 148 * 32 * 8);
      if (variant == 0) k<<<148, 512>>>(iters, cyc, sink, nw); else k2<<<148, 512>>>(iters, cyc, sink, nw);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[32]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
      double bytes = double(nw) * iters * 4 * 4096;
      printf("variant %d warps %2d: %lld cycles, %.1f B/clk/SM, %.1f clk per 4KB load per warp (%s)\n", variant, nw, mx,
             bytes / mx, double(mx) / (iters * 4), cudaGetErrorString(e));
    }
  return 0;
}
