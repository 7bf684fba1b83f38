// Microbenchmark: tcgen05.ld throughput while a 128x256x16 tcgen05.mma stream accumulates into the other half of TMEM.
#include <cstdio>
#include "../../fast3r_b200/csrc/common.cuh"
using namespace f3r;
__global__ void __launch_bounds__(384, 1) k(int iters, int mode, long long* cycles, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tptr; __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(&tptr);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  long long t0 = clock64();
  float acc = 0.f;
  if (warp == 1 && (mode & 1)) {
    if (lane == 0) {
      constexpr uint32_t id = make_idesc_bf16(128, 256, 0, 0);
      for (int it = 0; it < iters; ++it) {
        const uint64_t ad = make_smem_desc_sw128(smem_u32(smem + (it & 1) * 49152), 1);
        const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + (it & 1) * 49152 + 16384), 1);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tb, ad + 2 * kk, bd + 2 * kk, id, 1);
      }
      umma_commit(&bar); mbar_wait(&bar, 0);
    }
  } else if (warp >= 4 && (mode & 2)) {
    const uint32_t lb = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int nld = iters * 4 / 8;  // comparable duration
    for (int it = 0; it < nld; ++it) {
      uint32_t r[32];
      tmem_ld32(tb + lb + 256 + ((it * 32 + (warp >> 3) * 128) & 255), r);
      tmem_ld_wait();
      acc += __uint_as_float(r[0] & 1u);
    }
  }
  long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc<512>(tb); }
}
int main() {
  long long* cyc; float* sink; cudaMalloc(&cyc, 148 * 16 * 8); cudaMalloc(&sink, 148 * 384 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int mode : {1, 2, 3}) {
    cudaMemset(cyc, 0, 148 * 16 * 8);
    k<<<148, 384, 100 * 1024>>>(iters, mode, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long ld = 0; for (int i = 4; i < 12; ++i) ld = h[i] > ld ? h[i] : ld;
    printf("mode %d: MMA %.1f clk per 128x256x16 MMA | LDTM (8 warps) %.1f clk per 4KB load per warp = %.0f B/clk/SM (%s)\n", mode,
           double(h[1]) / (iters * 4.0), double(ld) / (iters * 4 / 8), 8.0 * 4096 * (iters * 4 / 8) / double(ld), cudaGetErrorString(e));
  }
  return 0;
}
