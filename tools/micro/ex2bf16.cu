#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
__global__ void rate(int iters, long long* cycles, uint32_t* sink) {
  uint32_t v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0xbf80bf00u + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0) cycles[blockIdx.x * 32 + threadIdx.x / 32] = t1 - t0;
  uint32_t s = 0; for (int i = 0; i < 8; ++i) s ^= v[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void acc(float* out) {  // out[i] = rel err stats over x in [-16, 0]
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float x = -16.0f * i / (gridDim.x * blockDim.x);
  uint32_t w; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(x), "f"(x));
  uint32_t r; asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(w));
  float got = __uint_as_float(r << 16);
  out[i] = got / exp2f(x) - 1.0f;
}
int main() {
  long long* cyc; uint32_t* sink; cudaMalloc(&cyc, 148 * 32 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  for (int warps : {4, 8, 16}) {
    rate<<<148, warps * 32>>>(2000, cyc, sink); cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("ex2.bf16x2 warps/SMSP %d: %.2f clk per warp-instr per SMSP (%s)\n", warps / 4, double(mx) / (2000.0 * 64) / (warps / 4), cudaGetErrorString(cudaGetLastError()));
  }
  float* out; cudaMalloc(&out, 65536 * 4); acc<<<256, 256>>>(out); cudaDeviceSynchronize();
  static float h[65536]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  for (int seg = 0; seg < 8; ++seg) { double m = 0, s2 = 0; for (int i = seg * 8192; i < (seg + 1) * 8192; ++i) { m = fmax(m, fabs(h[i])); s2 += double(h[i]) * h[i]; }
    printf("x in [%d,%d]: max rel err %.2e rms %.2e\n", -2 * (seg + 1), -2 * seg, m, sqrt(s2 / 8192)); }
  return 0;
}
