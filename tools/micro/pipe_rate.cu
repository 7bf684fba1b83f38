// Microbenchmark: issue rate (cycles per warp-instruction per SM sub-partition) of the softmax instruction mix.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#define REP 64
template <int OP>
__global__ void k(int iters, long long* cycles, float* sink, float a, float b) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = a + i * 0.001f + threadIdx.x * 1e-6f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < REP / 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        if (OP == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i + 1])); }
        if (OP == 1) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i + 1]) : "f"(a), "f"(b)); }
        if (OP == 2) { asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%2}; mov.b64 rc, {%3,%3}; fma.rn.f32x2 ra, ra, rb, rc; mov.b64 {%0,%1}, ra; }" : "+f"(v[i]), "+f"(v[i + 1]) : "f"(a), "f"(b));
                       asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%2}; mov.b64 rc, {%3,%3}; fma.rn.f32x2 ra, ra, rb, rc; mov.b64 {%0,%1}, ra; }" : "+f"(v[i]), "+f"(v[i + 1]) : "f"(a), "f"(b)); }
        if (OP == 3) { asm volatile("{ .reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%2}; add.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra; }" : "+f"(v[i]), "+f"(v[i + 1]) : "f"(a));
                       asm volatile("{ .reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%2}; add.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra; }" : "+f"(v[i]), "+f"(v[i + 1]) : "f"(a)); }
        if (OP == 4) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i + 1]) : "f"(a), "f"(b)); }
        if (OP == 5) { uint32_t p; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v[i]), "f"(v[i + 1])); v[i] = __uint_as_float(p);
                       asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v[i + 1]), "f"(v[i])); v[i + 1] = __uint_as_float(p); }
        if (OP == 6) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(a)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i + 1]) : "f"(a)); }
        if (OP == 7) { asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(a)); asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i + 1]) : "f"(a)); }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0) cycles[blockIdx.x * 32 + threadIdx.x / 32] = t1 - t0;
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, long long* cyc, float* sink) {
  for (int warps : {4, 8, 16}) {
    k<OP><<<148, warps * 32>>>(2000, cyc, sink, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-10s warps/SMSP %d: %.2f clk per warp-instr per SMSP\n", name, warps / 4, double(mx) / (2000.0 * REP) / (warps / 4));
  }
}
int main() {
  long long* cyc; float* sink; cudaMalloc(&cyc, 148 * 32 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  run<0>("MUFU.EX2", cyc, sink); run<1>("FFMA", cyc, sink); run<2>("FFMA2", cyc, sink); run<3>("FADD2", cyc, sink);
  run<4>("FMNMX3", cyc, sink); run<5>("F2FP", cyc, sink); run<6>("FADD", cyc, sink); run<7>("FMNMX", cyc, sink);
  return 0;
}
