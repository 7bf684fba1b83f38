// Microbenchmark: cost of the attention kernel's MMA stream and softmax stream in isolation and together
// (no data dependencies between them), cycles per key-block iteration (2 tiles x 128 keys, hd 64).
#include <cstdio>
#include "../../fast3r_b200/csrc/common.cuh"
using namespace f3r;

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%4}; mov.b64 rc, {%5,%5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{ .reg .b64 ra, rd; mov.b64 rd, {%0,%1}; mov.b64 ra, {%2,%3};\n\t"
      "add.rn.f32x2 rd, rd, ra; mov.b64 {%0,%1}, rd; }" : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}

__device__ __forceinline__ void fsub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {  // a - b
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%6};\n\t"
      "fma.rn.f32x2 rd, rb, rc, ra; mov.b64 {%0,%1}, rd; }" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(-1.0f));
}
__device__ __forceinline__ void fma2v(float& d0, float& d1, float a0, float a1, float b0, float b1, float c) {  // a*b + c
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c));
}
// exp2 of a pair on the FMA/ALU pipes: round-to-nearest split + degree-3 minimax polynomial (rel err 7.6e-5)
__device__ __forceinline__ void exp2_emu2(float& e0, float& e1, float x0, float x1) {
  x0 = fmaxf(x0, -126.f); x1 = fmaxf(x1, -126.f);
  float t0 = x0, t1 = x1;
  fadd2(t0, t1, 12582912.f, 12582912.f);        // t = x + 1.5*2^23 : low mantissa bits = round(x)
  float r0 = t0, r1 = t1;
  fadd2(r0, r1, -12582912.f, -12582912.f);      // r = round(x)
  float f0, f1;
  fsub2(f0, f1, x0, x1, r0, r1);                // f = x - r in [-0.5, 0.5]
  float p0, p1;
  fma2v(p0, p1, f0, f1, 0.05520550534f, 0.05520550534f, 0.2426139712f);
  fma2v(p0, p1, p0, p1, f0, f1, 0.6932547688f);
  fma2v(p0, p1, p0, p1, f0, f1, 0.9999276996f);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

#ifndef EMU_MASK
#define EMU_MASK 0x0  // bit k set: pair k of every 8 pairs uses the FMA-pipe exp2
#endif

// mode bit0: run MMA stream, bit1: run softmax stream; mma_kind: 0 both, 1 SS only, 2 TS only
__global__ void __launch_bounds__(384, 1) k(int iters, int mode, int mma_kind, long long* cycles, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(&tptr);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  long long t0 = clock64();
  float acc = 0.f;
  if (warp == 1 && (mode & 1)) {
    if (lane == 0) {
      constexpr uint32_t idqk = make_idesc_bf16(128, 128, 0, 0), idpv = make_idesc_bf16(128, 64, 0, 1);
      for (int it = 0; it < iters; ++it) {
        const int st = it & 3;
        const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + 32768 + st * 16384), 1);
        const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + 32768 + 65536 + st * 16384), 0);
        for (int t = 0; t < 2; ++t) {
          const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + t * 16384), 1);
          if (mma_kind != 2)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ss(tb + t * 128, qd + 2 * kk, kd + 2 * kk, idqk, kk > 0);
          if (mma_kind != 1)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_ts(tb + 384 + t * 64, tb + 256 + t * 64 + 8 * kk, vd + 128 * kk, idpv, 1);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
    }
  } else if (warp >= 4 && (mode & 2)) {
    const int t = (warp - 4) >> 2, quarter = warp & 3;
    const uint32_t lb = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tm_s = tb + lb + t * 128, tm_p = tb + lb + 256 + t * 64;
    float m_used = 0.f, l = 0.f;
    const float sl2 = 0.23f;
    for (int it = 0; it < iters; ++it) {
      uint32_t s[128];
      tmem_ld32(tm_s, s); tmem_ld32(tm_s + 32, s + 32); tmem_ld32(tm_s + 64, s + 64); tmem_ld32(tm_s + 96, s + 96);
      tmem_ld_wait();
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]), mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 8; i < 128; i += 8) {
        mx0 = max3(mx0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        mx1 = max3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = max3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = max3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if ((mx - m_used) * sl2 > 8.f) { l *= ex2a((m_used - mx) * sl2); m_used = mx; }
      const float nm = -m_used * sl2;
      float l0 = 0, l1 = 0, l2 = 0, l3 = 0;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        float x0, x1, x2, x3;
        ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sl2, nm);
        ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sl2, nm);
        float e0, e1, e2, e3;
        if ((EMU_MASK >> ((i / 2) & 7)) & 1) exp2_emu2(e0, e1, x0, x1); else { e0 = ex2a(x0); e1 = ex2a(x1); }
        if ((EMU_MASK >> ((i / 2 + 1) & 7)) & 1) exp2_emu2(e2, e3, x2, x3); else { e2 = ex2a(x2); e3 = ex2a(x3); }
        fadd2(l0, l1, e0, e1); fadd2(l2, l3, e2, e3);
        pk[i / 2] = pack_bf16(e0, e1); pk[i / 2 + 1] = pack_bf16(e2, e3);
      }
      l += (l0 + l1) + (l2 + l3);
      tmem_st32(tm_p, pk); tmem_st32(tm_p + 32, pk + 32);
      tmem_st_wait();
    }
    acc = l;
  }
  long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 16 * 8); cudaMalloc(&sink, 148 * 384 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int iters = 400;
  struct { int mode, kind; const char* name; } cases[] = {
    {2, 0, "softmax only (8 warps)"}, {3, 0, "MMA + softmax, independent"}};
  for (auto& c : cases) {
    cudaMemset(cyc, 0, 148 * 16 * 8);
    k<<<148, 384, 170 * 1024>>>(iters, c.mode, c.kind, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mma = h[1], sm = 0; for (int i = 4; i < 12; ++i) sm = h[i] > sm ? h[i] : sm;
    printf("%-32s: MMA warp %.0f clk/iter, softmax warps %.0f clk/iter (%s)\n", c.name, double(mma) / iters, double(sm) / iters,
           cudaGetErrorString(e));
  }
  return 0;
}
