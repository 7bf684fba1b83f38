// Parity-mode attention (head_dim 64): softmax(scale * Q K^T) V with fp32-level accuracy on the bf16 tensor pipe.
// (fast3r/croco/models/blocks.py:135-194 run WITHOUT autocast, i.e. the reference's fp32 path,
//  fast3r/dust3r/inference_multiview.py:41-49 dtype="32".)
//
// Every fp32 operand x is carried as a pair of bf16 numbers x = hi + lo (hi = bf16(x), lo = bf16(x - hi), 16 mantissa
// bits together) and every product is evaluated as hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator (the lo*lo term,
// 2^-18 relative, is dropped):
//   S = Q K^T : the head dimension is "concatenated" to 192: Q' = [Qhi | Qlo | Qhi], K' = [Khi | Khi | Klo]
//               (written by attn_split_kernel below) -> 12 k-steps of one 128x128x16 SS MMA chain instead of 4;
//   O += P V  : P is split in the softmax registers into Phi / Plo (two packed-bf16 TMEM operands), V arrives as
//               [Vhi | Vlo]; three TS MMA chains Phi*Vhi + Plo*Vhi + Phi*Vlo accumulate into the same O columns.
// Softmax statistics, the row sum (of the un-split fp32 p) and the output are fp32.  One CTA = 128 query rows of one
// (batch, head); K'/V' blocks of 128 keys stream through a 2-stage TMA ring.  This kernel trades speed for accuracy
// (3x the MMAs, no query-tile ping-pong); the bf16 kernel in attention.cu is the fast path.
#include "common.cuh"
#include "f3r_kernels.h"

namespace f3r {

constexpr int X3_THREADS = 256;  // warp 0 TMA, warp 1 MMA (+ TMEM alloc), warps 2-3 idle, warps 4-7 softmax
constexpr int X3_STAGES = 2;
constexpr int X3_TILE = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, 128B-swizzled
constexpr int X3_SMEM_BYTES = (3 + X3_STAGES * 5) * X3_TILE + 1024 + 256;
constexpr uint32_t X3_TM_S = 0, X3_TM_PHI = 128, X3_TM_PLO = 192, X3_TM_O = 256;

__global__ void __launch_bounds__(X3_THREADS, 1)
attention_x3_kernel(const __grid_constant__ CUtensorMap tmap_q3, const __grid_constant__ CUtensorMap tmap_k3,
                    const __grid_constant__ CUtensorMap tmap_v2, const __grid_constant__ AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                  // 3 sub-tiles
  uint8_t* smem_k = smem + 3 * X3_TILE;                    // X3_STAGES x 3 sub-tiles
  uint8_t* smem_v = smem_k + X3_STAGES * 3 * X3_TILE;      // X3_STAGES x 2 sub-tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + X3_STAGES * 2 * X3_TILE);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + X3_STAGES;
  uint64_t* v_full = k_empty + X3_STAGES;
  uint64_t* v_empty = v_full + X3_STAGES;
  uint64_t* s_full = v_empty + X3_STAGES;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int nkv = (p.skv + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q3);
    tma_prefetch_desc(&tmap_k3);
    tma_prefetch_desc(&tmap_v2);
    mbar_init(q_full, 1);
    for (int s = 0; s < X3_STAGES; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_full, 128); mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, 3 * X3_TILE);
      for (int s = 0; s < 3; ++s) tma_load_3d(smem_q + s * X3_TILE, &tmap_q3, q_full, h * 192 + s * 64, qt * 128, b);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&k_full[stage], 3 * X3_TILE);
        for (int s = 0; s < 3; ++s)
          tma_load_3d(smem_k + (stage * 3 + s) * X3_TILE, &tmap_k3, &k_full[stage], h * 192 + s * 64, j * 128, b);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&v_full[stage], 2 * X3_TILE);
        for (int s = 0; s < 2; ++s)
          tma_load_3d(smem_v + (stage * 2 + s) * X3_TILE, &tmap_v2, &v_full[stage], h * 128 + s * 64, j * 128, b);
        if (++stage == X3_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      auto issue_s = [&](int stage) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const uint64_t qd = make_smem_desc_sw128(smem_u32(smem_q + s * X3_TILE), 1);
          const uint64_t kd = make_smem_desc_sw128(smem_u32(smem_k + (stage * 3 + s) * X3_TILE), 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + X3_TM_S, qd + 2 * k, kd + 2 * k, idesc_qk, (s | k) ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(&k_empty[stage]);
      };
      auto issue_pv = [&](int stage, int j) {
        const uint64_t vhi = make_smem_desc_sw128(smem_u32(smem_v + (stage * 2 + 0) * X3_TILE), 0);
        const uint64_t vlo = make_smem_desc_sw128(smem_u32(smem_v + (stage * 2 + 1) * X3_TILE), 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tmem_base + X3_TM_O, tmem_base + X3_TM_PHI + 8 * k, vhi + 128 * k, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ts(tmem_base + X3_TM_O, tmem_base + X3_TM_PLO + 8 * k, vhi + 128 * k, idesc_pv, 1u);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ts(tmem_base + X3_TM_O, tmem_base + X3_TM_PHI + 8 * k, vlo + 128 * k, idesc_pv, 1u);
        umma_commit(pv_done);
        umma_commit(&v_empty[stage]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % X3_STAGES;
        const uint32_t ph = (j / X3_STAGES) & 1;
        if (j + 1 < nkv) {
          const int st1 = (j + 1) % X3_STAGES;
          const uint32_t ph1 = ((j + 1) / X3_STAGES) & 1;
          mbar_wait(&k_full[st1], ph1);
          mbar_wait(s_free, j & 1);  // the softmax warps hold S_j in registers
          tc_fence_after();
          issue_s(st1);
        }
        mbar_wait(&v_full[st], ph);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        issue_pv(st, j);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax warps: one thread per query row =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tm_s = tmem_base + lane_base + X3_TM_S;
    const uint32_t tm_phi = tmem_base + lane_base + X3_TM_PHI;
    const uint32_t tm_plo = tmem_base + lane_base + X3_TM_PLO;
    const uint32_t tm_o = tmem_base + lane_base + X3_TM_O;
    const float sl2 = p.scale_log2;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tm_s + 32 * c, s + 32 * c);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);
      if (j == nkv - 1) {
        const int valid = p.skv - j * 128;
        if (valid < 128) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) s[i] = 0xff800000u;  // -inf
        }
      }
      float mx = __uint_as_float(s[0]);
#pragma unroll
      for (int i = 1; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
      float alpha = 1.f;
      const bool need = (mx - m_used) * sl2 > 8.f;  // lazy reference move (first block: -inf reference => true)
      if (need) {
        alpha = exp2f((m_used - mx) * sl2);
        m_used = mx;
        l *= alpha;
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // O quiescent, P_{j-1} consumed
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          uint32_t o[32];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tmem_ld32(tm_o + 32 * c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tm_o + 32 * c, o);
          }
          tmem_st_wait();
        }
      }
      const float nm = -m_used * sl2;
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float e0 = exp2f(fmaf(__uint_as_float(s[32 * c + i]), sl2, nm));
          const float e1 = exp2f(fmaf(__uint_as_float(s[32 * c + i + 1]), sl2, nm));
          lsum += e0 + e1;
          const uint32_t h2 = pack_bf16(e0, e1);
          hi[i / 2] = h2;
          lo[i / 2] = pack_bf16(e0 - bf16_lo(h2), e1 - bf16_hi(h2));
        }
        tmem_st16(tm_phi + 16 * c, hi);
        tmem_st16(tm_plo + 16 * c, lo);
      }
      l += lsum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> fp32 global
    mbar_wait(pv_done, (nkv - 1) & 1);
    tc_fence_after();
    const int q = qt * 128 + row;
    const float inv = 1.f / l;
    float* dst = static_cast<float*>(p.out) + (static_cast<size_t>(b) * p.sq + q) * p.ldo + h * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(tm_o + 32 * c, o);
      tmem_ld_wait();
      if (q < p.sq) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(dst + 32 * c)[i] =
              make_float4(__uint_as_float(o[4 * i]) * inv, __uint_as_float(o[4 * i + 1]) * inv,
                          __uint_as_float(o[4 * i + 2]) * inv, __uint_as_float(o[4 * i + 3]) * inv);
      }
    }
    if (p.lse != nullptr && q < p.sq)
      p.lse[(static_cast<size_t>(b) * p.heads + h) * p.sq + q] = m_used * sl2 * 0.69314718056f + logf(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

cudaError_t launch_attention_x3(const CUtensorMap& tq3, const CUtensorMap& tk3, const CUtensorMap& tv2,
                                const AttnArgs& a, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(attention_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X3_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const int grid = a.batch * a.heads * a.q_tiles;
  attention_x3_kernel<<<grid, X3_THREADS, X3_SMEM_BYTES, stream>>>(tq3, tk3, tv2, a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- operand preparation
// fp32 q [rows, heads*64], kv [rows, 2*heads*64] (K | V)  ->  bf16 q3 [rows, heads*192] = per head [Qhi | Qlo | Qhi],
// k3 [rows, heads*192] = per head [Khi | Khi | Klo], v2 [rows, heads*128] = per head [Vhi | Vlo].
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  hi.x = pack_bf16(v.x, v.y); hi.y = pack_bf16(v.z, v.w);
  lo.x = pack_bf16(v.x - bf16_lo(hi.x), v.y - bf16_hi(hi.x));
  lo.y = pack_bf16(v.z - bf16_lo(hi.y), v.w - bf16_hi(hi.y));
}
__global__ void __launch_bounds__(256) attn_split_kernel(const float4* __restrict__ q, int ldq4, const float4* __restrict__ kv,
                                                         int ldkv4, uint2* __restrict__ q3, uint2* __restrict__ k3,
                                                         uint2* __restrict__ v2, size_t rows_q, size_t rows_kv, int heads) {
  const int d4 = heads * 16;  // float4 vectors per row of one of q / k / v
  const size_t nq = rows_q * d4, nk = rows_kv * d4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < nq + 2 * nk;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint2 hi, lo;
    if (idx < nq) {
      const size_t r = idx / d4; const int c = idx % d4, hh = c / 16, e = c % 16;
      split4(__ldg(q + r * ldq4 + c), hi, lo);
      uint2* o = q3 + (r * heads + hh) * 48 + e;
      o[0] = hi; o[16] = lo; o[32] = hi;
    } else if (idx < nq + nk) {
      const size_t i = idx - nq, r = i / d4; const int c = i % d4, hh = c / 16, e = c % 16;
      split4(__ldg(kv + r * ldkv4 + c), hi, lo);
      uint2* o = k3 + (r * heads + hh) * 48 + e;
      o[0] = hi; o[16] = hi; o[32] = lo;
    } else {
      const size_t i = idx - nq - nk, r = i / d4; const int c = i % d4, hh = c / 16, e = c % 16;
      split4(__ldg(kv + r * ldkv4 + d4 + c), hi, lo);
      uint2* o = v2 + (r * heads + hh) * 32 + e;
      o[0] = hi; o[16] = lo;
    }
  }
}
cudaError_t launch_attn_split(const float* q, int ldq, const float* kv, int ldkv, void* q3, void* k3, void* v2,
                              size_t rows_q, size_t rows_kv, int heads, cudaStream_t stream) {
  if (ldq % 4 || ldkv % 4) return cudaErrorInvalidValue;
  const size_t total = (rows_q + 2 * rows_kv) * heads * 16;
  if (total == 0) return cudaSuccess;
  const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  attn_split_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(q), ldq / 4,
                                              reinterpret_cast<const float4*>(kv), ldkv / 4, static_cast<uint2*>(q3),
                                              static_cast<uint2*>(k3), static_cast<uint2*>(v2), rows_q, rows_kv, heads);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- GEMM operand split
// fp32 x [rows, K] -> bf16 [rows, 3K] = [hi | lo | hi] (optionally of relu(x)); the matching weight layout is
// [Whi | Whi | Wlo] along K, so the ordinary bf16 GEMM over 3K accumulates hi*hi + lo*hi + hi*lo in fp32.
__global__ void __launch_bounds__(256) split3_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t rows,
                                                     int k4, int relu) {
  const size_t total = rows * k4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = idx / k4; const int c = idx % k4;
    float4 v = __ldg(in + idx);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    uint2 hi, lo;
    split4(v, hi, lo);
    uint2* o = out + r * 3 * k4 + c;
    o[0] = hi; o[k4] = lo; o[2 * k4] = hi;
  }
}
cudaError_t launch_split3(const float* in, void* out, size_t rows, int k, int relu, cudaStream_t stream) {
  if (k % 4) return cudaErrorInvalidValue;
  const size_t total = rows * (k / 4);
  if (total == 0) return cudaSuccess;
  const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  split3_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(in), static_cast<uint2*>(out), rows, k / 4, relu);
  return cudaGetLastError();
}

// dst += src (fp32; second residual operand of the DPT fusion blocks in parity mode)
__global__ void __launch_bounds__(256) add_f32_kernel(float4* __restrict__ dst, const float4* __restrict__ src, size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 a = dst[i];
    const float4 b = __ldg(src + i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    dst[i] = a;
  }
}
cudaError_t launch_add_f32(float* dst, const float* src, size_t n, cudaStream_t stream) {
  if (n % 4) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  const size_t n4 = n / 4;
  const int grid = static_cast<int>(n4 / 256 + 1 < 148 * 16 ? n4 / 256 + 1 : 148 * 16);
  add_f32_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), n4);
  return cudaGetLastError();
}

}  // namespace f3r
