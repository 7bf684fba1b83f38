// FlashAttention-style forward for head_dim 64 on sm_100a: softmax(scale * Q K^T) V, non-causal, no mask
// (fast3r/croco/models/blocks.py:135-194; encoder: batch = views, S = P; fusion decoder: batch = B, S = N*P).
//
// One CTA owns 256 query rows (two 128-row tiles that ping-pong) of one (batch, head) and streams all keys
// in blocks of 128.  K/V blocks arrive by TMA into a 128B-swizzled smem ring shared by both tiles;
// S = Q K^T and O += P V run on tcgen05 with S, P and O all resident in TMEM (512 columns:
// S0 S1 | P0 P1 | O0 O1).  P is written back to TMEM as packed bf16 and consumed as the TMEM A-operand of the
// PV MMA, V is consumed as an MN-major smem B operand, so no transposes or smem round trips are needed.
// Softmax is exact online softmax in fp32 (exp2 domain) with lazy O rescaling: the running reference max is
// only moved when the row max grows by more than 2^8, which makes the TMEM read-modify-write of O rare.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 4-7 = softmax of tile 0,
// warps 8-11 = softmax of tile 1 (one thread per query row).  setmaxnreg moves registers from warpgroup 0
// to the softmax warpgroups, which keep a whole 128-wide score row in registers.
#include <cstdlib>

#include "common.cuh"
#include "f3r_kernels.h"

namespace f3r {

// kSplit = threads per query row in the softmax (1: 8 softmax warps, 2: 16 softmax warps, better latency hiding)
template <int kSplit> constexpr int att_threads() { return 128 + 256 * kSplit; }  // warpgroup 0: TMA + MMA (+2 idle)
constexpr int ATT_STAGES = 4;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int ATT_SMEM_BYTES = (2 + 2 * ATT_STAGES) * ATT_TILE_BYTES + 1024 + 256 + 4096 /*row-max exchange*/;

constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_P0 = 256, TM_P1 = 320, TM_O0 = 384, TM_O1 = 448;

F3R_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Blackwell packed fp32x2 / 3-input ALU ops (SASS FFMA2 / FADD2 / FMNMX3): halve the issue slots of the softmax.
F3R_DEVICE float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
F3R_DEVICE void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%4}; mov.b64 rc, {%5,%5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
F3R_DEVICE void fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{ .reg .b64 ra, rd; mov.b64 rd, {%0,%1}; mov.b64 ra, {%2,%3};\n\t"
      "add.rn.f32x2 rd, rd, ra; mov.b64 {%0,%1}, rd; }"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}

F3R_DEVICE void fsub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {  // a - b
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%6};\n\t"
      "fma.rn.f32x2 rd, rb, rc, ra; mov.b64 {%0,%1}, rd; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(-1.0f));
}
F3R_DEVICE void fma2v(float& d0, float& d1, float a0, float a1, float b0, float b1, float c) {  // a*b + c
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c));
}
// exp2 of a pair on the FMA / ALU pipes instead of the 16-per-clock-per-SM special-function unit: round-to-nearest
// split x = n + f (magic-number add), degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (max rel. error 7.6e-5, far
// below the bf16 rounding of P), exponent patched in with an integer add.  x is clamped at -126 (2^-126 ~ 0).
F3R_DEVICE void exp2_emu2(float& e0, float& e1, float x0, float x1) {
  x0 = fmaxf(x0, -126.f); x1 = fmaxf(x1, -126.f);
  float t0 = x0, t1 = x1;
  fadd2(t0, t1, 12582912.f, 12582912.f);        // t = x + 1.5*2^23: low mantissa bits hold round(x)
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
// bit k set: pair k of every 8 score pairs takes the FMA-pipe exp2.  Measured (profiles/r01_notes.md): in the isolated
// softmax stream 25 % emulation is 12 % faster (2386 -> 2110 clk / iteration), but in the full kernel the decoder layer
// at N=32 went from 2.70 ms to 2.99 ms, so it is compiled out by default (build with -DF3R_ATT_EMU_MASK=0x11 to try).
// 1: deferred row max (one-pass softmax, see the loop); 0: classic max pass before the exponentials.
// Measured at N=32: 3.06 ms per decoder layer vs 2.70 ms for the classic order (the early pv_done wait and the chunked
// P stores cost more than the max pre-pass saves), so the classic order stays the default.
#ifndef F3R_ATT_DEFER_MAX
#define F3R_ATT_DEFER_MAX 0
#endif
#ifndef F3R_ATT_EMU_MASK
#define F3R_ATT_EMU_MASK 0x0
#endif

template <int kSplit>
__global__ void __launch_bounds__(att_threads<kSplit>(), 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                 const __grid_constant__ AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // 2 tiles
  uint8_t* smem_k = smem + 2 * ATT_TILE_BYTES;              // ATT_STAGES tiles
  uint8_t* smem_v = smem_k + ATT_STAGES * ATT_TILE_BYTES;   // ATT_STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + ATT_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // ATT_STAGES
  uint64_t* k_empty = k_full + ATT_STAGES;
  uint64_t* v_full = k_empty + ATT_STAGES;
  uint64_t* v_empty = v_full + ATT_STAGES;
  uint64_t* s_full = v_empty + ATT_STAGES;       // 2
  uint64_t* s_free = s_full + 2;                 // 2
  uint64_t* p_full = s_free + 2;                 // 2
  uint64_t* pv_done = p_full + 2;                // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xbuf = reinterpret_cast<float*>(smem_v + ATT_STAGES * ATT_TILE_BYTES + 256);  // [tile][parity][half][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int nkv = (p.skv + 127) / 128;
  const int dmodel = p.heads * 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_STAGES; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&s_free[t], 128 * kSplit);
      mbar_init(&p_full[t], 128 * kSplit); mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
    // register budget (must balance inside the CTA's launch allocation):
    //   kSplit 1: 384 thr x 168 = 64512 = 128 x 72 + 256 x 216      kSplit 2: 640 thr x 96 = 61440 = 128 x 64 + 512 x 104
    if constexpr (kSplit == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      tma_load_3d(smem_q, &tmap_q, q_full, h * 64, qt * 256, b);
      tma_load_3d(smem_q + ATT_TILE_BYTES, &tmap_q, q_full, h * 64, qt * 256 + 128, b);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&k_full[stage], ATT_TILE_BYTES);
        tma_load_3d(smem_k + stage * ATT_TILE_BYTES, &tmap_kv, &k_full[stage], h * 64, j * 128, b);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&v_full[stage], ATT_TILE_BYTES);
        tma_load_3d(smem_v + stage * ATT_TILE_BYTES, &tmap_kv, &v_full[stage], dmodel + h * 64, j * 128, b);
        if (++stage == ATT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);  // S[128x128] = Q[128x64] K^T, both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);   // O[128x64] += P[128x128] V, V MN-major
      const uint64_t qdesc0 = make_smem_desc_sw128(smem_u32(smem_q), 1);
      const uint64_t qdesc1 = make_smem_desc_sw128(smem_u32(smem_q + ATT_TILE_BYTES), 1);
      auto issue_s = [&](int t, int stage) {
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(smem_k + stage * ATT_TILE_BYTES), 1);
        const uint64_t qd = t ? qdesc1 : qdesc0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_base + (t ? TM_S1 : TM_S0), qd + 2 * k, kdesc + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int stage, int j) {
        const uint64_t vdesc = make_smem_desc_sw128(smem_u32(smem_v + stage * ATT_TILE_BYTES), 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // 16 keys per MMA: 8 packed-bf16 TMEM columns of P, 16 smem rows (2048 B) of V
          umma_ts(tmem_base + (t ? TM_O1 : TM_O0), tmem_base + (t ? TM_P1 : TM_P0) + 8 * k, vdesc + 128 * k,
                  idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&pv_done[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % ATT_STAGES;
        const uint32_t ph = (j / ATT_STAGES) & 1;
        const int st1 = (j + 1) % ATT_STAGES;
        const uint32_t ph1 = ((j + 1) / ATT_STAGES) & 1;
        const bool more = (j + 1) < nkv;
        if (more) {
          mbar_wait(&k_full[st1], ph1);
          mbar_wait(&s_free[0], j & 1);
          tc_fence_after();
          issue_s(0, st1);
        }
        mbar_wait(&v_full[st], ph);
        mbar_wait(&p_full[0], j & 1);
        tc_fence_after();
        issue_pv(0, st, j);
        if (more) {
          mbar_wait(&s_free[1], j & 1);
          tc_fence_after();
          issue_s(1, st1);
          umma_commit(&k_empty[st1]);
        }
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        issue_pv(1, st, j);
        umma_commit(&v_empty[st]);
      }
    }
  }
  } else {
    // ===================== softmax warps =====================
    if constexpr (kSplit == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    constexpr int COLS = 128 / kSplit;      // score columns per thread
    constexpr int OCOLS = 64 / kSplit;      // output columns per thread
    const int sw = warp - 4;
    const int t = sw / (4 * kSplit);        // query tile 0 / 1
    const int half = (sw >> 2) % kSplit;    // which column slice of the row this thread owns
    const int quarter = warp & 3;           // TMEM lane quarter accessible to this warp
    const int row = quarter * 32 + lane;    // row in the 128-row tile
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tm_s = tmem_base + lane_base + (t ? TM_S1 : TM_S0) + half * COLS;
    const uint32_t tm_p = tmem_base + lane_base + (t ? TM_P1 : TM_P0) + half * (COLS / 2);
    const uint32_t tm_o = tmem_base + lane_base + (t ? TM_O1 : TM_O0) + half * OCOLS;
    const float sl2 = p.scale_log2;
    float m_used = -INFINITY;  // raw-score reference max the exponentials are taken against
    float l = 0.f;             // this thread's partial row sum

    float mx_prev = -INFINITY;  // (deferred-max variant) raw row max of the previous key block
    for (int j = 0; j < nkv; ++j) {
      if constexpr (kSplit == 1 && F3R_ATT_DEFER_MAX) {
        // ---- one-pass variant: the exponentials of block j are taken against the reference max known BEFORE the block
        // (decided from block j-1's max), so the row-max reduction runs inside the exp loop on the ALU pipe instead of
        // as a serial pre-pass.  Exact: P_j, l and O always share one reference; the reference moves one block late.
        // If block j overshoots the reference by more than 2^64 the block is recomputed against its own max.
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t s[128];
        tmem_ld32(tm_s + 0, s + 0);
        tmem_ld32(tm_s + 32, s + 32);
        tmem_ld32(tm_s + 64, s + 64);
        tmem_ld32(tm_s + 96, s + 96);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[t]);
        if (j == nkv - 1) {
          const int valid = p.skv - j * 128;
          if (valid < 128) {
#pragma unroll
            for (int i = 0; i < 128; ++i)
              if (i >= valid) s[i] = 0xff800000u;
          }
        }
        auto rowmax = [&]() {
          float a0 = max3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
          float a1 = max3(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
          float a2 = __uint_as_float(s[6]), a3 = __uint_as_float(s[7]);
#pragma unroll
          for (int i = 8; i < 128; i += 8) {
            a0 = max3(a0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
            a1 = max3(a1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
            a2 = max3(a2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
            a3 = max3(a3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
          }
          return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        };
        auto rescale_o = [&](float alpha) {  // O_t *= alpha (PV_{j-1} must have completed)
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
        };
        if (j == 0) {
          m_used = rowmax();  // the first block has no earlier reference
        } else {
          float alpha = 1.f;
          const bool need = (mx_prev - m_used) * sl2 > 8.f;
          if (need) {
            alpha = ex2_approx((m_used - mx_prev) * sl2);
            m_used = mx_prev;
            l *= alpha;
          }
          mbar_wait(&pv_done[t], (j - 1) & 1);  // O_t quiescent and P_t consumed (needed before the P stores anyway)
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) rescale_o(alpha);
        }
        const float l_before = l;
        float mxc0, mxc1;
        auto exp_pass = [&](bool track) {
          const float nm = -m_used * sl2;
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
          mxc0 = -INFINITY; mxc1 = -INFINITY;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 32 * c; i < 32 * c + 32; i += 4) {
              float x0, x1, x2, x3;
              ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sl2, nm);
              ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sl2, nm);
              if (track) {
                mxc0 = max3(mxc0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
                mxc1 = max3(mxc1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
              }
              const float e0 = ex2_approx(x0), e1 = ex2_approx(x1), e2 = ex2_approx(x2), e3 = ex2_approx(x3);
              fadd2(l0, l1, e0, e1);
              fadd2(l2, l3, e2, e3);
              pk[(i - 32 * c) / 2] = pack_bf16(e0, e1);
              pk[(i - 32 * c) / 2 + 1] = pack_bf16(e2, e3);
            }
            tmem_st16(tm_p + 16 * c, pk);
          }
          return (l0 + l1) + (l2 + l3);
        };
        float lsum = exp_pass(j > 0);
        if (j > 0) {
          const float mx_cur = fmaxf(mxc0, mxc1);
          mx_prev = mx_cur;
          const bool over = (mx_cur - m_used) * sl2 > 64.f;  // reference too stale for this block: redo it exactly
          if (__any_sync(0xffffffffu, over)) {
            float alpha = 1.f;
            if (over) { alpha = ex2_approx((m_used - mx_cur) * sl2); m_used = mx_cur; }
            l = l_before * alpha;
            tmem_st_wait();
            rescale_o(alpha);
            lsum = exp_pass(false);
          }
        } else {
          mx_prev = m_used;
        }
        l += lsum;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        continue;
      }
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s[COLS];
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c) tmem_ld32(tm_s + 32 * c, s + 32 * c);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[t]);  // S_t may be overwritten by the next QK^T

      if (j == nkv - 1) {
        const int valid = p.skv - j * 128 - half * COLS;
        if (valid < COLS) {
#pragma unroll
          for (int i = 0; i < COLS; ++i)
            if (i >= valid) s[i] = 0xff800000u;  // -inf
        }
      }
      float mx0 = max3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
      float mx1 = max3(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
      float mx2 = __uint_as_float(s[6]), mx3 = __uint_as_float(s[7]);
#pragma unroll
      for (int i = 8; i < COLS; i += 8) {
        mx0 = max3(mx0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        mx1 = max3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = max3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = max3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if constexpr (kSplit == 2) {
        // combine the two half-row maxima (both threads must take the same rescale decision)
        float* xb = xbuf + ((t * 2 + (j & 1)) * 2) * 128;
        xb[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
        mx = fmaxf(mx, xb[(half ^ 1) * 128 + row]);
      }
      // lazy rescale: move the reference only if the max grew by more than 8 (log2 domain)
      float alpha = 1.f;
      const bool need = (mx - m_used) * sl2 > 8.f;  // (-inf reference => true)
      if (need) {
        alpha = ex2_approx((m_used - mx) * sl2);  // exp2(-inf) = 0 on the first block
        m_used = mx;
        l *= alpha;
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        mbar_wait(&pv_done[t], (j - 1) & 1);  // O_t must be quiescent
        tc_fence_after();
        uint32_t o[32];
#pragma unroll
        for (int c = 0; c < OCOLS / 32; ++c) {
          tmem_ld32(tm_o + 32 * c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tm_o + 32 * c, o);
        }
        tmem_st_wait();
      }
      const float nm = -m_used * sl2;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      uint32_t pk[COLS / 2];
#pragma unroll
      for (int i = 0; i < COLS; i += 4) {
        float x0, x1, x2, x3;
        ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sl2, nm);
        ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sl2, nm);
        float e0, e1, e2, e3;
        if ((F3R_ATT_EMU_MASK >> ((i / 2) & 7)) & 1) exp2_emu2(e0, e1, x0, x1);
        else { e0 = ex2_approx(x0); e1 = ex2_approx(x1); }
        if ((F3R_ATT_EMU_MASK >> ((i / 2 + 1) & 7)) & 1) exp2_emu2(e2, e3, x2, x3);
        else { e2 = ex2_approx(x2); e3 = ex2_approx(x3); }
        fadd2(l0, l1, e0, e1);
        fadd2(l2, l3, e2, e3);
        pk[i / 2] = pack_bf16(e0, e1);
        pk[i / 2 + 1] = pack_bf16(e2, e3);
      }
      l += (l0 + l1) + (l2 + l3);
      if (j > 0) {
        mbar_wait(&pv_done[t], (j - 1) & 1);  // previous PV has consumed P_t
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < COLS / 64; ++c) tmem_st32(tm_p + 32 * c, pk + 32 * c);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }

    // ---- epilogue: O / l -> bf16 -> global
    if constexpr (kSplit == 2) {
      float* xb = xbuf + ((t * 2 + (nkv & 1)) * 2) * 128;  // slot not used by the last iteration's exchange
      xb[half * 128 + row] = l;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      l += xb[(half ^ 1) * 128 + row];
    }
    mbar_wait(&pv_done[t], (nkv - 1) & 1);
    tc_fence_after();
    const int q = qt * 256 + t * 128 + row;
    const float inv = 1.f / l;
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + (static_cast<size_t>(b) * p.sq + q) * p.ldo + h * 64 +
                         half * OCOLS;
#pragma unroll
    for (int c = 0; c < OCOLS / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(tm_o + 32 * c, o);
      tmem_ld_wait();
      if (q < p.sq) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + 32 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          d4[i] = w;
        }
      }
    }
    if (p.lse != nullptr && q < p.sq && half == 0)
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

template <int kSplit>
static cudaError_t launch_attention_t(const CUtensorMap& tq, const CUtensorMap& tkv, const AttnArgs& a,
                                      cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ATT_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = a.batch * a.heads * a.q_tiles;
  attention_kernel<kSplit><<<grid, att_threads<kSplit>(), ATT_SMEM_BYTES, stream>>>(tq, tkv, a);
  return cudaGetLastError();
}

cudaError_t launch_attention(const CUtensorMap& tq, const CUtensorMap& tkv, const AttnArgs& a, cudaStream_t stream) {
  static int variant = -1;
  if (variant < 0) {
    // default: one thread per score row (measured 2.68 ms vs 3.11 ms per decoder layer at N=32 for the
    // two-threads-per-row variant, profiles/r01_notes.md); F3R_ATTN_SPLIT=2 selects the latter for experiments
    const char* e = getenv("F3R_ATTN_SPLIT");
    variant = (e && e[0] == '2') ? 2 : 1;
  }
  return variant == 1 ? launch_attention_t<1>(tq, tkv, a, stream) : launch_attention_t<2>(tq, tkv, a, stream);
}

}  // namespace f3r
