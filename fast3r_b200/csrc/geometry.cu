// Geometry tail that every caller runs right after the forward (first slice):
//   conf_quantile      torch.quantile(conf.reshape(-1), q) per view            (multiview_dust3r_module.py:477, :1093)
//   similarity_fit     mask = conf >= thr & valid (with the two "< 3 points" fallbacks) and the least-squares
//                      similarity y ~ s R x + t over the masked points        (multiview_dust3r_module.py:427-525;
//                      roma.rigid_points_registration(compute_scaling=True) = Umeyama / Kabsch)
//   similarity_apply   out = s (x R^T) + t on ALL points                       (multiview_dust3r_module.py:517-521)
//   focal_weiszfeld    IRLS focal from a pointmap                              (dust3r/post_process.py:19-79, :82-142)
// All of it is HBM-bound streaming / reduction work: one pass over 12-28 bytes per pixel per kernel, fp64 accumulators,
// fixed reduction order (no float atomics), so results do not depend on scheduling.
#include <math.h>

#include "f3r_kernels.h"
#include "geometry_math.h"

namespace f3r {

// ------------------------------------------------------------------------------------------------- quantile
// One CTA per view: 4-pass 8-bit radix select of the order statistic floor(q (n-1)) on the order-preserving integer image
// of the floats, one more pass for its successor, then ATen's lerp.  The histogram is warp-aggregated (confidences share
// their exponent byte, so naive shared atomics would serialise on one bin).
namespace {

constexpr int QT = 1024;

__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void __launch_bounds__(QT) conf_quantile_kernel(const float* __restrict__ conf, int n, float q,
                                                           float* __restrict__ thr) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_k, s_cnt_le, s_min_gt;
  const float* c = conf + static_cast<size_t>(blockIdx.x) * n;
  const int tid = threadIdx.x;
  const float rank = __fmul_rn(q, static_cast<float>(n - 1));
  const int lo = static_cast<int>(floorf(rank));
  const int hi = static_cast<int>(ceilf(rank));
  const float w = __fsub_rn(rank, static_cast<float>(lo));

  uint32_t prefix = 0, mask = 0, k = static_cast<uint32_t>(lo);
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int base = 0; base < n; base += QT) {
      const int i = base + tid;
      bool ok = false;
      uint32_t d = 0;
      if (i < n) {
        const uint32_t u = fkey(c[i]);
        ok = (u & mask) == prefix;
        d = (u >> shift) & 255u;
      }
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const unsigned peers = __match_any_sync(act, d);
        if ((__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[d], static_cast<uint32_t>(__popc(peers)));
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t cum = 0;
      int d = 0;
      for (; d < 255; ++d) {
        const uint32_t h = hist[d];
        if (cum + h > k) break;
        cum += h;
      }
      s_prefix = prefix | (static_cast<uint32_t>(d) << shift);
      s_k = k - cum;
    }
    __syncthreads();
    prefix = s_prefix;
    k = s_k;
    mask |= 0xffu << shift;
  }
  const uint32_t u_lo = prefix;
  if (hi == lo) {  // integral rank: ATen still evaluates lerp(a, a, 0) = fma(0, a - a, a), which is NaN for an infinite a
    if (tid == 0) {
      const float a = fkey_inv(u_lo);
      thr[blockIdx.x] = fmaf(w, __fsub_rn(a, a), a);
    }
    return;
  }
  if (tid == 0) {
    s_cnt_le = 0;
    s_min_gt = 0xffffffffu;
  }
  __syncthreads();
  uint32_t cnt = 0, mn = 0xffffffffu;
  for (int i = tid; i < n; i += QT) {
    const uint32_t u = fkey(c[i]);
    if (u <= u_lo) ++cnt;
    else mn = min(mn, u);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  mn = __reduce_min_sync(0xffffffffu, mn);
  if ((tid & 31) == 0) {
    atomicAdd(&s_cnt_le, cnt);
    atomicMin(&s_min_gt, mn);
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t u_hi = (static_cast<uint32_t>(hi) < s_cnt_le) ? u_lo : s_min_gt;
    const float a = fkey_inv(u_lo), b = fkey_inv(u_hi);
    const float diff = __fsub_rn(b, a);
    // ATen lerp (Lerp.h; fused multiply-add on both the vectorised CPU path and CUDA):
    // |w| < 0.5 ? fma(w, b - a, a) : fma(w - 1, b - a, b)
    thr[blockIdx.x] = (fabsf(w) < 0.5f) ? fmaf(w, diff, a) : fmaf(__fsub_rn(w, 1.0f), diff, b);
  }
}

// ------------------------------------------------------------------------------------------------- similarity fit
// MOM (17 moments per point set) comes from geometry_math.h
constexpr int FIT_THREADS = 256;
constexpr int FIT_CHUNKS = 32;   // partial sums per view (fixed, so the reduction order is)

struct Moments {
  double v[MOM];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < MOM; ++i) v[i] = 0.0;
  }
  __device__ __forceinline__ void add(float x0, float x1, float x2, float y0, float y1, float y2) {
    const double a0 = x0, a1 = x1, a2 = x2, b0 = y0, b1 = y1, b2 = y2;
    v[0] += 1.0;
    v[1] += a0; v[2] += a1; v[3] += a2;
    v[4] += b0; v[5] += b1; v[6] += b2;
    v[7] += a0 * a0 + a1 * a1 + a2 * a2;
    v[8] += b0 * a0;  v[9] += b0 * a1;  v[10] += b0 * a2;
    v[11] += b1 * a0; v[12] += b1 * a1; v[13] += b1 * a2;
    v[14] += b2 * a0; v[15] += b2 * a1; v[16] += b2 * a2;
  }
};

// set A: conf >= thr & valid;  set B: valid only (the reference's first fallback).  partial [views][FIT_CHUNKS][2][MOM]
__global__ void __launch_bounds__(FIT_THREADS) similarity_moments_kernel(
    const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ conf,
    const float* __restrict__ thr, const uint8_t* __restrict__ valid, int n, double* __restrict__ partial) {
  __shared__ double red[FIT_THREADS / 32][2 * MOM];
  const int view = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const size_t base = static_cast<size_t>(view) * n;
  const float* xv = x + base * 3;
  const float* yv = y + base * 3;
  const float* cv = conf ? conf + base : nullptr;
  const uint8_t* vv = valid ? valid + base : nullptr;
  const float t = (cv && thr) ? thr[view] : 0.f;
  const int per = (n + FIT_CHUNKS - 1) / FIT_CHUNKS;
  const int i0 = chunk * per, i1 = min(n, i0 + per);
  Moments a, b;
  a.zero();
  b.zero();
  for (int i = i0 + tid; i < i1; i += FIT_THREADS) {
    const bool in_b = vv ? (vv[i] != 0) : true;
    if (!in_b) continue;
    const float x0 = xv[3 * i], x1 = xv[3 * i + 1], x2 = xv[3 * i + 2];
    const float y0 = yv[3 * i], y1 = yv[3 * i + 1], y2 = yv[3 * i + 2];
    b.add(x0, x1, x2, y0, y1, y2);
    if (!(cv && thr) || cv[i] >= t) a.add(x0, x1, x2, y0, y1, y2);
  }
#pragma unroll
  for (int j = 0; j < MOM; ++j) {
    double va = a.v[j], vb = b.v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      va += __shfl_down_sync(0xffffffffu, va, o);
      vb += __shfl_down_sync(0xffffffffu, vb, o);
    }
    if ((tid & 31) == 0) {
      red[tid >> 5][j] = va;
      red[tid >> 5][MOM + j] = vb;
    }
  }
  __syncthreads();
  if (tid < 2 * MOM) {
    double s = 0.0;
#pragma unroll
    for (int wp = 0; wp < FIT_THREADS / 32; ++wp) s += red[wp][tid];
    partial[(static_cast<size_t>(view) * FIT_CHUNKS + chunk) * 2 * MOM + tid] = s;
  }
}

__global__ void __launch_bounds__(64) similarity_solve_kernel(const double* __restrict__ partial, float* __restrict__ rts) {
  __shared__ double tot[2 * MOM];
  const int view = blockIdx.x, tid = threadIdx.x;
  if (tid < 2 * MOM) {
    double s = 0.0;
    for (int c = 0; c < FIT_CHUNKS; ++c) s += partial[(static_cast<size_t>(view) * FIT_CHUNKS + c) * 2 * MOM + tid];
    tot[tid] = s;
  }
  __syncthreads();
  if (tid != 0) return;
  float* out = rts + static_cast<size_t>(view) * 13;
  const double* m = nullptr;
  if (tot[0] >= 3.0) m = tot;                  // confidence mask & valid_mask
  else if (tot[MOM] >= 3.0) m = tot + MOM;     // valid_mask only (multiview_dust3r_module.py:493-501)
  if (!m) {                                    // identity (:504-509)
    for (int i = 0; i < 13; ++i) out[i] = 0.f;
    out[0] = out[4] = out[8] = 1.f;
    out[12] = 1.f;
    return;
  }
  umeyama_from_moments(m, out);
}

// ------------------------------------------------------------------------------------------------- similarity apply
constexpr int APPLY_THREADS = 256;

__device__ __forceinline__ void apply_pt(const float* r, float x0, float x1, float x2, float& o0, float& o1, float& o2) {
  o0 = fmaf(r[12], fmaf(r[2], x2, fmaf(r[1], x1, r[0] * x0)), r[9]);
  o1 = fmaf(r[12], fmaf(r[5], x2, fmaf(r[4], x1, r[3] * x0)), r[10]);
  o2 = fmaf(r[12], fmaf(r[8], x2, fmaf(r[7], x1, r[6] * x0)), r[11]);
}

// kVec: n % 4 == 0 and 16-byte aligned bases - each thread moves 4 points as 3 float4
template <bool kVec>
__global__ void __launch_bounds__(APPLY_THREADS) similarity_apply_kernel(const float* __restrict__ x,
                                                                         const float* __restrict__ rts,
                                                                         float* __restrict__ out, int n) {
  __shared__ float r[13];
  const int view = blockIdx.y;
  if (threadIdx.x < 13) r[threadIdx.x] = rts[static_cast<size_t>(view) * 13 + threadIdx.x];
  __syncthreads();
  const size_t base = static_cast<size_t>(view) * n * 3;
  if (kVec) {
    const int g = blockIdx.x * APPLY_THREADS + threadIdx.x;  // group of 4 points
    if (g * 4 >= n) return;
    const float4* src = reinterpret_cast<const float4*>(x + base) + static_cast<size_t>(g) * 3;
    float4* dst = reinterpret_cast<float4*>(out + base) + static_cast<size_t>(g) * 3;
    const float4 a = __ldcs(src), b = __ldcs(src + 1), c = __ldcs(src + 2);
    float4 oa, ob, oc;
    apply_pt(r, a.x, a.y, a.z, oa.x, oa.y, oa.z);
    apply_pt(r, a.w, b.x, b.y, oa.w, ob.x, ob.y);
    apply_pt(r, b.z, b.w, c.x, ob.z, ob.w, oc.x);
    apply_pt(r, c.y, c.z, c.w, oc.y, oc.z, oc.w);
    __stcs(dst, oa);
    __stcs(dst + 1, ob);
    __stcs(dst + 2, oc);
  } else {
    const int i = blockIdx.x * APPLY_THREADS + threadIdx.x;
    if (i >= n) return;
    const float* p = x + base + static_cast<size_t>(i) * 3;
    float* o = out + base + static_cast<size_t>(i) * 3;
    float o0, o1, o2;
    apply_pt(r, p[0], p[1], p[2], o0, o1, o2);
    o[0] = o0;
    o[1] = o1;
    o[2] = o2;
  }
}

// ------------------------------------------------------------------------------------------------- Weiszfeld focal
constexpr int FOC_THREADS = 256;
constexpr int FOC_CHUNKS = 64;   // partial sums per view and iteration
constexpr int FOC_P = 3;         // numerator, denominator, selected points

// One IRLS iteration over all views: every block first re-derives the current focal of its view from the previous
// iteration's partial sums (prev == NULL: the closed-form L2 initialisation, unit weights), then reduces its chunk.
__global__ void __launch_bounds__(FOC_THREADS) weiszfeld_iter_kernel(
    const float* __restrict__ pts, const float* __restrict__ conf, const float* __restrict__ thr,
    const float* __restrict__ pp, int H, int W, const double* __restrict__ prev, double* __restrict__ next) {
  __shared__ double red[FOC_THREADS / 32][FOC_P];
  const int view = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int n = H * W;
  float focal = 0.f;
  if (prev) {
    double num = 0.0, den = 0.0;
    const double* pv = prev + static_cast<size_t>(view) * FOC_CHUNKS * FOC_P;
    for (int c = 0; c < FOC_CHUNKS; ++c) {
      num += pv[c * FOC_P];
      den += pv[c * FOC_P + 1];
    }
    focal = static_cast<float>(num / den);
  }
  const float cx = pp ? pp[2 * view] : 0.5f * static_cast<float>(W);
  const float cy = pp ? pp[2 * view + 1] : 0.5f * static_cast<float>(H);
  const size_t base = static_cast<size_t>(view) * n;
  const float* pv3 = pts + base * 3;
  const float* cv = (conf && thr) ? conf + base : nullptr;
  const float t = cv ? thr[view] : 0.f;
  const int per = (n + FOC_CHUNKS - 1) / FOC_CHUNKS;
  const int i0 = chunk * per, i1 = min(n, i0 + per);
  double num = 0.0, den = 0.0, cnt = 0.0;
  for (int i = i0 + tid; i < i1; i += FOC_THREADS) {
    if (cv && !(cv[i] >= t)) continue;
    const float x = pv3[3 * i], y = pv3[3 * i + 1], z = pv3[3 * i + 2];
    float xz = __fdiv_rn(x, z), yz = __fdiv_rn(y, z);
    if (!isfinite(xz)) xz = 0.f;   // nan_to_num(posinf=0, neginf=0), NaN -> 0
    if (!isfinite(yz)) yz = 0.f;
    const float u = static_cast<float>(i % W) - cx, v = static_cast<float>(i / W) - cy;
    const float dpx = __fadd_rn(__fmul_rn(xz, u), __fmul_rn(yz, v));
    const float dxx = __fadd_rn(__fmul_rn(xz, xz), __fmul_rn(yz, yz));
    float wgt = 1.f;
    if (prev) {
      const float du = u - focal * xz, dv = v - focal * yz;
      wgt = 1.f / fmaxf(sqrtf(du * du + dv * dv), 1e-8f);
    }
    num += static_cast<double>(wgt * dpx);
    den += static_cast<double>(wgt * dxx);
    cnt += 1.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_down_sync(0xffffffffu, num, o);
    den += __shfl_down_sync(0xffffffffu, den, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if ((tid & 31) == 0) {
    red[tid >> 5][0] = num;
    red[tid >> 5][1] = den;
    red[tid >> 5][2] = cnt;
  }
  __syncthreads();
  if (tid < FOC_P) {
    double s = 0.0;
#pragma unroll
    for (int wp = 0; wp < FOC_THREADS / 32; ++wp) s += red[wp][tid];
    next[(static_cast<size_t>(view) * FOC_CHUNKS + chunk) * FOC_P + tid] = s;
  }
}

__global__ void weiszfeld_final_kernel(const double* __restrict__ last, int views, int H, int W, float* __restrict__ focal) {
  const int view = blockIdx.x * blockDim.x + threadIdx.x;
  if (view >= views) return;
  double num = 0.0, den = 0.0, cnt = 0.0;
  const double* pv = last + static_cast<size_t>(view) * FOC_CHUNKS * FOC_P;
  for (int c = 0; c < FOC_CHUNKS; ++c) {
    num += pv[c * FOC_P];
    den += pv[c * FOC_P + 1];
    cnt += pv[c * FOC_P + 2];
  }
  float f;
  if (cnt == 0.0) {
    f = static_cast<float>(static_cast<double>(max(H, W)) / (2.0 * tan(M_PI / 6.0)));  // post_process.py:108
  } else {
    f = static_cast<float>(num / den);
    if (f < 0.f) f = 0.f;  // focal.clip(min=0 * focal_base, max=inf)
  }
  focal[view] = f;
}

}  // namespace

// ------------------------------------------------------------------------------------------------- launchers
cudaError_t launch_conf_quantile(const float* conf, int views, int n, float q, float* thr, cudaStream_t stream) {
  conf_quantile_kernel<<<views, QT, 0, stream>>>(conf, n, q, thr);
  return cudaGetLastError();
}

size_t similarity_fit_workspace(int views) { return static_cast<size_t>(views) * FIT_CHUNKS * 2 * MOM * sizeof(double); }

cudaError_t launch_similarity_fit(const float* x, const float* y, const float* conf, const float* thr,
                                  const uint8_t* valid, int views, int n, float* rts, double* workspace,
                                  cudaStream_t stream) {
  similarity_moments_kernel<<<dim3(FIT_CHUNKS, views), FIT_THREADS, 0, stream>>>(x, y, conf, thr, valid, n, workspace);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  similarity_solve_kernel<<<views, 64, 0, stream>>>(workspace, rts);
  return cudaGetLastError();
}

cudaError_t launch_similarity_apply(const float* x, const float* rts, float* out, int views, int n, cudaStream_t stream) {
  const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    const int groups = n / 4;
    similarity_apply_kernel<true><<<dim3((groups + APPLY_THREADS - 1) / APPLY_THREADS, views), APPLY_THREADS, 0, stream>>>(
        x, rts, out, n);
  } else {
    similarity_apply_kernel<false><<<dim3((n + APPLY_THREADS - 1) / APPLY_THREADS, views), APPLY_THREADS, 0, stream>>>(
        x, rts, out, n);
  }
  return cudaGetLastError();
}

size_t focal_workspace(int views) { return 2 * static_cast<size_t>(views) * FOC_CHUNKS * FOC_P * sizeof(double); }

cudaError_t launch_focal_weiszfeld(const float* pts, const float* conf, const float* thr, const float* pp, int views,
                                   int H, int W, int iters, float* focal, double* workspace, cudaStream_t stream) {
  const size_t half = static_cast<size_t>(views) * FOC_CHUNKS * FOC_P;
  double* buf[2] = {workspace, workspace + half};
  const dim3 grid(FOC_CHUNKS, views);
  for (int it = 0; it <= iters; ++it) {
    weiszfeld_iter_kernel<<<grid, FOC_THREADS, 0, stream>>>(pts, conf, thr, pp, H, W, it ? buf[(it - 1) & 1] : nullptr,
                                                            buf[it & 1]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  weiszfeld_final_kernel<<<(views + 127) / 128, 128, 0, stream>>>(buf[iters & 1], views, H, W, focal);
  return cudaGetLastError();
}

}  // namespace f3r
