// Shared host-side plumbing for the C-ABI: per-thread error string, launch counter, 16-bit helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/viscy_b200.h"

namespace vb {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(VB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return VB200_OK;
}

// Per-device one-time work (cudaFuncSetAttribute opt-ins, attribute queries): function attributes are per device /
// context, so a process that drives several GPUs must repeat them on each; thread-safe (autograd's backward threads).
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  static int device() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 63;
  }
  bool need(int dev) const { return ((mask.load(std::memory_order_acquire) >> dev) & 1ull) == 0; }
  void done(int dev) { mask.fetch_or(1ull << dev, std::memory_order_release); }
};

#define VB_REQUIRE(cond, ...) \
  do {                        \
    if (!(cond)) return vb::fail(VB200_ERR_INVALID, __VA_ARGS__); \
  } while (0)
#define VB_SUPPORTED(cond, ...) \
  do {                          \
    if (!(cond)) return vb::fail(VB200_ERR_UNSUPPORTED, __VA_ARGS__); \
  } while (0)

// ---- 16-bit element traits ---------------------------------------------------------------------
template <bool BF16>
struct H16;
template <>
struct H16<true> {
  using T = __nv_bfloat16;
  using T2 = __nv_bfloat162;
  static __device__ __forceinline__ float to_f(T v) { return __bfloat162float(v); }
  static __device__ __forceinline__ T from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  }
};
template <>
struct H16<false> {
  using T = __half;
  using T2 = __half2;
  static __device__ __forceinline__ float to_f(T v) { return __half2float(v); }
  static __device__ __forceinline__ T from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 r = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t v) {
    return __half22float2(*reinterpret_cast<__half2*>(&v));
  }
};

// ---- runtime-typed element access (dt: VB200_BF16 | VB200_FP16 | VB200_FP32) for the loss / prediction kernels ----
__device__ __forceinline__ float ld_any(const void* p, long long i, int dt) {
  if (dt == 2) return __ldg(reinterpret_cast<const float*>(p) + i);
  if (dt == 0) return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(p) + i));
  return __half2float(__ldg(reinterpret_cast<const __half*>(p) + i));
}
__device__ __forceinline__ float round_any(float v, int dt) {
  if (dt == 2) return v;
  if (dt == 0) return __bfloat162float(__float2bfloat16_rn(v));
  return __half2float(__float2half_rn(v));
}
__device__ __forceinline__ void st_any(void* p, long long i, int dt, float v) {
  if (dt == 2) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == 0) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}

// exact-form (erf) GELU, as timm's act_layer='gelu' -> nn.GELU().  erf through
// Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below 16-bit output rounding) so that GELU and its
// derivative share one exp:  erf(x) = 1 - (a1 t + ... + a5 t^5) e^{-x^2},  t = 1/(1 + p x),  x >= 0.
__device__ __forceinline__ void gelu_parts(float u, float& cdf, float& pdf) {
  const float e = __expf(-0.5f * u * u);
  const float x = fabsf(u) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, e, 1.0f);
  cdf = 0.5f * (1.0f + copysignf(erf_abs, u));
  pdf = 0.3989422804014327f * e;
}
// two elements at a time on the packed fp32x2 pipe (sm_100 FFMA2 / FMUL2 / FADD2)
__device__ __forceinline__ void gelu_parts2(float2 u, float2& cdf, float2& pdf) {
  const float2 uu = __fmul2_rn(u, u);
  float2 e;
  e.x = exp2f(-0.72134752044448170f * uu.x);  // exp(-u^2/2) = 2^(-u^2 / (2 ln 2))
  e.y = exp2f(-0.72134752044448170f * uu.y);
  const float2 x = make_float2(fabsf(u.x), fabsf(u.y));
  const float2 den = __ffma2_rn(x, make_float2(0.23164190541f, 0.23164190541f), make_float2(1.f, 1.f));  // p/sqrt(2)
  float2 t;
  t.x = __fdividef(1.0f, den.x);
  t.y = __fdividef(1.0f, den.y);
  float2 poly = __ffma2_rn(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
  poly = __ffma2_rn(poly, t, make_float2(1.421413741f, 1.421413741f));
  poly = __ffma2_rn(poly, t, make_float2(-0.284496736f, -0.284496736f));
  poly = __ffma2_rn(poly, t, make_float2(0.254829592f, 0.254829592f));
  poly = __fmul2_rn(poly, t);
  // 0.5 * erfc(|u|/sqrt2) = 0.5 * poly * e ;  cdf = u >= 0 ? 1 - that : that
  const float2 q = __fmul2_rn(__fmul2_rn(poly, e), make_float2(0.5f, 0.5f));
  cdf.x = u.x >= 0.f ? 1.0f - q.x : q.x;
  cdf.y = u.y >= 0.f ? 1.0f - q.y : q.y;
  pdf = __fmul2_rn(e, make_float2(0.3989422804014327f, 0.3989422804014327f));
}

// raw MUFU ops: no range fix-up code around them (the arguments here are always in range)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// u -> d = gelu'(u), g = gelu(u) for two elements, one exp and one reciprocal each, everything else on the packed
// fp32x2 pipe.  With q = erfc(|u|/sqrt2)/2 (Abramowitz-Stegun 7.1.26, coefficients pre-scaled by -1/2) and h = 1/2 - q:
//   cdf = 1/2 + sgn(u) h,   gelu = u/2 + |u| h,   gelu' = cdf + u pdf = 1/2 + sgn(u) (h + |u| pdf)
__device__ __forceinline__ void gelu_gp2(float2 u, float2& d, float2& g) {
  const float2 uu = __fmul2_rn(u, u);
  const float2 ea = __fmul2_rn(uu, make_float2(-0.72134752044448170f, -0.72134752044448170f));
  const float2 e = make_float2(ex2_approx(ea.x), ex2_approx(ea.y));  // exp(-u^2/2)
  const float2 x = make_float2(fabsf(u.x), fabsf(u.y));
  const float2 den = __ffma2_rn(x, make_float2(0.23164190541f, 0.23164190541f), make_float2(1.f, 1.f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 poly = __ffma2_rn(t, make_float2(-0.5307027145f, -0.5307027145f), make_float2(0.7265760135f, 0.7265760135f));
  poly = __ffma2_rn(poly, t, make_float2(-0.7107068705f, -0.7107068705f));
  poly = __ffma2_rn(poly, t, make_float2(0.142248368f, 0.142248368f));
  poly = __ffma2_rn(poly, t, make_float2(-0.127414796f, -0.127414796f));
  poly = __fmul2_rn(poly, t);                                       // -q / e
  const float2 h = __ffma2_rn(poly, e, make_float2(0.5f, 0.5f));    // 1/2 - q
  const float2 pdf = __fmul2_rn(e, make_float2(0.3989422804014327f, 0.3989422804014327f));
  const float2 s = __ffma2_rn(x, pdf, h);
  d.x = 0.5f + copysignf(s.x, u.x);
  d.y = 0.5f + copysignf(s.y, u.y);
  g = __ffma2_rn(x, h, __fmul2_rn(u, make_float2(0.5f, 0.5f)));
}

__device__ __forceinline__ float gelu_f(float u) {
  float cdf, pdf;
  gelu_parts(u, cdf, pdf);
  return u * cdf;
}
__device__ __forceinline__ float dgelu_f(float u) {
  float cdf, pdf;
  gelu_parts(u, cdf, pdf);
  return fmaf(u, pdf, cdf);
}

// ---- column reductions over rows of [M, C] tensors (8 channels per thread) ------------------------------------------
// Block = 512 threads arranged as CW column threads x (512 / CW) row lanes, CW = power of two in [4, 128] >= min(C/8, 128),
// so that narrow tensors (C = 32: 4 vectors, C = 96: 12 vectors) still use every thread.  colred_combine sums the row lanes through shared
// memory and issues one atomicAdd per (block, channel) and quantity.
struct ColRedShape {
  int cw_log2;  // log2(CW)
  int colb;     // column blocks
  __host__ static ColRedShape make(int C8) {
    ColRedShape s;
    s.cw_log2 = 2;  // narrow tensors (C = 32: 4 vectors): 4 column threads x 128 row lanes, every thread loads
    while ((1 << s.cw_log2) < C8 && s.cw_log2 < 7) ++s.cw_log2;
    s.colb = (C8 + (1 << s.cw_log2) - 1) >> s.cw_log2;
    return s;
  }
};

template <int NQ>
__device__ __forceinline__ void colred_combine(const float (&acc)[NQ][8], float* red /* [NQ*8][512] */, int cw_log2,
                                               int c8_base, int C8, float* const (&out)[NQ]) {
  const int tid = threadIdx.x;
  const int CW = 1 << cw_log2, RL = 512 >> cw_log2;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int k = 0; k < 8; ++k) red[(q * 8 + k) * 512 + tid] = acc[q][k];
  __syncthreads();
  for (int o = tid; o < NQ * 8 * CW; o += 512) {
    const int cx = o & (CW - 1);
    const int qk = o >> cw_log2;
    const int c8 = c8_base + cx;
    if (c8 < C8) {
      float s = 0.f;
      for (int r = 0; r < RL; ++r) s += red[qk * 512 + r * CW + cx];
      atomicAdd(out[qk >> 3] + (long long)c8 * 8 + (qk & 7), s);
    }
  }
}

}  // namespace vb
