// MixedLoss / ms_ssim_25d (VU/losses/mixed_loss.py:42-69, VU/evaluation/metrics.py:174-349) as fused box-filter kernels.
//
// One pyramid level of the reference = five depthwise uniform-window conv3d (kernel (D, kh, kw), bf16 inputs / bf16 kernel /
// fp32 accumulation / bf16 outputs) + the SSIM / contrast-sensitivity maps + their per-sample means + avg_pool3d((1,2,2)) of both
// volumes + target.max() of the next level.  Here: ONE pass over the two volumes per level.  A block owns a 32 x 64 pixel
// tile of one (sample, channel) plane stack: it sums the five quantities over depth for the haloed tile into shared memory
// (rounded to bf16 element by element exactly where the reference rounds), slides the kw- and kh-wide window sums in place,
// forms mu = bf16(bf16(1/N) * sum) and the maps, and reduces.  The same pass emits the L1 / L2 sums of MixedLoss, the pooled
// volumes of the next level and its data range.  The backward pass is the adjoint box filter over the three gradient
// maps (d/d mu_x, d/d mu_xx, d/d mu_xy) applied to every depth slice, fused with the L1 / L2 and pooling gradients.
// HBM-bound: algorithmic traffic per level = both volumes read once (+ pooled volumes written).
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace vb {
namespace ssim {

constexpr int NT = 256, KMAX = 16;
// tile of owned pixels per block: 32 x 64 on large planes (halo overhead 1.5x), 16 x 32 when that would leave SMs idle

__device__ __forceinline__ float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// compile-time typed element access (DT: VB200_BF16 | VB200_FP16 | VB200_FP32) so that the unrolled depth loops carry no
// dtype branches
template <int DT>
__device__ __forceinline__ float ldt(const void* p, long long i) {
  if constexpr (DT == 2) return __ldg(reinterpret_cast<const float*>(p) + i);
  else if constexpr (DT == 0) return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(p) + i));
  else return __half2float(__ldg(reinterpret_cast<const __half*>(p) + i));
}
template <int DT>
__device__ __forceinline__ float rnd(float v) {
  if constexpr (DT == 2) return v;
  else if constexpr (DT == 0) return __bfloat162float(__float2bfloat16_rn(v));
  else return __half2float(__float2half_rn(v));
}
template <int DT>
__device__ __forceinline__ void stt(void* p, long long i, float v) {
  if constexpr (DT == 2) reinterpret_cast<float*>(p)[i] = v;
  else if constexpr (DT == 0) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}
constexpr int DU = 8;  // depth slices in flight per thread: independent loads issued together (the loops are latency-bound)

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// in-place "valid" sliding window sums over a [rows][cols] fp32 plane with row pitch `pitch`: along columns (horizontal)
// one thread per row, along rows (vertical) one thread per column; out[i] = sum_{j<k} in[i+j], i < n - k + 1
__device__ __forceinline__ void slide(float* line, int stride, int n, int k) {
  float s = 0.f;
  for (int j = 0; j < k; ++j) s += line[j * stride];
  for (int i = 0; i + k <= n; ++i) {
    const float gone = line[i * stride];  // still the input: position i is written below, positions > i later
    line[i * stride] = s;
    if (i + k < n) s += line[(i + k) * stride] - gone;
  }
}

struct Params {
  const void* x;
  const void* y;
  void* xp;
  void* yp;
  float* mu;            // [5][BC][Ho][Wo] or null
  const float* dmax;    // data range of this level (device scalar)
  float* dmax_next;     // max of the pooled target (pre-set to -inf) or null
  float* acc;           // [B][4] sums: ssim, cs, |x-y|, (x-y)^2
  int xdt, ydt, C, D, H, W, kh, kw, Ho, Wo, flags;  // flags: 1 ssim, 2 l1/l2 sums, 4 pooled volumes
  float wbf, k1, k2;
};

template <int XDT, int YDT, int TH, int TW>
__global__ void __launch_bounds__(NT, 2) level_fwd_kernel(const Params p) {
  extern __shared__ float sm[];  // [5][RH][RW]
  __shared__ float red[4][NT / 32];
  const int RH = TH + p.kh - 1, RW = TW + p.kw - 1, RP = RH * RW;
  const int h0 = blockIdx.y * TH, w0 = blockIdx.x * TW;
  const int bc = blockIdx.z, b = bc / p.C;
  const long long plane = (long long)p.H * p.W;
  const long long base = (long long)bc * p.D * plane;
  float l1 = 0.f, l2 = 0.f, s_ssim = 0.f, s_cs = 0.f;

  if (p.flags & 1) {
    for (int idx = threadIdx.x; idx < RP; idx += NT) {
      const int r = idx / RW, c = idx - r * RW;
      const int h = h0 + r, w = w0 + c;
      float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
      if (h < p.H && w < p.W) {
        const long long o = base + (long long)h * p.W + w;
        for (int d0 = 0; d0 < p.D; d0 += DU) {
          float xv[DU], yv[DU];
#pragma unroll
          for (int u = 0; u < DU; ++u) {
            const bool on = d0 + u < p.D;
            xv[u] = on ? ldt<XDT>(p.x, o + (d0 + u) * plane) : 0.f;
            yv[u] = on ? ldt<YDT>(p.y, o + (d0 + u) * plane) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < DU; ++u) {  // zeros past the last slice add nothing
            sx += bf(xv[u]);
            sy += bf(yv[u]);
            sxx += bf(xv[u] * xv[u]);
            syy += bf(yv[u] * yv[u]);
            sxy += bf(xv[u] * yv[u]);
          }
        }
      }
      sm[idx] = sx;
      sm[RP + idx] = sy;
      sm[2 * RP + idx] = sxx;
      sm[3 * RP + idx] = syy;
      sm[4 * RP + idx] = sxy;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 5 * RH; t += NT) slide(sm + (t / RH) * RP + (t % RH) * RW, 1, RW, p.kw);
    __syncthreads();
    for (int t = threadIdx.x; t < 5 * TW; t += NT) slide(sm + (t / TW) * RP + (t % TW), RW, RH, p.kh);
    __syncthreads();
    const float dr = __ldg(p.dmax);
    const float c1 = (p.k1 * dr) * (p.k1 * dr), c2 = (p.k2 * dr) * (p.k2 * dr);
    for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
      const int r = idx / TW, c = idx - r * TW;
      const int ho = h0 + r, wo = w0 + c;
      if (ho >= p.Ho || wo >= p.Wo) continue;
      const int q = r * RW + c;
      const float mx = bf(p.wbf * sm[q]), my = bf(p.wbf * sm[RP + q]);
      const float mxx = bf(p.wbf * sm[2 * RP + q]), myy = bf(p.wbf * sm[3 * RP + q]), mxy = bf(p.wbf * sm[4 * RP + q]);
      const float vx = mxx - mx * mx, vy = myy - my * my, vxy = mxy - mx * my;
      const float cs = (2.f * vxy + c2) / (vx + vy + c2);
      const float ss = ((2.f * mx * my + c1) / (mx * mx + my * my + c1)) * cs;
      s_ssim += ss;
      s_cs += cs;
      if (p.mu) {
        const long long mo = ((long long)bc * p.Ho + ho) * p.Wo + wo, ms = (long long)gridDim.z * p.Ho * p.Wo;
        p.mu[mo] = mx;
        p.mu[ms + mo] = my;
        p.mu[2 * ms + mo] = mxx;
        p.mu[3 * ms + mo] = myy;
        p.mu[4 * ms + mo] = mxy;
      }
    }
  }
  if (p.flags & 2) {
    for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
      const int r = idx / TW, c = idx - r * TW;
      const int h = h0 + r, w = w0 + c;
      if (h >= p.H || w >= p.W) continue;
      const long long o = base + (long long)h * p.W + w;
      for (int d0 = 0; d0 < p.D; d0 += DU) {
        float df[DU];
#pragma unroll
        for (int u = 0; u < DU; ++u)
          df[u] = d0 + u < p.D ? ldt<XDT>(p.x, o + (d0 + u) * plane) - ldt<YDT>(p.y, o + (d0 + u) * plane) : 0.f;
#pragma unroll
        for (int u = 0; u < DU; ++u) {
          l1 += fabsf(df[u]);
          l2 += df[u] * df[u];
        }
      }
    }
  }
  float ymax = -INFINITY;
  if (p.flags & 4) {
    const int H2 = p.H / 2, W2 = p.W / 2;
    const long long plane2 = (long long)H2 * W2, base2 = (long long)bc * p.D * plane2;
    for (int idx = threadIdx.x; idx < (TH / 2) * (TW / 2); idx += NT) {
      const int r = idx / (TW / 2), c = idx - r * (TW / 2);
      const int h2 = h0 / 2 + r, w2 = w0 / 2 + c;
      if (h2 >= H2 || w2 >= W2) continue;
      const long long o = base + (long long)(2 * h2) * p.W + 2 * w2;
      constexpr int PU = 4;
      for (int d0 = 0; d0 < p.D; d0 += PU) {
        float xs[PU], ys[PU];
#pragma unroll
        for (int u = 0; u < PU; ++u) {
          const long long od = o + (d0 + u) * plane;
          const bool on = d0 + u < p.D;
          xs[u] = on ? ldt<XDT>(p.x, od) + ldt<XDT>(p.x, od + 1) + ldt<XDT>(p.x, od + p.W) + ldt<XDT>(p.x, od + p.W + 1) : 0.f;
          ys[u] = on ? ldt<YDT>(p.y, od) + ldt<YDT>(p.y, od + 1) + ldt<YDT>(p.y, od + p.W) + ldt<YDT>(p.y, od + p.W + 1) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < PU; ++u) {
          if (d0 + u < p.D) {
            const long long o2 = base2 + (d0 + u) * plane2 + (long long)h2 * W2 + w2;
            stt<XDT>(p.xp, o2, 0.25f * xs[u]);
            stt<YDT>(p.yp, o2, 0.25f * ys[u]);
            ymax = fmaxf(ymax, rnd<YDT>(0.25f * ys[u]));
          }
        }
      }
    }
  }
  // block reduction: 4 sums + 1 max
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float v[4] = {s_ssim, s_cs, l1, l2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][wid] = v[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  if (lane == 0 && p.dmax_next && ymax > -INFINITY) atomic_max_float(p.dmax_next, ymax);
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int k = 0; k < NT / 32; ++k) s += red[threadIdx.x][k];
    const bool on = threadIdx.x < 2 ? (p.flags & 1) : (p.flags & 2);
    if (on) atomicAdd(p.acc + b * 4 + threadIdx.x, s);
  }
}

struct BwdParams {
  const void* x;
  const void* y;
  void* dx;                // [BC][D][H][W] in x's dtype
  const float* mu;         // [5][BC][Ho][Wo]
  const float* dmax;
  const float* g_ssim;     // [B] upstream gradients of the per-sample MEANS (or null)
  const float* g_cs;       // [B]
  const float* g_l1;       // scalar upstream of the l1 / l2 MEANS (or null)
  const float* g_l2;
  const void* g_pool;      // [BC][D][H/2][W/2] gradient of the pooled volume in x's dtype (or null)
  int xdt, ydt, B, C, D, H, W, kh, kw, Ho, Wo;
  float wbf, k1, k2;
};

template <int XDT, int YDT, int TH, int TW>
__global__ void __launch_bounds__(NT, 3) level_bwd_kernel(const BwdParams p) {
  extern __shared__ float sm[];  // [3][RH][RW]: G_mux, G_muxx, G_muxy at output pixels (h0 - kh + 1 + r, w0 - kw + 1 + c)
  const int RH = TH + p.kh - 1, RW = TW + p.kw - 1, RP = RH * RW;
  const int h0 = blockIdx.y * TH, w0 = blockIdx.x * TW;
  const int bc = blockIdx.z, b = bc / p.C;
  const long long plane = (long long)p.H * p.W;
  const long long base = (long long)bc * p.D * plane;
  const bool do_ssim = p.mu != nullptr && (p.g_ssim != nullptr || p.g_cs != nullptr);
  if (do_ssim) {
    const float inv_n = 1.0f / ((float)p.C * p.Ho * p.Wo);
    const float gs = p.g_ssim ? __ldg(p.g_ssim + b) * inv_n : 0.f, gc = p.g_cs ? __ldg(p.g_cs + b) * inv_n : 0.f;
    const float dr = __ldg(p.dmax);
    const float c1 = (p.k1 * dr) * (p.k1 * dr), c2 = (p.k2 * dr) * (p.k2 * dr);
    const long long ms = (long long)gridDim.z * p.Ho * p.Wo;
    // branch-free, four pixels per iteration: the 20 window-mean loads of an iteration are issued together
    for (int i0 = threadIdx.x; i0 < RP; i0 += 4 * NT) {
      float m[4][5];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i0 + u * NT;
        const int r = idx / RW, c = idx - r * RW;
        const int ho = h0 - p.kh + 1 + r, wo = w0 - p.kw + 1 + c;
        ok[u] = idx < RP && ho >= 0 && wo >= 0 && ho < p.Ho && wo < p.Wo;
        const long long mo = ok[u] ? ((long long)bc * p.Ho + ho) * p.Wo + wo : 0;
#pragma unroll
        for (int q = 0; q < 5; ++q) m[u][q] = __ldg(p.mu + q * ms + mo);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i0 + u * NT;
        if (idx >= RP) continue;
        float gx = 0.f, gxx = 0.f, gxy = 0.f;
        if (ok[u]) {
          const float mx = m[u][0], my = m[u][1], mxx = m[u][2], myy = m[u][3], mxy = m[u][4];
          const float vx = mxx - mx * mx, vy = myy - my * my, vxy = mxy - mx * my;
          const float dcs = 1.0f / (vx + vy + c2), cs = (2.f * vxy + c2) * dcs;
          const float dl = 1.0f / (mx * mx + my * my + c1), l = (2.f * mx * my + c1) * dl;
          const float g_cs_tot = gc + gs * l, g_l_tot = gs * cs;
          gxy = g_cs_tot * 2.f * dcs;
          gxx = -g_cs_tot * cs * dcs;
          gx = g_l_tot * (2.f * my - 2.f * l * mx) * dl - 2.f * mx * gxx - my * gxy;
        }
        sm[idx] = gx;
        sm[RP + idx] = gxx;
        sm[2 * RP + idx] = gxy;
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * RH; t += NT) slide(sm + (t / RH) * RP + (t % RH) * RW, 1, RW, p.kw);
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * TW; t += NT) slide(sm + (t / TW) * RP + (t % TW), RW, RH, p.kh);
    __syncthreads();
  }
  const float n_tot = (float)p.B * p.C * p.D * (float)plane;
  const float gl1 = p.g_l1 ? __ldg(p.g_l1) / n_tot : 0.f, gl2 = p.g_l2 ? 2.f * __ldg(p.g_l2) / n_tot : 0.f;
  const bool l12 = p.g_l1 != nullptr || p.g_l2 != nullptr;
  const int H2 = p.H / 2, W2 = p.W / 2;
  const long long plane2 = (long long)H2 * W2, base2 = (long long)bc * p.D * plane2;
  for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
    const int r = idx / TW, c = idx - r * TW;
    const int h = h0 + r, w = w0 + c;
    if (h >= p.H || w >= p.W) continue;
    float A = 0.f, Bq = 0.f, Cq = 0.f;
    if (do_ssim) {
      const int q = r * RW + c;
      A = p.wbf * sm[q];
      Bq = 2.f * p.wbf * sm[RP + q];
      Cq = p.wbf * sm[2 * RP + q];
    }
    const bool pooled = p.g_pool != nullptr && (h >> 1) < H2 && (w >> 1) < W2;
    const long long o = base + (long long)h * p.W + w;
    const long long o2 = base2 + (long long)(h >> 1) * W2 + (w >> 1);
    for (int d0 = 0; d0 < p.D; d0 += DU) {
      float xv[DU], yv[DU], gp[DU];
#pragma unroll
      for (int u = 0; u < DU; ++u) {
        const bool on = d0 + u < p.D;
        xv[u] = (on && (do_ssim || l12)) ? ldt<XDT>(p.x, o + (d0 + u) * plane) : 0.f;
        yv[u] = (on && (do_ssim || l12)) ? ldt<YDT>(p.y, o + (d0 + u) * plane) : 0.f;
        gp[u] = (on && pooled) ? ldt<XDT>(p.g_pool, o2 + (d0 + u) * plane2) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < DU; ++u) {
        if (d0 + u < p.D) {
          float g = A + xv[u] * Bq + yv[u] * Cq;
          if (l12) {
            const float df = xv[u] - yv[u];
            g += gl1 * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) + gl2 * df;
          }
          g += 0.25f * gp[u];
          stt<XDT>(p.dx, o + (d0 + u) * plane, g);
        }
      }
    }
  }
}

// max over a flat array (any dtype) into a float pre-set to -inf
__global__ void __launch_bounds__(256) max_kernel(const void* x, int dt, long long n, float* out) {
  float m = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, ld_any(x, i, dt));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomic_max_float(out, m);
}

static int dt_ok(int dt) { return dt == VB200_BF16 || dt == VB200_FP16 || dt == 2; }

}  // namespace ssim
}  // namespace vb

using namespace vb;
using namespace vb::ssim;

static float bf16_round_host(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&v, &u, 4);
  return v;
}

extern "C" int vb200_max_f(const void* x, int dtype, int64_t n, float* out, vb200_stream_t stream) {
  VB_REQUIRE(x && out, "null pointer");
  VB_SUPPORTED(dt_ok(dtype), "dtype %d", dtype);
  if (n <= 0) return VB200_OK;
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 8);
  max_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, dtype, n, out);
  return check_launch("max_f");
}

static int ssim_smem_opt_in(const void* fn, PerDeviceOnce& once) {
  const int dev = PerDeviceOnce::device();
  if (once.need(dev)) {
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    once.done(dev);
  }
  return 0;
}

extern "C" int vb200_ssim25d_level_fwd(const void* x, const void* y, int x_dtype, int y_dtype, int B, int C, int D, int H,
                                       int W, int kh, int kw, const float* data_range, float* acc, float* mu, void* x_pool,
                                       void* y_pool, float* data_range_next, int flags, vb200_stream_t stream) {
  VB_REQUIRE(x && y && acc, "null pointer");
  VB_SUPPORTED(dt_ok(x_dtype) && dt_ok(y_dtype), "dtypes %d / %d", x_dtype, y_dtype);
  VB_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "empty volume");
  VB_REQUIRE(!(flags & 1) || (kh >= 1 && kw >= 1 && kh <= KMAX && kw <= KMAX && H >= kh && W >= kw && data_range),
             "window %dx%d on a %dx%d plane", kh, kw, H, W);
  VB_REQUIRE(!(flags & 4) || (x_pool && y_pool), "pooled outputs missing");
  Params p{};
  p.x = x; p.y = y; p.xp = x_pool; p.yp = y_pool; p.mu = mu; p.dmax = data_range; p.dmax_next = data_range_next; p.acc = acc;
  p.xdt = x_dtype; p.ydt = y_dtype; p.C = C; p.D = D; p.H = H; p.W = W; p.kh = kh; p.kw = kw;
  p.Ho = H - kh + 1; p.Wo = W - kw + 1; p.flags = flags;
  p.wbf = bf16_round_host(1.0f / (float)(D * kh * kw));
  p.k1 = 0.01f; p.k2 = 0.03f;
  const bool big = (long long)((W + 63) / 64) * ((H + 31) / 32) * B * C >= 2 * 148;
  const int TH = big ? 32 : 16, TW = big ? 64 : 32;
  const size_t smem = (flags & 1) ? (size_t)5 * (TH + kh - 1) * (TW + kw - 1) * sizeof(float) : 0;
  dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B * C);
  static PerDeviceOnce once[18];
#define SSIM_FWD(XD, YD)                                                                               \
  if (x_dtype == XD && y_dtype == YD) {                                                                \
    if (big) {                                                                                         \
      ssim_smem_opt_in((const void*)level_fwd_kernel<XD, YD, 32, 64>, once[XD * 3 + YD]);              \
      level_fwd_kernel<XD, YD, 32, 64><<<grid, NT, smem, (cudaStream_t)stream>>>(p);                   \
    } else {                                                                                           \
      ssim_smem_opt_in((const void*)level_fwd_kernel<XD, YD, 16, 32>, once[9 + XD * 3 + YD]);          \
      level_fwd_kernel<XD, YD, 16, 32><<<grid, NT, smem, (cudaStream_t)stream>>>(p);                   \
    }                                                                                                  \
  }
  SSIM_FWD(0, 0) SSIM_FWD(0, 1) SSIM_FWD(0, 2) SSIM_FWD(1, 0) SSIM_FWD(1, 1) SSIM_FWD(1, 2) SSIM_FWD(2, 0) SSIM_FWD(2, 1)
  SSIM_FWD(2, 2)
#undef SSIM_FWD
  return check_launch("ssim25d_level_fwd");
}

extern "C" int vb200_ssim25d_level_bwd(const void* x, const void* y, int x_dtype, int y_dtype, int B, int C, int D, int H,
                                       int W, int kh, int kw, const float* data_range, const float* mu, const float* g_ssim,
                                       const float* g_cs, const float* g_l1, const float* g_l2, const void* g_pool, void* dx,
                                       vb200_stream_t stream) {
  VB_REQUIRE(x && y && dx, "null pointer");
  VB_SUPPORTED(dt_ok(x_dtype) && dt_ok(y_dtype), "dtypes %d / %d", x_dtype, y_dtype);
  const bool do_ssim = mu != nullptr && (g_ssim != nullptr || g_cs != nullptr);
  VB_REQUIRE(!do_ssim || (kh >= 1 && kw >= 1 && kh <= KMAX && kw <= KMAX && H >= kh && W >= kw && data_range),
             "window %dx%d on a %dx%d plane", kh, kw, H, W);
  BwdParams p{};
  p.x = x; p.y = y; p.dx = dx; p.mu = mu; p.dmax = data_range; p.g_ssim = g_ssim; p.g_cs = g_cs; p.g_l1 = g_l1; p.g_l2 = g_l2;
  p.g_pool = g_pool; p.xdt = x_dtype; p.ydt = y_dtype; p.B = B; p.C = C; p.D = D; p.H = H; p.W = W; p.kh = kh; p.kw = kw;
  p.Ho = H - kh + 1; p.Wo = W - kw + 1;
  p.wbf = bf16_round_host(1.0f / (float)(D * kh * kw));
  p.k1 = 0.01f; p.k2 = 0.03f;
  const bool big = (long long)((W + 63) / 64) * ((H + 31) / 32) * B * C >= 2 * 148;
  const int TH = big ? 32 : 16, TW = big ? 64 : 32;
  const size_t smem = do_ssim ? (size_t)3 * (TH + kh - 1) * (TW + kw - 1) * sizeof(float) : 0;
  dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B * C);
  static PerDeviceOnce once[18];
#define SSIM_BWD(XD, YD)                                                                               \
  if (x_dtype == XD && y_dtype == YD) {                                                                \
    if (big) {                                                                                         \
      ssim_smem_opt_in((const void*)level_bwd_kernel<XD, YD, 32, 64>, once[XD * 3 + YD]);              \
      level_bwd_kernel<XD, YD, 32, 64><<<grid, NT, smem, (cudaStream_t)stream>>>(p);                   \
    } else {                                                                                           \
      ssim_smem_opt_in((const void*)level_bwd_kernel<XD, YD, 16, 32>, once[9 + XD * 3 + YD]);          \
      level_bwd_kernel<XD, YD, 16, 32><<<grid, NT, smem, (cudaStream_t)stream>>>(p);                   \
    }                                                                                                  \
  }
  SSIM_BWD(0, 0) SSIM_BWD(0, 1) SSIM_BWD(0, 2) SSIM_BWD(1, 0) SSIM_BWD(1, 1) SSIM_BWD(1, 2) SSIM_BWD(2, 0) SSIM_BWD(2, 1)
  SSIM_BWD(2, 2)
#undef SSIM_BWD
  return check_launch("ssim25d_level_bwd");
}
