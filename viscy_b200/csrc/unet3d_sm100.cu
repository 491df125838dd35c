// HBM-bound kernels of the 3-D U-Net family (UNet3DBase / Unet3d, VM/unet/unet3d_base.py:145-198, VM/unet/blocks.py:88-113):
// NCDHW <-> channels-last conversion at the model boundary, BatchNorm3d (+ReLU) apply / backward on channels-last rows,
// channel concat / split, bias and residual adds.  The 3x3x3 convolutions themselves run on the tcgen05 GEMM
// (layout_sm100.cu im2col3d/col2im3d lowering; conv3d_sm100.cu implicit GEMM for small channel counts).
#include "common.cuh"

namespace vb {

static inline unsigned nblocks(long long total, int bs = 256) { return (unsigned)((total + bs - 1) / bs); }

// x (N,C,S) any float type -> y (N,S,Cpad) 16-bit, channels >= C zero.  S = D*H*W.  Thread = (n, s): reads C strided
// values (coalesced across s), writes Cpad contiguous.
template <typename TIN, bool BF16>
__global__ void __launch_bounds__(256)
to_cl_kernel(const TIN* __restrict__ x, uint16_t* __restrict__ y, int C, int Cpad, long long S, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / S, s = idx - n * S;
  const TIN* src = x + n * C * S + s;
  uint16_t* dst = y + idx * Cpad;
  for (int c = 0; c < Cpad; ++c) {
    float v = c < C ? static_cast<float>(src[(long long)c * S]) : 0.f;
    typename H16<BF16>::T hv = H16<BF16>::from_f(v);
    dst[c] = *reinterpret_cast<uint16_t*>(&hv);
  }
}

// y (N,S,Cpad) 16-bit -> x (N,C,S) 16-bit
__global__ void __launch_bounds__(256)
from_cl_kernel(const uint16_t* __restrict__ y, uint16_t* __restrict__ x, int C, int Cpad, long long S, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / S, s = idx - n * S;
  for (int c = 0; c < C; ++c) x[(n * C + c) * S + s] = y[idx * Cpad + c];
}

// y = act(x * scale[c] + shift[c]); 8 channels per thread
template <bool BF16>
__global__ void __launch_bounds__(256)
affine_act_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                  uint4* __restrict__ y, int C8, int relu, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = (int)(i % C8);
  const uint4 q = __ldg(x + i);
  const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = H16<BF16>::unpack(w4[k]);
    float a = fmaf(f.x, __ldg(scale + c8 * 8 + 2 * k), __ldg(shift + c8 * 8 + 2 * k));
    float b = fmaf(f.y, __ldg(scale + c8 * 8 + 2 * k + 1), __ldg(shift + c8 * 8 + 2 * k + 1));
    if (relu) {
      a = fmaxf(a, 0.f);
      b = fmaxf(b, 0.f);
    }
    o[k] = H16<BF16>::pack(a, b);
  }
  y[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// BatchNorm backward reductions over rows: s1[c] += sum dy', s2[c] += sum dy' * xhat, dy' = dy * (y > 0 if relu)
// (block shape: see ColRedShape in common.cuh)
template <bool BF16>
__global__ void __launch_bounds__(512)
bn_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ y,
                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ s1,
                     float* __restrict__ s2, long long M, int C8, int relu, int rows_per_block, int cw_log2) {
  __shared__ float red[16 * 512];
  const int cx = threadIdx.x & ((1 << cw_log2) - 1), ry = threadIdx.x >> cw_log2, RL = 512 >> cw_log2;
  const int c8 = (blockIdx.x << cw_log2) + cx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float a[2][8], mu[8], rs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[0][k] = a[1][k] = 0.f;
    mu[k] = c8 < C8 ? mean[c8 * 8 + k] : 0.f;
    rs[k] = c8 < C8 ? rstd[c8 * 8 + k] : 0.f;
  }
  if (c8 < C8) {
    for (long long r = r0 + ry; r < r1; r += RL) {
      const uint4 qd = __ldg(dy + r * C8 + c8), qx = __ldg(x + r * C8 + c8);
      uint4 qy = make_uint4(0, 0, 0, 0);
      if (relu) qy = __ldg(y + r * C8 + c8);
      const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w}, wy[4] = {qy.x, qy.y, qy.z, qy.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 d = H16<BF16>::unpack(wd[k]);
        const float2 xv = H16<BF16>::unpack(wx[k]);
        if (relu) {
          const float2 yv = H16<BF16>::unpack(wy[k]);
          if (!(yv.x > 0.f)) d.x = 0.f;
          if (!(yv.y > 0.f)) d.y = 0.f;
        }
        a[0][2 * k] += d.x;
        a[0][2 * k + 1] += d.y;
        a[1][2 * k] = fmaf(d.x, (xv.x - mu[2 * k]) * rs[2 * k], a[1][2 * k]);
        a[1][2 * k + 1] = fmaf(d.y, (xv.y - mu[2 * k + 1]) * rs[2 * k + 1], a[1][2 * k + 1]);
      }
    }
  }
  float* const outs[2] = {s1, s2};
  colred_combine<2>(a, red, cw_log2, blockIdx.x << cw_log2, C8, outs);
}

// dx = g[c] * (dy' - m1[c] - xhat * m2[c]),  g = gamma * rstd, m1 = s1/M, m2 = s2/M (training) or 0 (eval)
template <bool BF16>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ y,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ g,
                    const float* __restrict__ m1, const float* __restrict__ m2, uint4* __restrict__ dx, int C8,
                    int relu, long long total8, int raw, float inv_m) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c0 = (int)(i % C8) * 8;
  const uint4 qd = __ldg(dy + i), qx = __ldg(x + i);
  uint4 qy = make_uint4(0, 0, 0, 0);
  if (relu) qy = __ldg(y + i);
  const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w}, wy[4] = {qy.x, qy.y, qy.z, qy.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 d = H16<BF16>::unpack(wd[k]);
    const float2 xv = H16<BF16>::unpack(wx[k]);
    if (relu) {
      const float2 yv = H16<BF16>::unpack(wy[k]);
      if (!(yv.x > 0.f)) d.x = 0.f;
      if (!(yv.y > 0.f)) d.y = 0.f;
    }
    const int c = c0 + 2 * k;
    const float xa = (xv.x - __ldg(mean + c)) * __ldg(rstd + c), xb = (xv.y - __ldg(mean + c + 1)) * __ldg(rstd + c + 1);
    // raw: g = gamma, m1 / m2 = the column sums of the reduce pass (scaled here instead of by three tiny torch launches)
    const float ga = raw ? __ldg(g + c) * __ldg(rstd + c) : __ldg(g + c);
    const float gb = raw ? __ldg(g + c + 1) * __ldg(rstd + c + 1) : __ldg(g + c + 1);
    const float sc = raw ? inv_m : 1.f;
    o[k] = H16<BF16>::pack(ga * (d.x - __ldg(m1 + c) * sc - xa * (__ldg(m2 + c) * sc)),
                           gb * (d.y - __ldg(m1 + c + 1) * sc - xb * (__ldg(m2 + c + 1) * sc)));
  }
  dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// out[m, 0:Ca] = a[m], out[m, Ca:Ca+Cb] = b[m]   (inverse: split out into a and b); 8 channels per thread
__global__ void __launch_bounds__(256)
cat2_kernel(uint4* __restrict__ a, uint4* __restrict__ b, uint4* __restrict__ out, int Ca8, int Cb8, long long total8,
            int inverse) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int Co8 = Ca8 + Cb8;
  const int c = (int)(i % Co8);
  const long long m = i / Co8;
  uint4* p = c < Ca8 ? a + m * Ca8 + c : b + m * Cb8 + (c - Ca8);
  if (!inverse)
    out[i] = *p;
  else
    *p = out[i];
}

// y = x (+ other) (+ bias[c]); 8 channels per thread
template <bool BF16>
__global__ void __launch_bounds__(256)
add_rows_kernel(const uint4* __restrict__ x, const uint4* __restrict__ other, const float* __restrict__ bias,
                uint4* __restrict__ y, int C8, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c0 = (int)(i % C8) * 8;
  const uint4 q = __ldg(x + i);
  uint4 r = make_uint4(0, 0, 0, 0);
  if (other != nullptr) r = __ldg(other + i);
  const uint32_t wq[4] = {q.x, q.y, q.z, q.w}, wr[4] = {r.x, r.y, r.z, r.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 f = H16<BF16>::unpack(wq[k]);
    if (other != nullptr) {
      const float2 g = H16<BF16>::unpack(wr[k]);
      f.x += g.x;
      f.y += g.y;
    }
    if (bias != nullptr) {
      f.x += __ldg(bias + c0 + 2 * k);
      f.y += __ldg(bias + c0 + 2 * k + 1);
    }
    o[k] = H16<BF16>::pack(f.x, f.y);
  }
  y[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace vb

using namespace vb;

#define DT_SWITCH(dtype, ...)                                                   \
  do {                                                                          \
    if ((dtype) == VB200_BF16) { constexpr bool BF = true; __VA_ARGS__; }       \
    else if ((dtype) == VB200_FP16) { constexpr bool BF = false; __VA_ARGS__; } \
    else return vb::fail(VB200_ERR_UNSUPPORTED, "dtype %d", (int)(dtype));      \
  } while (0)

extern "C" int vb200_to_channels_last(const void* x, int x_dtype, void* y, int64_t N, int C, int Cpad, int64_t S,
                                      int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && y && Cpad >= C, "bad arguments");
  const long long total = N * S;
  cudaStream_t st = (cudaStream_t)stream;
#define TOCL(TIN, BF) to_cl_kernel<TIN, BF><<<nblocks(total), 256, 0, st>>>((const TIN*)x, (uint16_t*)y, C, Cpad, S, total)
  if (dtype == VB200_BF16) {
    if (x_dtype == 2) TOCL(float, true);
    else if (x_dtype == 0) TOCL(__nv_bfloat16, true);
    else return fail(VB200_ERR_UNSUPPORTED, "input dtype %d for bf16 path", x_dtype);
  } else if (dtype == VB200_FP16) {
    if (x_dtype == 2) TOCL(float, false);
    else if (x_dtype == 1) TOCL(__half, false);
    else return fail(VB200_ERR_UNSUPPORTED, "input dtype %d for fp16 path", x_dtype);
  } else {
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  }
#undef TOCL
  return check_launch("vb200_to_channels_last");
}

extern "C" int vb200_from_channels_last(const void* y, void* x, int64_t N, int C, int Cpad, int64_t S,
                                        vb200_stream_t stream) {
  VB_REQUIRE(x && y && Cpad >= C, "bad arguments");
  const long long total = N * S;
  from_cl_kernel<<<nblocks(total), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)y, (uint16_t*)x, C, Cpad, S, total);
  return check_launch("vb200_from_channels_last");
}

extern "C" int vb200_affine_act(const void* x, const float* scale, const float* shift, void* y, int64_t M, int C,
                                int relu, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && scale && shift && y, "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total8 = M * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DT_SWITCH(dtype, affine_act_kernel<BF><<<nblocks(total8), 256, 0, st>>>((const uint4*)x, scale, shift, (uint4*)y, C / 8, relu, total8));
  return check_launch("vb200_affine_act");
}

extern "C" int vb200_bn_bwd_reduce(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                                   float* s1, float* s2, int64_t M, int C, int relu, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(dy && x && mean && rstd && s1 && s2 && (y || !relu), "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const int C8 = C / 8;
  const ColRedShape sh = ColRedShape::make(C8);
  long long rpb = (M * sh.colb + 148 * 2 - 1) / (148 * 2);
  const long long min_rows = 4LL * (512 >> sh.cw_log2);
  if (rpb < min_rows) rpb = min_rows;
  if (rpb > M) rpb = M;
  dim3 grid(sh.colb, (unsigned)((M + rpb - 1) / rpb));
  cudaStream_t st = (cudaStream_t)stream;
  DT_SWITCH(dtype, bn_bwd_reduce_kernel<BF><<<grid, 512, 0, st>>>((const uint4*)dy, (const uint4*)x, (const uint4*)y, mean, rstd, s1, s2, M, C8, relu, (int)rpb, sh.cw_log2));
  return check_launch("vb200_bn_bwd_reduce");
}

extern "C" int vb200_bn_bwd_apply(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                                  const float* g, const float* m1, const float* m2, void* dx, int64_t M, int C,
                                  int relu, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(dy && x && mean && rstd && g && m1 && m2 && dx && (y || !relu), "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total8 = M * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DT_SWITCH(dtype, bn_bwd_apply_kernel<BF><<<nblocks(total8), 256, 0, st>>>((const uint4*)dy, (const uint4*)x, (const uint4*)y, mean, rstd, g, m1, m2, (uint4*)dx, C / 8, relu, total8, 0, 1.f));
  return check_launch("vb200_bn_bwd_apply");
}

extern "C" int vb200_bn_bwd_apply_raw(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                                      const float* gamma, const float* s1, const float* s2, float inv_m, void* dx, int64_t M,
                                      int C, int relu, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(dy && x && mean && rstd && gamma && s1 && s2 && dx && (y || !relu), "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total8 = M * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DT_SWITCH(dtype, bn_bwd_apply_kernel<BF><<<nblocks(total8), 256, 0, st>>>((const uint4*)dy, (const uint4*)x, (const uint4*)y, mean, rstd, gamma, s1, s2, (uint4*)dx, C / 8, relu, total8, 1, inv_m));
  return check_launch("vb200_bn_bwd_apply_raw");
}

// BatchNorm statistics -> everything the apply pass and the backward need, plus the running-statistics update, in one
// launch (was ~14 torch launches on [C]-sized vectors per layer and step).  sums = [2][Cc] column sums of (x - pivot) and
// (x - pivot)^2 (training) or null (eval: running statistics).  weight / bias / run_* have Cn entries, outputs Cc
// (channel-padded rows: padded channels get scale = shift = 0).  momentum < 0: no running-statistics update.
namespace vb {
__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ pivot,
                                   const float* __restrict__ weight, const float* __restrict__ bias, float* run_mean,
                                   float* run_var, int Cn, int Cc, float inv_m, float unbias, float eps, float momentum,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cc) return;
  const bool real = c < Cn;
  float mean, var;
  if (sums != nullptr) {
    const float a = sums[c] * inv_m, q = sums[Cc + c] * inv_m;
    var = fmaxf(q - a * a, 0.f);
    mean = pivot != nullptr ? a + pivot[c] : a;
    if (real && momentum >= 0.f && run_mean != nullptr) {
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * (var * unbias);
    }
  } else {
    mean = real ? run_mean[c] : 0.f;
    var = real ? run_var[c] : 1.f;
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = real ? weight[c] * rstd : 0.f;
  scale[c] = sc;
  shift[c] = real ? bias[c] - mean * sc : 0.f;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
}
}  // namespace vb

extern "C" int vb200_bn_finalize(const float* sums, const float* pivot, const float* weight, const float* bias,
                                 float* run_mean, float* run_var, int Cn, int Cc, double M, float eps, float momentum,
                                 float* scale, float* shift, float* mean, float* rstd, vb200_stream_t stream) {
  VB_REQUIRE(weight && bias && scale && shift && mean && rstd, "null pointer");
  VB_REQUIRE(sums != nullptr || (run_mean != nullptr && run_var != nullptr), "eval mode needs running statistics");
  VB_REQUIRE(Cn > 0 && Cc >= Cn && M >= 1.0, "Cn %d Cc %d M %g", Cn, Cc, M);
  const float unbias = (float)(M / (M > 1.0 ? M - 1.0 : 1.0));
  vb::bn_finalize_kernel<<<(Cc + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, pivot, weight, bias, run_mean, run_var, Cn, Cc,
                                                                         (float)(1.0 / M), unbias, eps, momentum, scale, shift,
                                                                         mean, rstd);
  return check_launch("vb200_bn_finalize");
}

extern "C" int vb200_cat2(void* a, void* b, void* out, int64_t M, int Ca, int Cb, int inverse, vb200_stream_t stream) {
  VB_REQUIRE(a && b && out, "null pointer");
  VB_SUPPORTED(Ca % 8 == 0 && Cb % 8 == 0, "concat needs channel counts %% 8 == 0 (%d, %d)", Ca, Cb);
  const long long total8 = M * ((Ca + Cb) / 8);
  cat2_kernel<<<nblocks(total8), 256, 0, (cudaStream_t)stream>>>((uint4*)a, (uint4*)b, (uint4*)out, Ca / 8, Cb / 8, total8, inverse);
  return check_launch("vb200_cat2");
}

extern "C" int vb200_add_rows(const void* x, const void* other, const float* bias, void* y, int64_t M, int C, int dtype,
                              vb200_stream_t stream) {
  VB_REQUIRE(x && y, "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total8 = M * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DT_SWITCH(dtype, add_rows_kernel<BF><<<nblocks(total8), 256, 0, st>>>((const uint4*)x, (const uint4*)other, bias, (uint4*)y, C / 8, total8));
  return check_launch("vb200_add_rows");
}
