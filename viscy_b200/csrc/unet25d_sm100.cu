// HBM-bound kernels of the 2.5-D U-Net (Unet25d / ConvBlock3D, VM/unet/unet25d.py:206-251,
// VM/components/conv_block_3d.py:261-298) on channels-last rows [N, D, H, W, C], 8 channels (16 B) per thread:
//   scale_relu      : nn.Dropout3d (one keep/scale factor per sample and channel) fused with nn.ReLU, forward / backward
//   avgpool_hw2     : nn.AvgPool3d((1,2,2), stride (1,2,2)), forward / backward
//   upsample2x_hw   : nn.Upsample(scale_factor=(1,2,2), mode="trilinear", align_corners=False), forward / backward
// The convolutions of the blocks run on the tcgen05 implicit GEMM (gemm_sm100.cu, vb200_conv3d_igemm*).
#include "common.cuh"

namespace vb {

static inline unsigned nblocks25(long long total, int bs = 256) { return (unsigned)((total + bs - 1) / bs); }

template <bool BF16>
__device__ __forceinline__ void unpack8f(const uint4& q, float* v) {
  const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = H16<BF16>::unpack(w4[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
template <bool BF16>
__device__ __forceinline__ uint4 pack8f(const float* v) {
  return make_uint4(H16<BF16>::pack(v[0], v[1]), H16<BF16>::pack(v[2], v[3]), H16<BF16>::pack(v[4], v[5]),
                    H16<BF16>::pack(v[6], v[7]));
}

// forward : y = relu?(x * scale[n, c])                       (gate == nullptr)
// backward: y = x * scale[n, c] * (gate > 0 if relu)         (x = dy, gate = forward output)
template <bool BF16>
__global__ void __launch_bounds__(256)
scale_relu_kernel(const uint4* __restrict__ x, const uint4* __restrict__ gate, const float* __restrict__ scale,
                  uint4* __restrict__ y, long long rows_per_sample, int C8, int relu, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = (int)(i % C8);
  const long long n = (i / C8) / rows_per_sample;
  float v[8], g[8];
  unpack8f<BF16>(__ldg(x + i), v);
  if (gate != nullptr) unpack8f<BF16>(__ldg(gate + i), g);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (scale != nullptr) v[k] *= __ldg(scale + (n * C8 + c8) * 8 + k);
    if (relu) {
      if (gate != nullptr) {
        if (!(g[k] > 0.f)) v[k] = 0.f;
      } else {
        v[k] = fmaxf(v[k], 0.f);
      }
    }
  }
  y[i] = pack8f<BF16>(v);
}

// x [P, H, W, C] -> y [P, H/2, W/2, C]: mean of the 2x2 window (floor semantics: a trailing odd row / column is dropped)
template <bool BF16>
__global__ void __launch_bounds__(256)
avgpool_hw2_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int C8, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int OH = H >> 1, OW = W >> 1;
  const int c8 = (int)(i % C8);
  long long r = i / C8;
  const int ow = (int)(r % OW);
  r /= OW;
  const int oh = (int)(r % OH);
  const long long p = r / OH;
  const uint4* src = x + ((p * H + 2 * oh) * W + 2 * ow) * C8 + c8;
  float a[8], b[8], acc[8];
  unpack8f<BF16>(__ldg(src), a);
  unpack8f<BF16>(__ldg(src + C8), b);
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = a[k] + b[k];
  unpack8f<BF16>(__ldg(src + (long long)W * C8), a);
  unpack8f<BF16>(__ldg(src + (long long)W * C8 + C8), b);
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.25f * (acc[k] + a[k] + b[k]);
  y[i] = pack8f<BF16>(acc);
}

// dx [P, H, W, C] = 0.25 * dy [P, H/2, W/2, C] broadcast over the window (zero for a dropped trailing row / column)
template <bool BF16>
__global__ void __launch_bounds__(256)
avgpool_hw2_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int H, int W, int C8, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int OH = H >> 1, OW = W >> 1;
  const int c8 = (int)(i % C8);
  long long r = i / C8;
  const int w = (int)(r % W);
  r /= W;
  const int h = (int)(r % H);
  const long long p = r / H;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 0.f;
  if ((h >> 1) < OH && (w >> 1) < OW) {
    unpack8f<BF16>(__ldg(dy + ((p * OH + (h >> 1)) * OW + (w >> 1)) * C8 + c8), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= 0.25f;
  }
  dx[i] = pack8f<BF16>(v);
}

// bilinear x2 in H and W, align_corners=False: source coordinate of output o is max(o/2 - 1/4, 0):
//   o = 2i   -> 0.25 x[i-1] + 0.75 x[i]   (x[0] alone for i = 0)
//   o = 2i+1 -> 0.75 x[i]   + 0.25 x[i+1] (x[L-1] alone for i = L-1)
__device__ __forceinline__ void up2_taps(int o, int L, int& i0, int& i1, float& w0, float& w1) {
  const int i = o >> 1;
  if (o & 1) {
    i0 = i;
    i1 = min(i + 1, L - 1);
    w0 = 0.75f;
    w1 = 0.25f;
  } else {
    i0 = max(i - 1, 0);
    i1 = i;
    w0 = i > 0 ? 0.25f : 0.f;
    w1 = i > 0 ? 0.75f : 1.f;
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256)
upsample2x_hw_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int C8, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int OH = 2 * H, OW = 2 * W;
  const int c8 = (int)(i % C8);
  long long r = i / C8;
  const int ow = (int)(r % OW);
  r /= OW;
  const int oh = (int)(r % OH);
  const long long p = r / OH;
  int h0, h1, w0, w1;
  float a0, a1, b0, b1;
  up2_taps(oh, H, h0, h1, a0, a1);
  up2_taps(ow, W, w0, w1, b0, b1);
  const uint4* base = x + p * H * W * C8 + c8;
  float v00[8], v01[8], v10[8], v11[8], o[8];
  unpack8f<BF16>(__ldg(base + ((long long)h0 * W + w0) * C8), v00);
  unpack8f<BF16>(__ldg(base + ((long long)h0 * W + w1) * C8), v01);
  unpack8f<BF16>(__ldg(base + ((long long)h1 * W + w0) * C8), v10);
  unpack8f<BF16>(__ldg(base + ((long long)h1 * W + w1) * C8), v11);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = a0 * (b0 * v00[k] + b1 * v01[k]) + a1 * (b0 * v10[k] + b1 * v11[k]);
  y[i] = pack8f<BF16>(o);
}

// adjoint of the above, gather form: input i receives from outputs 2i-1 (0.25), 2i (0.75 | 1), 2i+1 (0.75 | 1), 2i+2 (0.25)
__device__ __forceinline__ int up2_adj(int i, int L, int* o, float* w) {
  int n = 0;
  if (i >= 1) { o[n] = 2 * i - 1; w[n++] = 0.25f; }
  o[n] = 2 * i;     w[n++] = i > 0 ? 0.75f : 1.f;
  o[n] = 2 * i + 1; w[n++] = i < L - 1 ? 0.75f : 1.f;
  if (i + 1 <= L - 1) { o[n] = 2 * i + 2; w[n++] = 0.25f; }
  return n;
}

template <bool BF16>
__global__ void __launch_bounds__(256)
upsample2x_hw_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int H, int W, int C8, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int OH = 2 * H, OW = 2 * W;
  const int c8 = (int)(i % C8);
  long long r = i / C8;
  const int w = (int)(r % W);
  r /= W;
  const int h = (int)(r % H);
  const long long p = r / H;
  int oh[4], ow[4];
  float ah[4], aw[4];
  const int nh = up2_adj(h, H, oh, ah), nw = up2_adj(w, W, ow, aw);
  const uint4* base = dy + p * OH * OW * C8 + c8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int a = 0; a < nh; ++a)
    for (int b = 0; b < nw; ++b) {
      float v[8];
      unpack8f<BF16>(__ldg(base + ((long long)oh[a] * OW + ow[b]) * C8), v);
      const float wt = ah[a] * aw[b];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(wt, v[k], acc[k]);
    }
  dx[i] = pack8f<BF16>(acc);
}

}  // namespace vb

using namespace vb;

#define DT_SWITCH25(dtype, ...)                                                 \
  do {                                                                          \
    if ((dtype) == VB200_BF16) { constexpr bool BF = true; __VA_ARGS__; }       \
    else if ((dtype) == VB200_FP16) { constexpr bool BF = false; __VA_ARGS__; } \
    else return vb::fail(VB200_ERR_UNSUPPORTED, "dtype %d", (int)(dtype));      \
  } while (0)

extern "C" int vb200_scale_relu(const void* x, const void* gate, const float* scale, void* y, int64_t N,
                                int64_t rows_per_sample, int C, int relu, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && y && N > 0 && rows_per_sample > 0, "bad arguments");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total8 = N * rows_per_sample * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DT_SWITCH25(dtype, scale_relu_kernel<BF><<<nblocks25(total8), 256, 0, st>>>((const uint4*)x, (const uint4*)gate, scale, (uint4*)y, rows_per_sample, C / 8, relu, total8));
  return check_launch("vb200_scale_relu");
}

extern "C" int vb200_avgpool_hw2(const void* src, void* dst, int64_t P, int H, int W, int C, int backward, int dtype,
                                 vb200_stream_t stream) {
  VB_REQUIRE(src && dst && P > 0 && H >= 2 && W >= 2, "bad arguments");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const int C8 = C / 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (!backward) {
    const long long total8 = P * (H / 2) * (W / 2) * C8;
    DT_SWITCH25(dtype, avgpool_hw2_fwd_kernel<BF><<<nblocks25(total8), 256, 0, st>>>((const uint4*)src, (uint4*)dst, H, W, C8, total8));
  } else {
    const long long total8 = P * H * W * C8;
    DT_SWITCH25(dtype, avgpool_hw2_bwd_kernel<BF><<<nblocks25(total8), 256, 0, st>>>((const uint4*)src, (uint4*)dst, H, W, C8, total8));
  }
  return check_launch("vb200_avgpool_hw2");
}

extern "C" int vb200_upsample2x_hw(const void* src, void* dst, int64_t P, int H, int W, int C, int backward, int dtype,
                                   vb200_stream_t stream) {
  VB_REQUIRE(src && dst && P > 0 && H >= 1 && W >= 1, "bad arguments");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const int C8 = C / 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (!backward) {
    const long long total8 = P * 4 * H * W * C8;
    DT_SWITCH25(dtype, upsample2x_hw_fwd_kernel<BF><<<nblocks25(total8), 256, 0, st>>>((const uint4*)src, (uint4*)dst, H, W, C8, total8));
  } else {
    const long long total8 = P * H * W * C8;
    DT_SWITCH25(dtype, upsample2x_hw_bwd_kernel<BF><<<nblocks25(total8), 256, 0, st>>>((const uint4*)src, (uint4*)dst, H, W, C8, total8));
  }
  return check_launch("vb200_upsample2x_hw");
}
