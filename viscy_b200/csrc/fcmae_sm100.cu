// FCMAE pieces (VM/unet/fcmae.py, VM/components/heads.py:657-695) that the ConvNeXt kernels do not already cover:
//   * rows_select: the sparse (masked) path's row gather / scatter / masking on channels-last rows
//     (masked_patchify / masked_unpatchify / `x *= unmasked`, fcmae.py:91-141, 215-227);
//   * shuffle_pool: PixelToVoxelShuffleHead = pixel shuffle by r + ConstantPad2d((r-1,0,r-1,0)) + AvgPool2d(r, stride 1),
//     NHWC decoder rows -> NCHW (= NCDHW after the head's reshape) and its adjoint.
// HBM-bound, 16-byte vectors along the channel dimension, NCHW side coalesced along X.
#include <algorithm>

#include "common.cuh"

namespace vb {

// dst[r, :] = (map[r] >= 0 ? src[map[r], :] : 0) + (base ? base[r, :] : 0);  C8 = channels / 8 (16-byte vectors)
template <bool BF16>
__global__ void __launch_bounds__(256)
rows_select_kernel(const uint4* __restrict__ src, const int* __restrict__ map, const uint4* __restrict__ base,
                   uint4* __restrict__ dst, long long n_vec, int C8) {
  using H = H16<BF16>;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n_vec; v += (long long)gridDim.x * blockDim.x) {
    const long long r = v / C8;
    const int c8 = (int)(v - r * C8);
    const int m = __ldg(map + r);
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (m >= 0) q = __ldg(src + (long long)m * C8 + c8);
    if (base) {
      const uint4 b = __ldg(base + v);
      const uint32_t qa[4] = {q.x, q.y, q.z, q.w}, ba[4] = {b.x, b.y, b.z, b.w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = H::unpack(qa[k]), y = H::unpack(ba[k]);
        o[k] = H::pack(x.x + y.x, x.y + y.y);
      }
      q = make_uint4(o[0], o[1], o[2], o[3]);
    }
    dst[v] = q;
  }
}

constexpr int SP_W = 16;  // decoder pixels per block along w

// out[j] = sum_{a<R, b<R} P[a][j + b] (pool) or P[off][j + off] (no pool), j < R, from a window whose first element is 4-byte
// aligned and whose row pitch is even: R rows x 2R columns read as 32-bit pairs
template <typename H, int R>
__device__ __forceinline__ void window_outputs(const uint16_t* base, int pitch, int pool, int off, float (&o)[R]) {
  if (!pool) {
#pragma unroll
    for (int j = 0; j < R; ++j) o[j] = H::to_f(*reinterpret_cast<const typename H::T*>(base + off * pitch + j + off));
    return;
  }
  if constexpr (R == 1) {
    o[0] = H::to_f(*reinterpret_cast<const typename H::T*>(base));
  } else {
    float col[2 * R];
#pragma unroll
    for (int b = 0; b < 2 * R; ++b) col[b] = 0.f;
#pragma unroll
    for (int a = 0; a < R; ++a) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(base + a * pitch);
#pragma unroll
      for (int b2 = 0; b2 < R; ++b2) {
        const float2 f = H::unpack(row[b2]);
        col[2 * b2] += f.x;
        col[2 * b2 + 1] += f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      float v = 0.f;
#pragma unroll
      for (int b = 0; b < R; ++b) v += col[j + b];
      o[j] = v;
    }
  }
}

// forward: block = (w segment, hh, n).  Shared memory holds the SHUFFLED image region the block's outputs look at,
// P[Cq][2R-1][SP_W*R + R-1] (rows hh*R-(R-1) .. hh*R+R-1, columns w0*R-(R-1) .. (w0+SP_W)*R-1; zeros outside the image = the
// front padding), filled from 16-byte channel vectors of decoder rows hh-1, hh; every output is then an R x R window sum
// at compile-time strides.
template <bool BF16, int R>
__global__ void __launch_bounds__(256)
shuffle_pool_fwd_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int h, int w, int Cq, int pool) {
  using H = H16<BF16>;
  extern __shared__ uint16_t sm[];
  constexpr int PR = 2 * R - 1, PW = SP_W * R + R + (R & 1), XW = SP_W * R, RR = R * R;  // PW even: rows stay 4-byte aligned
  const int C = Cq * RR, C8 = C / 8;
  const int w0 = blockIdx.x * SP_W, hh = blockIdx.y;
  const long long n = blockIdx.z;
  for (int idx = threadIdx.x; idx < Cq * PR; idx += blockDim.x)  // pad columns behind the XW + R - 1 real ones
    for (int c = XW + R - 1; c < PW; ++c) sm[idx * PW + c] = 0;
  for (int idx = threadIdx.x; idx < 2 * (SP_W + 1) * C8; idx += blockDim.x) {
    const int c8 = idx % C8;
    int t = idx / C8;
    const int xl = t % (SP_W + 1), rr = t / (SP_W + 1);
    const int y = hh - 1 + rr, x = w0 - 1 + xl;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && x >= 0 && x < w) q = __ldg(reinterpret_cast<const uint4*>(src + ((n * h + y) * w + x) * (long long)C) + c8);
    const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c8 * 8 + k;
      const int cq = c / RR, ij = c % RR, i = ij / R, j = ij % R;
      const int pr = rr * R + i - 1, pc = xl * R + j - 1;
      if (pr >= 0 && pc >= 0) sm[(cq * PR + pr) * PW + pc] = static_cast<uint16_t>(k & 1 ? qw[k >> 1] >> 16 : qw[k >> 1] & 0xffffu);
    }
  }
  __syncthreads();
  const int Ws = w * R, Hs = h * R;
  constexpr float inv = 1.0f / (float)RR;
  // thread = (cq, i, group of R consecutive x): column sums of the R x 2R window, then R sliding outputs
  for (int idx = threadIdx.x; idx < Cq * R * SP_W; idx += blockDim.x) {
    const int g = idx % SP_W;
    const int t = idx / SP_W;
    const int i = t % R, cq = t / R;
    const int X = (w0 + g) * R;
    if (w0 + g >= w) continue;
    float o[R];
    window_outputs<H, R>(sm + (cq * PR + i) * PW + g * R, PW, pool, R - 1, o);
    uint16_t* dp = dst + ((n * Cq + cq) * Hs + hh * R + i) * (long long)Ws + X;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const typename H::T v = H::from_f(pool ? o[j] * inv : o[j]);
      dp[j] = *reinterpret_cast<const uint16_t*>(&v);
    }
  }
}

// backward: d_src[n,hh,ww, cq R^2 + i R + j] = pool ? R^-2 sum_{a,b<R} dP[cq, hh R + i + a, ww R + j + b] : dP[cq, hh R + i, ww R + j]
// smem D[Cq][2R-1][XW + R - 1] (16-bit), loaded coalesced along X.
template <bool BF16, int R>
__global__ void __launch_bounds__(256)
shuffle_pool_bwd_kernel(const uint16_t* __restrict__ dP, uint16_t* __restrict__ dsrc, int h, int w, int Cq, int pool) {
  using H = H16<BF16>;
  extern __shared__ uint16_t sm[];
  constexpr int XW = SP_W * R, RW = XW + R + (R & 1), RH = 2 * R - 1, RR = R * R;  // RW even (zero pad columns)
  const int w0 = blockIdx.x * SP_W, hh = blockIdx.y;
  const long long n = blockIdx.z;
  const int Ws = w * R, Hs = h * R;
  for (int idx = threadIdx.x; idx < Cq * RH * RW; idx += blockDim.x) {
    const int xl = idx % RW;
    const int t = idx / RW;
    const int yl = t % RH, cq = t / RH;
    const int Y = hh * R + yl, X = w0 * R + xl;
    sm[idx] = (Y < Hs && X < Ws) ? __ldg(dP + ((n * Cq + cq) * Hs + Y) * (long long)Ws + X) : (uint16_t)0;
  }
  __syncthreads();
  const int C = Cq * RR;
  constexpr float inv = 1.0f / (float)RR;
  // thread = (decoder pixel wl, cq, i): the R channels j = 0..R-1 are R sliding sums over the same R x (2R-1) window
  for (int idx = threadIdx.x; idx < SP_W * Cq * R; idx += blockDim.x) {
    const int ci = idx % (Cq * R), wl = idx / (Cq * R);
    const int ww = w0 + wl;
    if (ww >= w) continue;
    const int cq = ci / R, i = ci % R;
    float o[R];
    window_outputs<H, R>(sm + (cq * RH + i) * RW + wl * R, RW, pool, 0, o);
    uint16_t* dp = dsrc + ((n * h + hh) * w + ww) * (long long)C + cq * RR + i * R;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const typename H::T v = H::from_f(pool ? o[j] * inv : o[j]);
      dp[j] = *reinterpret_cast<const uint16_t*>(&v);
    }
  }
}

template <bool BF16, int R>
static int shuffle_pool_launch(const uint16_t* s, uint16_t* d, int B, int h, int w, int Cq, int pool, int backward,
                               cudaStream_t st) {
  const size_t smem = (size_t)Cq * (2 * R - 1) * (SP_W * R + R + (R & 1)) * 2;  // same extent both ways
  if (smem > 200 * 1024) return fail(VB200_ERR_UNSUPPORTED, "tile of %zu bytes does not fit shared memory", smem);
  static PerDeviceOnce once[2];
  const int dev = PerDeviceOnce::device();
  if (once[backward ? 1 : 0].need(dev)) {
    const void* fn = backward ? (const void*)shuffle_pool_bwd_kernel<BF16, R> : (const void*)shuffle_pool_fwd_kernel<BF16, R>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    once[backward ? 1 : 0].done(dev);
  }
  dim3 grid((w + SP_W - 1) / SP_W, h, B);
  if (!backward) shuffle_pool_fwd_kernel<BF16, R><<<grid, 256, smem, st>>>(s, d, h, w, Cq, pool);
  else shuffle_pool_bwd_kernel<BF16, R><<<grid, 256, smem, st>>>(s, d, h, w, Cq, pool);
  return check_launch("shuffle_pool");
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_rows_select(const void* src, const int32_t* map, const void* base, void* dst, int64_t n_dst, int C,
                                 int dtype, vb200_stream_t stream) {
  VB_REQUIRE(src && map && dst, "null pointer");
  VB_REQUIRE(C > 0 && C % 8 == 0, "channels (%d) must be a multiple of 8", C);
  VB_SUPPORTED(dtype == VB200_BF16 || dtype == VB200_FP16, "dtype %d", dtype);
  if (n_dst <= 0) return VB200_OK;
  const int C8 = C / 8;
  const long long n_vec = (long long)n_dst * C8;
  const unsigned blocks = (unsigned)std::min<long long>((n_vec + 255) / 256, 148LL * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB200_BF16)
    rows_select_kernel<true><<<blocks, 256, 0, st>>>((const uint4*)src, map, (const uint4*)base, (uint4*)dst, n_vec, C8);
  else
    rows_select_kernel<false><<<blocks, 256, 0, st>>>((const uint4*)src, map, (const uint4*)base, (uint4*)dst, n_vec, C8);
  return check_launch("rows_select");
}

extern "C" int vb200_shuffle_pool(const void* src, void* dst, int B, int h, int w, int Cq, int r, int pool, int backward,
                                  int dtype, vb200_stream_t stream) {
  VB_REQUIRE(src && dst, "null pointer");
  VB_REQUIRE(r >= 1 && Cq > 0, "scale %d / channels %d", r, Cq);
  VB_SUPPORTED(r == 1 || r == 2 || r == 4 || r == 8, "pixel-shuffle scale %d (1, 2, 4, 8 are built)", r);
  VB_SUPPORTED(dtype == VB200_BF16 || dtype == VB200_FP16, "dtype %d", dtype);
  VB_SUPPORTED((Cq * r * r) % 8 == 0, "decoder channels (%d) must be a multiple of 8", Cq * r * r);
  cudaStream_t st = (cudaStream_t)stream;
  const uint16_t* s = (const uint16_t*)src;
  uint16_t* d = (uint16_t*)dst;
#define SP_CASE(RV)                                                                                            \
  case RV:                                                                                                     \
    return dtype == VB200_BF16 ? shuffle_pool_launch<true, RV>(s, d, B, h, w, Cq, pool, backward, st)           \
                               : shuffle_pool_launch<false, RV>(s, d, B, h, w, Cq, pool, backward, st);
  switch (r) {
    SP_CASE(1) SP_CASE(2) SP_CASE(4) SP_CASE(8)
  }
#undef SP_CASE
  return fail(VB200_ERR_UNSUPPORTED, "scale %d", r);
}
