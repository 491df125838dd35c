// GroupNorm (+ timestep scale / shift) + SiLU | ReLU on channels-last 16-bit rows [N, R, C]: the norm / activation half
// of the reference's conv block  Block.forward: act(norm(proj(x)) * (scale + 1) + shift)  (VM/unet/blocks.py:88-113, the
// UNet3DBase defaults norm="group", activation="silu": VM/unet/unet3d_base.py:58-72; CELLDiff family).
//
// GroupNorm statistics are per (sample, group); everything the elementwise passes need folds into per-(sample, channel)
// fp32 coefficients computed on the host side from the column sums (tiny [N, C] tensors):
//   forward :  v = a[n,c] * x + b[n,c],  y = act(v)            a = rstd_g gamma (1 + scale),  b = (beta - mean_g rstd_g gamma)(1 + scale) + shift
//   backward:  dv = dy * act'(v);  S1[n,c] = sum_r dv,  S2[n,c] = sum_r dv * x    (one reduction pass over dy and x)
//              dx = c1[n,c] * dv + c2[n,c] * x + c3[n,c]                            (one elementwise pass)
// HBM-bound: forward reads x once and writes y once; backward reads dy and x twice and writes dx once.
#include "common.cuh"

namespace vb {

enum { GN_ACT_NONE = 0, GN_ACT_RELU = 1, GN_ACT_SILU = 2, GN_ACT_LEAKY = 3, GN_ACT_ELU = 4, GN_ACT_SELU = 5 };
constexpr float SELU_ALPHA = 1.6732632423543772f, SELU_SCALE = 1.0507009873554805f;  // torch.nn.SELU

template <int ACT>
__device__ __forceinline__ float gn_act(float v) {
  if (ACT == GN_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == GN_ACT_SILU) return v * __frcp_rn(1.0f + __expf(-v));
  if (ACT == GN_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;  // nn.LeakyReLU() default slope
  if (ACT == GN_ACT_ELU) return v > 0.f ? v : expm1f(v);     // nn.ELU() alpha = 1
  if (ACT == GN_ACT_SELU) return SELU_SCALE * (v > 0.f ? v : SELU_ALPHA * expm1f(v));
  return v;
}
template <int ACT>
__device__ __forceinline__ float gn_dact(float v) {
  if (ACT == GN_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (ACT == GN_ACT_SILU) {
    const float s = __frcp_rn(1.0f + __expf(-v));
    return s * fmaf(v, 1.0f - s, 1.0f);
  }
  if (ACT == GN_ACT_LEAKY) return v > 0.f ? 1.f : 0.01f;
  if (ACT == GN_ACT_ELU) return v > 0.f ? 1.f : __expf(v);
  if (ACT == GN_ACT_SELU) return SELU_SCALE * (v > 0.f ? 1.f : SELU_ALPHA * __expf(v));
  return 1.f;
}

// y = act(a[n,c] * x + b[n,c]); thread = 8 channels of one row
template <bool BF16, int ACT>
__global__ void __launch_bounds__(256)
affine_nc_act_kernel(const uint4* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                     uint4* __restrict__ y, long long R, int C8, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = (int)(i % C8);
  const long long n = (i / C8) / R;
  const float4* ap = reinterpret_cast<const float4*>(a + (n * C8 + c8) * 8);
  const float4* bp = reinterpret_cast<const float4*>(b + (n * C8 + c8) * 8);
  const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
  const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const uint4 q = __ldg(x + i);
  const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = H16<BF16>::unpack(w4[k]);
    o[k] = H16<BF16>::pack(gn_act<ACT>(fmaf(av[2 * k], f.x, bv[2 * k])), gn_act<ACT>(fmaf(av[2 * k + 1], f.y, bv[2 * k + 1])));
  }
  y[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// S1[n,c] += sum_r dv, S2[n,c] += sum_r dv * x, dv = dy * act'(a x + b)   (block shape: ColRedShape; grid z = sample)
template <bool BF16, int ACT>
__global__ void __launch_bounds__(512)
gn_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const float* __restrict__ a,
                     const float* __restrict__ b, float* __restrict__ s1, float* __restrict__ s2, int R, int C8,
                     int rows_per_block, int cw_log2) {
  __shared__ float red[16 * 512];
  const int cx = threadIdx.x & ((1 << cw_log2) - 1), ry = threadIdx.x >> cw_log2, RL = 512 >> cw_log2;
  const int c8 = (blockIdx.x << cw_log2) + cx;
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  float acc[2][8], av[8], bv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    acc[0][k] = acc[1][k] = 0.f;
    av[k] = c8 < C8 ? a[((long long)n * C8 + c8) * 8 + k] : 0.f;
    bv[k] = c8 < C8 ? b[((long long)n * C8 + c8) * 8 + k] : 0.f;
  }
  if (c8 < C8) {
    const long long base = (long long)n * R * C8 + c8;
#pragma unroll 4
    for (int r = r0 + ry; r < r1; r += RL) {
      const uint4 qd = __ldg(dy + base + (long long)r * C8), qx = __ldg(x + base + (long long)r * C8);
      const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 d = H16<BF16>::unpack(wd[k]), xv = H16<BF16>::unpack(wx[k]);
        const float d0 = d.x * gn_dact<ACT>(fmaf(av[2 * k], xv.x, bv[2 * k]));
        const float d1 = d.y * gn_dact<ACT>(fmaf(av[2 * k + 1], xv.y, bv[2 * k + 1]));
        acc[0][2 * k] += d0;
        acc[0][2 * k + 1] += d1;
        acc[1][2 * k] = fmaf(d0, xv.x, acc[1][2 * k]);
        acc[1][2 * k + 1] = fmaf(d1, xv.y, acc[1][2 * k + 1]);
      }
    }
  }
  float* const outs[2] = {s1 + (long long)n * C8 * 8, s2 + (long long)n * C8 * 8};
  colred_combine<2>(acc, red, cw_log2, blockIdx.x << cw_log2, C8, outs);
}

// dx = c1[n,c] * dv + c2[n,c] * x + c3[n,c], dv = dy * act'(a x + b)
template <bool BF16, int ACT>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const float* __restrict__ coef,
                    uint4* __restrict__ dx, long long R, int C8, long long total8, long long nc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = (int)(i % C8);
  const long long n = (i / C8) / R;
  const long long o0 = (n * C8 + c8) * 8;
  float cf[5][8];  // a, b, c1, c2, c3: five [N, C] planes of `coef`
#pragma unroll
  for (int p = 0; p < 5; ++p) {
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(coef + p * nc + o0));
    const float4 u1 = __ldg(reinterpret_cast<const float4*>(coef + p * nc + o0) + 1);
    cf[p][0] = u0.x; cf[p][1] = u0.y; cf[p][2] = u0.z; cf[p][3] = u0.w;
    cf[p][4] = u1.x; cf[p][5] = u1.y; cf[p][6] = u1.z; cf[p][7] = u1.w;
  }
  const uint4 qd = __ldg(dy + i), qx = __ldg(x + i);
  const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 d = H16<BF16>::unpack(wd[k]), xv = H16<BF16>::unpack(wx[k]);
    const int k0 = 2 * k, k1 = 2 * k + 1;
    const float d0 = d.x * gn_dact<ACT>(fmaf(cf[0][k0], xv.x, cf[1][k0]));
    const float d1 = d.y * gn_dact<ACT>(fmaf(cf[0][k1], xv.y, cf[1][k1]));
    o[k] = H16<BF16>::pack(fmaf(cf[2][k0], d0, fmaf(cf[3][k0], xv.x, cf[4][k0])),
                           fmaf(cf[2][k1], d1, fmaf(cf[3][k1], xv.y, cf[4][k1])));
  }
  dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

static unsigned gn_blocks(long long total) { return (unsigned)((total + 255) / 256); }

}  // namespace vb

using namespace vb;

#define GN_ACTS(...)                                                    \
  switch (act) {                                                        \
    case 0: { constexpr int ACT = 0; __VA_ARGS__; } break;              \
    case 1: { constexpr int ACT = 1; __VA_ARGS__; } break;              \
    case 2: { constexpr int ACT = 2; __VA_ARGS__; } break;              \
    case 3: { constexpr int ACT = 3; __VA_ARGS__; } break;              \
    case 4: { constexpr int ACT = 4; __VA_ARGS__; } break;              \
    default: { constexpr int ACT = 5; __VA_ARGS__; } break;             \
  }
#define GN_DISPATCH(dtype, act, ...)                                                                  \
  do {                                                                                                \
    if ((act) < 0 || (act) > 5) return vb::fail(VB200_ERR_INVALID, "activation %d", (int)(act));      \
    if ((dtype) == VB200_BF16) {                                                                      \
      constexpr bool BF = true;                                                                       \
      GN_ACTS(__VA_ARGS__)                                                                            \
    } else if ((dtype) == VB200_FP16) {                                                               \
      constexpr bool BF = false;                                                                      \
      GN_ACTS(__VA_ARGS__)                                                                            \
    } else {                                                                                          \
      return vb::fail(VB200_ERR_UNSUPPORTED, "dtype %d", (int)(dtype));                               \
    }                                                                                                 \
  } while (0)

extern "C" int vb200_affine_nc_act(const void* x, const float* a, const float* b, void* y, int64_t N, int64_t R, int C,
                                   int act, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && a && b && y, "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total8 = N * R * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  GN_DISPATCH(dtype, act, (affine_nc_act_kernel<BF, ACT><<<gn_blocks(total8), 256, 0, st>>>((const uint4*)x, a, b, (uint4*)y, R, C / 8, total8)));
  return check_launch("vb200_affine_nc_act");
}

extern "C" int vb200_gn_bwd_reduce(const void* dy, const void* x, const float* a, const float* b, float* s1, float* s2,
                                   int64_t N, int64_t R, int C, int act, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(dy && x && a && b && s1 && s2, "null pointer");
  VB_SUPPORTED(C % 8 == 0 && R < (1LL << 31) && N < 65536, "C (%d) %% 8", C);
  const int C8 = C / 8;
  const ColRedShape sh = ColRedShape::make(C8);
  long long rpb = (R * sh.colb * N + 148 * 2 - 1) / (148 * 2);
  const long long min_rows = 4LL * (512 >> sh.cw_log2);
  if (rpb < min_rows) rpb = min_rows;
  if (rpb > R) rpb = R;
  dim3 grid(sh.colb, (unsigned)((R + rpb - 1) / rpb), (unsigned)N);
  cudaStream_t st = (cudaStream_t)stream;
  GN_DISPATCH(dtype, act, (gn_bwd_reduce_kernel<BF, ACT><<<grid, 512, 0, st>>>((const uint4*)dy, (const uint4*)x, a, b, s1, s2, (int)R, C8, (int)rpb, sh.cw_log2)));
  return check_launch("vb200_gn_bwd_reduce");
}

extern "C" int vb200_gn_bwd_apply(const void* dy, const void* x, const float* coef, void* dx, int64_t N, int64_t R, int C,
                                  int act, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(dy && x && coef && dx, "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total8 = N * R * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  GN_DISPATCH(dtype, act, (gn_bwd_apply_kernel<BF, ACT><<<gn_blocks(total8), 256, 0, st>>>((const uint4*)dy, (const uint4*)x, coef, (uint4*)dx, R, C / 8, total8, N * C)));
  return check_launch("vb200_gn_bwd_apply");
}
