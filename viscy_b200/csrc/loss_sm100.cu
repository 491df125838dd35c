// Mean squared error of the training step (the reference's default loss: CY/engine.py:197 `nn.MSELoss()`, applied to the
// 16-bit autocast prediction and the fp32 target): one pass for the sum, one for the gradient, instead of torch's
// cast-to-fp32 copy + elementwise + reduction (forward) and elementwise + cast back (backward) over the 22 M-voxel output.
// HBM-bound: forward reads pred + target, backward reads both and writes dpred in pred's dtype.
#include <algorithm>

#include "common.cuh"

namespace vb {

__device__ __forceinline__ float loss_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 8 consecutive elements of a bf16 / fp16 / fp32 array as floats (p 16-byte aligned at element i for the 16-bit types,
// 32-byte for fp32: the callers check the bases and keep i % 8 == 0)
template <int DT>
__device__ __forceinline__ void ld8(const void* p, long long i, float* f) {
  if constexpr (DT == 2) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + i));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 v = H16<DT == 0>::unpack(w[k]);
      f[2 * k] = v.x;
      f[2 * k + 1] = v.y;
    }
  }
}

// *sum += scale * sum_i (pred_i - target_i)^2
template <int PDT, int TDT>
__global__ void __launch_bounds__(256)
mse_sum_kernel(const void* __restrict__ pred, const void* __restrict__ target, long long n, float scale, float* __restrict__ sum) {
  __shared__ float part[8];
  float acc = 0.f;
  const long long n8 = n >> 3;
  for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < n8; j += (long long)gridDim.x * 256) {
    float a[8], b[8];
    ld8<PDT>(pred, j * 8, a);
    ld8<TDT>(target, j * 8, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = a[k] - b[k];
      acc = fmaf(d, d, acc);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 7)) {  // tail
    const long long i = (n8 << 3) + threadIdx.x;
    const float d = ld_any(pred, i, PDT) - ld_any(target, i, TDT);
    acc = fmaf(d, d, acc);
  }
  acc = loss_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w];
    atomicAdd(sum, s * scale);
  }
}

// dpred_i = (pred_i - target_i) * (*gout) * scale, in pred's dtype
template <int PDT, int TDT>
__global__ void __launch_bounds__(256)
mse_bwd_kernel(const void* __restrict__ pred, const void* __restrict__ target, long long n, const float* __restrict__ gout,
               float scale, void* __restrict__ dpred) {
  const float g = __ldg(gout) * scale;
  const long long n8 = n >> 3;
  for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < n8; j += (long long)gridDim.x * 256) {
    float a[8], b[8];
    ld8<PDT>(pred, j * 8, a);
    ld8<TDT>(target, j * 8, b);
    if constexpr (PDT == 2) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(dpred) + j * 8);
      o[0] = make_float4((a[0] - b[0]) * g, (a[1] - b[1]) * g, (a[2] - b[2]) * g, (a[3] - b[3]) * g);
      o[1] = make_float4((a[4] - b[4]) * g, (a[5] - b[5]) * g, (a[6] - b[6]) * g, (a[7] - b[7]) * g);
    } else {
      uint4 q;
      q.x = H16<PDT == 0>::pack((a[0] - b[0]) * g, (a[1] - b[1]) * g);
      q.y = H16<PDT == 0>::pack((a[2] - b[2]) * g, (a[3] - b[3]) * g);
      q.z = H16<PDT == 0>::pack((a[4] - b[4]) * g, (a[5] - b[5]) * g);
      q.w = H16<PDT == 0>::pack((a[6] - b[6]) * g, (a[7] - b[7]) * g);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(dpred) + j * 8) = q;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 7)) {
    const long long i = (n8 << 3) + threadIdx.x;
    st_any(dpred, i, PDT, (ld_any(pred, i, PDT) - ld_any(target, i, TDT)) * g);
  }
}

static bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

}  // namespace vb

using namespace vb;

#define MSE_DISPATCH(pdt, tdt, ...)                                                                   \
  do {                                                                                                \
    if (pdt == 0 && tdt == 0) { constexpr int P = 0, T = 0; __VA_ARGS__; }                            \
    else if (pdt == 0 && tdt == 1) { constexpr int P = 0, T = 1; __VA_ARGS__; }                       \
    else if (pdt == 0 && tdt == 2) { constexpr int P = 0, T = 2; __VA_ARGS__; }                       \
    else if (pdt == 1 && tdt == 0) { constexpr int P = 1, T = 0; __VA_ARGS__; }                       \
    else if (pdt == 1 && tdt == 1) { constexpr int P = 1, T = 1; __VA_ARGS__; }                       \
    else if (pdt == 1 && tdt == 2) { constexpr int P = 1, T = 2; __VA_ARGS__; }                       \
    else if (pdt == 2 && tdt == 0) { constexpr int P = 2, T = 0; __VA_ARGS__; }                       \
    else if (pdt == 2 && tdt == 1) { constexpr int P = 2, T = 1; __VA_ARGS__; }                       \
    else if (pdt == 2 && tdt == 2) { constexpr int P = 2, T = 2; __VA_ARGS__; }                       \
    else return fail(VB200_ERR_UNSUPPORTED, "dtypes %d / %d", pdt, tdt);                              \
  } while (0)

extern "C" int vb200_mse_sum(const void* pred, const void* target, int pred_dtype, int target_dtype, int64_t n, float scale,
                             float* sum, vb200_stream_t stream) {
  VB_REQUIRE(pred && target && sum, "null pointer");
  VB_REQUIRE(n > 0, "empty tensors");
  VB_SUPPORTED(aligned32(pred) && aligned32(target), "pred / target must be 32-byte aligned");
  const unsigned blocks = (unsigned)std::min<long long>((n / 8 + 255) / 256 + 1, 148LL * 8);
  cudaStream_t st = (cudaStream_t)stream;
  MSE_DISPATCH(pred_dtype, target_dtype, (mse_sum_kernel<P, T><<<blocks, 256, 0, st>>>(pred, target, n, scale, sum)));
  return check_launch("vb200_mse_sum");
}

extern "C" int vb200_mse_bwd(const void* pred, const void* target, int pred_dtype, int target_dtype, int64_t n,
                             const float* gout, float scale, void* dpred, vb200_stream_t stream) {
  VB_REQUIRE(pred && target && gout && dpred, "null pointer");
  VB_REQUIRE(n > 0, "empty tensors");
  VB_SUPPORTED(aligned32(pred) && aligned32(target) && aligned32(dpred), "pred / target / dpred must be 32-byte aligned");
  const unsigned blocks = (unsigned)std::min<long long>((n / 8 + 255) / 256 + 1, 148LL * 16);
  cudaStream_t st = (cudaStream_t)stream;
  MSE_DISPATCH(pred_dtype, target_dtype, (mse_bwd_kernel<P, T><<<blocks, 256, 0, st>>>(pred, target, n, gout, scale, dpred)));
  return check_launch("vb200_mse_bwd");
}
