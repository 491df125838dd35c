// Prediction path (CY/engine.py:61-71, 711-805; VU/callbacks/prediction_writer.py:74-111): the per-window epilogue of
// sliding-window inference in one pass - centre-crop the (padded) window prediction, cast, and blend it into the output
// volume's Z range with the linear feathering factors of _blend_in.  HBM-bound; one read of the window, one
// read-modify-write of the output slab.
#include <type_traits>

#include "common.cuh"

namespace vb {

// dst [B,C,Z,H,W] (any dtype), src [B,C,d,Hs,Ws] (any dtype), window at z0, crop offset (oy, ox).
// _blend_in: z0 == 0 -> dst = src;  else samples = min(z0 + 1, d), f_i = min(d - i, samples), dst = dst (f-1)/f + src / f
// grid = (x chunks of 1024, H, B*C*d): one division per block, four elements per thread, coalesced rows
template <int DDT, int SDT>
__global__ void __launch_bounds__(256)
blend_window_kernel(void* __restrict__ dst, const void* __restrict__ src, int Z, int H, int W, int d, int Hs, int Ws, int z0,
                    int oy, int ox) {
  const int y = blockIdx.y;
  const int z = blockIdx.z % d;
  const long long bc = blockIdx.z / d;
  const long long so = ((bc * d + z) * Hs + y + oy) * (long long)Ws + ox;
  const long long o = ((bc * Z + z0 + z) * H + y) * (long long)W;
  const float f = (float)min(d - z, min(z0 + 1, d));
  const float a = (f - 1.f) / f, b = 1.f / f;
  auto ld = [](const void* p, long long i, auto tag) -> float {
    constexpr int DT = decltype(tag)::value;
    if constexpr (DT == 2) return reinterpret_cast<const float*>(p)[i];
    else if constexpr (DT == 0) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    else return __half2float(reinterpret_cast<const __half*>(p)[i]);
  };
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int x = blockIdx.x * 1024 + u * 256 + threadIdx.x;
    if (x >= W) continue;
    const float s = ld(src, so + x, std::integral_constant<int, SDT>{});
    float v = s;
    if (z0 != 0) v = ld(dst, o + x, std::integral_constant<int, DDT>{}) * a + s * b;
    st_any(dst, o + x, DDT, v);
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_blend_window(void* dst, const void* src, int dst_dtype, int src_dtype, int64_t BC, int Z, int H, int W,
                                  int d, int Hs, int Ws, int z0, int oy, int ox, vb200_stream_t stream) {
  VB_REQUIRE(dst && src, "null pointer");
  VB_SUPPORTED(dst_dtype >= 0 && dst_dtype <= 2 && src_dtype >= 0 && src_dtype <= 2, "dtypes %d / %d", dst_dtype, src_dtype);
  VB_REQUIRE(z0 >= 0 && d > 0 && z0 + d <= Z, "window [%d, %d) outside a volume of depth %d", z0, z0 + d, Z);
  VB_REQUIRE(oy >= 0 && ox >= 0 && oy + H <= Hs && ox + W <= Ws, "crop (%d,%d)+(%d,%d) outside (%d,%d)", oy, ox, H, W, Hs, Ws);
  if ((long long)BC * d * H * W <= 0) return VB200_OK;
  VB_SUPPORTED((long long)BC * d < 65536 && H < 65536, "blend grid: B*C*d = %lld, H = %d", (long long)BC * d, H);
  dim3 grid((W + 1023) / 1024, H, (unsigned)(BC * d));
  cudaStream_t st = (cudaStream_t)stream;
#define BW(DD, SD)                                                                                                     \
  if (dst_dtype == DD && src_dtype == SD) blend_window_kernel<DD, SD><<<grid, 256, 0, st>>>(dst, src, Z, H, W, d, Hs, Ws, z0, oy, ox);
  BW(0, 0) BW(0, 1) BW(0, 2) BW(1, 0) BW(1, 1) BW(1, 2) BW(2, 0) BW(2, 1) BW(2, 2)
#undef BW
  return check_launch("blend_window");
}
