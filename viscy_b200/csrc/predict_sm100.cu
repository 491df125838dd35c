// Prediction path (CY/engine.py:61-71, 711-805; VU/callbacks/prediction_writer.py:74-111): the per-window epilogue of
// sliding-window inference in one pass - centre-crop the (padded) window prediction, cast, and blend it into the output
// volume's Z range with the linear feathering factors of _blend_in.  HBM-bound; one read of the window, one
// read-modify-write of the output slab.
#include "common.cuh"

namespace vb {

// dst [B,C,Z,H,W] (any dtype), src [B,C,d,Hs,Ws] (any dtype), window at z0, crop offset (oy, ox).
// _blend_in: z0 == 0 -> dst = src;  else samples = min(z0 + 1, d), f_i = min(d - i, samples), dst = dst (f-1)/f + src / f
__global__ void __launch_bounds__(256)
blend_window_kernel(void* __restrict__ dst, const void* __restrict__ src, int ddt, int sdt, int BC, int Z, int H, int W, int d,
                    int Hs, int Ws, int z0, int oy, int ox) {
  const long long n = (long long)BC * d * H * W;
  const int samples = min(z0 + 1, d);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    long long t = i / W;
    const int y = (int)(t % H);
    t /= H;
    const int z = (int)(t % d);
    const long long bc = t / d;
    const float s = ld_any(src, ((bc * d + z) * Hs + y + oy) * (long long)Ws + x + ox, sdt);
    const long long o = ((bc * Z + z0 + z) * H + y) * (long long)W + x;
    float v = s;
    if (z0 != 0) {
      const float f = (float)min(d - z, samples);
      v = ld_any(dst, o, ddt) * (f - 1.f) / f + s / f;
    }
    st_any(dst, o, ddt, v);
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
  const long long n = (long long)BC * d * H * W;
  if (n <= 0) return VB200_OK;
  long long want = (n + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148LL * 16 ? want : 148LL * 16);
  blend_window_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, dst_dtype, src_dtype, (int)BC, Z, H, W, d, Hs, Ws,
                                                                 z0, oy, ox);
  return check_launch("blend_window");
}
