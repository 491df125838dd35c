// Depthwise 7x7 convolution (timm ConvNeXtBlock.conv_dw: nn.Conv2d(C, C, 7, padding=3, groups=C); VM/components/blocks.py:60-74
// via timm ConvNeXtStage) on channels-last 16-bit activations: forward, data gradient (same
// kernel, flipped taps, residual gradient folded into the store) and weight / bias gradient.
//
// Bound: the packed fp32 FMA pipe, not HBM (49 MACs per element at 4 B of traffic: decoder stage 2 of BASELINE config 2
// is 1.18 GMAC = 33 us at 128 FMA lanes x 148 SMs x 1.9 GHz against 15 us of HBM time), so the design minimises every
// instruction that is not an FFMA2:
//   * a CTA (4 warps) stages a (rows + 6) x (4 WC + 6) pixel x 64 channel tile in shared memory with 16-byte cp.async
//     (zero-filled outside the image = the padding; 128 B per pixel: coalesced, conflict-free);
//   * lane = channel pair (bf16x2 / fp16x2 word), warp = 4 output columns x a strip of rows; the 49 taps of the lane's
//     channel pair live in registers (98) for the whole kernel;
//   * the warp streams down its rows: each input row is read ONCE from shared memory (10 words per lane) and feeds the
//     7 output rows it touches, held in a 7-deep register ring of 4-wide accumulators (196 FFMA2 per 10 LDS.32);
//     the ring slot of a row is static because the row loop is unrolled by 7.
// The weight gradient is the same walk with the roles exchanged: 49 accumulators in registers, a 7-deep ring of
// dy rows, warps combined through shared memory, one red.global.add.v4.f32 per 4 channels and tap.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vb {

// legacy register-tile kernels (any even C): convnext_sm100.cu
int dwconv7_legacy(const void* x, const float* wt, const float* bias, const void* add, void* y, int B, int H, int W,
                   int C, int dtype, cudaStream_t st);
int dwconv7_wgrad_legacy(const void* x, const void* dy, float* dwt, float* db, int B, int H, int W, int C, int dtype,
                         cudaStream_t st);
int sm_count();

namespace dw {

constexpr int TW = 4;        // output columns per thread
constexpr int NWARP = 4;     // warps per CTA
constexpr int RED_BYTES = NWARP * 49 * 32 * 8 + NWARP * 32 * 8;  // cross-warp combine of the weight gradient

struct Geom {
  int B, H, W, C2;   // C2 = C / 2 channel pairs
  int rsw;           // output rows per warp
  int htiles, wtiles;
};

// (rows x COLS pixels x 64 channels) tile: pixel p of the tile at word p * 32, lane l reads word p * 32 + l
template <int COLS>
__device__ __forceinline__ void fill_tile(uint32_t dst, const uint4* __restrict__ src, int n, int h_first, int w_first,
                                          int rows, int H, int W, int C8, int chunk) {
  // 128 threads = 16 pixels x 8 parts per step: the thread keeps its part and walks pixels 16 apart with a carry from the
  // column into the row (no divisions in the loop)
  constexpr int STEP = NWARP * 32 / 8;
  const int part = threadIdx.x & 7;
  const int c8 = chunk * 8 + part;
  const bool cok = c8 < C8;
  int pix = threadIdx.x >> 3;
  int r = pix / COLS, c = pix - r * COLS;
  const int total = rows * COLS;
  uint32_t d = dst + threadIdx.x * 16;
  const uint4* rowp = src + ((long long)(n * H + h_first + r) * W + w_first) * C8 + c8;
  for (; pix < total; pix += STEP) {
    const int ih = h_first + r, iw = w_first + c;
    const bool ok = cok && ih >= 0 && ih < H && iw >= 0 && iw < W;
    cp_async_16(d, ok ? rowp + (long long)c * C8 : src, ok ? 16u : 0u);
    d += STEP * 8 * 16;
    c += STEP;
#pragma unroll
    for (int k = 0; k < (STEP + COLS - 1) / COLS; ++k)  // carries: one for the wide tiles, up to four for 4-column tiles
      if (c >= COLS) {
        c -= COLS;
        ++r;
        rowp += (long long)W * C8;
      }
  }
}

template <bool BF16, int WC, bool ADD>
__global__ void __launch_bounds__(NWARP * 32, 2)
dwconv7_tile_kernel(const uint4* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                    const uint4* __restrict__ add, uint32_t* __restrict__ y, const Geom g) {
  constexpr int WR = NWARP / WC, TC = TW * WC + 6, GC = TW * WC;
  extern __shared__ __align__(16) uint32_t tile[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wc = warp % WC, wr = warp / WC;
  int t = blockIdx.x;
  const int wti = t % g.wtiles;
  t /= g.wtiles;
  const int hti = t % g.htiles;
  const int n = t / g.htiles;
  const int chunk = blockIdx.y;
  const int h0 = hti * (WR * g.rsw), w0 = wti * GC;
  const int orows = min(WR * g.rsw, g.H - h0);
  uint32_t* at = tile + (WR * g.rsw + 6) * TC * 32;  // tile of the addend (data gradient: the shortcut's gradient)
  fill_tile<TC>(smem_u32(tile), x, n, h0 - 3, w0 - 3, orows + 6, g.H, g.W, g.C2 / 4, chunk);
  if constexpr (ADD) fill_tile<GC>(smem_u32(at), add, n, h0, w0, orows, g.H, g.W, g.C2 / 4, chunk);
  cp_async_commit();
  // the lane's 49 taps (tap-major fp32 [49][C]: 256 B per warp and tap) while the tile is in flight
  const int cp = chunk * 32 + lane;
  const bool active = cp < g.C2;
  const int C = 2 * g.C2;
  float2 wreg[49];
#pragma unroll
  for (int k = 0; k < 49; ++k)
    wreg[k] = active ? __ldg(reinterpret_cast<const float2*>(wt + k * C) + cp) : make_float2(0.f, 0.f);
  float2 b2 = make_float2(0.f, 0.f);
  if (bias != nullptr && active) b2 = __ldg(reinterpret_cast<const float2*>(bias) + cp);
  cp_async_wait<0>();
  __syncthreads();
  const int r0 = wr * g.rsw;
  const int oh0 = h0 + r0;  // first output row of this warp
  const int rs = min(g.rsw, g.H - oh0);
  const int ow0 = w0 + TW * wc;
  if (!active || rs <= 0 || ow0 >= g.W) return;
  const uint32_t* srow = tile + (r0 * TC + TW * wc) * 32 + lane;
  const uint32_t* arow = at + (r0 * GC + TW * wc) * 32 + lane;
  float2 acc[7][TW];
#pragma unroll
  for (int s = 0; s < 7; ++s)
#pragma unroll
    for (int j = 0; j < TW; ++j) acc[s][j] = b2;
  const long long obase = ((long long)(n * g.H + oh0) * g.W + ow0) * g.C2 + cp;
  const int cols_ok = min(TW, g.W - ow0);
  for (int q = 0; q * 7 < rs + 6; ++q) {
#pragma unroll
    for (int p = 0; p < 7; ++p) {
      const int ir = q * 7 + p;
      if (ir < rs + 6) {  // warp-uniform
        float2 in[TW + 6];
#pragma unroll
        for (int j = 0; j < TW + 6; ++j) in[j] = H16<BF16>::unpack(srow[(ir * TC + j) * 32]);
        // output row o = ir - kh is fed through filter row kh; its ring slot o % 7 = (p - kh) mod 7 is static
#pragma unroll
        for (int kh = 0; kh < 7; ++kh) {
          const int o = ir - kh;
          if (o >= 0 && o < rs) {  // warp-uniform: edge rows skip the filter rows that fall outside the strip
#pragma unroll
            for (int kw = 0; kw < 7; ++kw)
#pragma unroll
              for (int j = 0; j < TW; ++j)
                acc[(p - kh + 7) % 7][j] = __ffma2_rn(in[j + kw], wreg[kh * 7 + kw], acc[(p - kh + 7) % 7][j]);
          }
        }
        const int o = ir - 6;  // this row received its last filter row: store it, recycle the slot
        if (o >= 0) {
          const long long ob = obase + (long long)o * g.W * g.C2;
#pragma unroll
          for (int j = 0; j < TW; ++j) {
            if (j < cols_ok) {
              float2 v = acc[(p + 1) % 7][j];
              if constexpr (ADD) v = __fadd2_rn(v, H16<BF16>::unpack(arow[(o * GC + j) * 32]));
              y[ob + j * g.C2] = H16<BF16>::pack(v.x, v.y);
            }
            acc[(p + 1) % 7][j] = b2;
          }
        }
      }
    }
  }
}

// dwt[tap][c] += sum_pixels dy[p][c] * x[p + tap - 3][c];  db[c] += sum_pixels dy[p][c]
template <bool BF16, int WC>
__global__ void __launch_bounds__(NWARP * 32, 2)
dwconv7_wgrad_tile_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, float* __restrict__ dwt,
                          float* __restrict__ db, const Geom g) {
  constexpr int WR = NWARP / WC, TC = TW * WC + 6, GC = TW * WC;
  extern __shared__ __align__(16) uint32_t tile[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wc = warp % WC, wr = warp / WC;
  int t = blockIdx.x;
  const int wti = t % g.wtiles;
  t /= g.wtiles;
  const int hti = t % g.htiles;
  const int n = t / g.htiles;
  const int chunk = blockIdx.y;
  const int h0 = hti * (WR * g.rsw), w0 = wti * GC;
  const int orows = min(WR * g.rsw, g.H - h0);
  uint32_t* gt = tile + (WR * g.rsw + 6) * TC * 32;  // dy tile behind the x tile
  fill_tile<TC>(smem_u32(tile), x, n, h0 - 3, w0 - 3, orows + 6, g.H, g.W, g.C2 / 4, chunk);
  fill_tile<GC>(smem_u32(gt), dy, n, h0, w0, orows, g.H, g.W, g.C2 / 4, chunk);
  cp_async_commit();
  const int cp = chunk * 32 + lane;
  const bool active = cp < g.C2;
  float2 acc[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) acc[k] = make_float2(0.f, 0.f);
  float2 bsum = make_float2(0.f, 0.f);
  cp_async_wait<0>();
  __syncthreads();
  const int r0 = wr * g.rsw;
  const int rs = min(g.rsw, orows - r0);
  if (active && rs > 0 && w0 + TW * wc < g.W) {
    const uint32_t* srow = tile + (r0 * TC + TW * wc) * 32 + lane;
    const uint32_t* grow = gt + (r0 * GC + TW * wc) * 32 + lane;
    float2 gr[7][TW];  // ring of dy rows: row o in slot o % 7
    for (int q = 0; q * 7 < rs + 6; ++q) {
#pragma unroll
      for (int p = 0; p < 7; ++p) {
        const int ir = q * 7 + p;
        if (ir < rs + 6) {  // warp-uniform
          float2 in[TW + 6];
#pragma unroll
          for (int j = 0; j < TW + 6; ++j) in[j] = H16<BF16>::unpack(srow[(ir * TC + j) * 32]);
          if (ir < rs) {
#pragma unroll
            for (int j = 0; j < TW; ++j) {
              gr[p][j] = H16<BF16>::unpack(grow[(ir * GC + j) * 32]);
              bsum = __fadd2_rn(bsum, gr[p][j]);
            }
          }
          // dy row o = ir - kh pairs with this x row through filter row kh; j outermost: the seven accumulators of a
          // filter row touched by consecutive FFMA2 are all different
#pragma unroll
          for (int kh = 0; kh < 7; ++kh) {
            const int o = ir - kh;
            if (o >= 0 && o < rs) {
#pragma unroll
              for (int j = 0; j < TW; ++j)
#pragma unroll
                for (int kw = 0; kw < 7; ++kw)
                  acc[kh * 7 + kw] = __ffma2_rn(gr[(p - kh + 7) % 7][j], in[j + kw], acc[kh * 7 + kw]);
            }
          }
        }
      }
    }
  }
  // combine the four warps: red[warp][tap][lane] float2, then one vector reduction per 4 channels and tap
  __syncthreads();
  float2* red = reinterpret_cast<float2*>(tile);
#pragma unroll
  for (int k = 0; k < 49; ++k) red[(warp * 49 + k) * 32 + lane] = acc[k];
  float2* redb = red + NWARP * 49 * 32;
  redb[warp * 32 + lane] = bsum;
  __syncthreads();
  const int C = 2 * g.C2;
  for (int item = threadIdx.x; item < 50 * 16; item += NWARP * 32) {
    const int k = item >> 4, l2 = item & 15;  // k == 49: the bias row
    const int c = chunk * 64 + 4 * l2;
    if (c >= C) continue;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
      const float4 v = k < 49 ? *reinterpret_cast<const float4*>(red + (w * 49 + k) * 32 + 2 * l2)
                              : *reinterpret_cast<const float4*>(redb + w * 32 + 2 * l2);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* dst = k < 49 ? dwt + (long long)k * C + c : db + c;
    if (k < 49 || db != nullptr)
      atomicAdd(reinterpret_cast<float4*>(dst), s);  // RED.ADD.F32x4
  }
}

// tile shape: WC warps across the columns (4 columns each), 4 / WC down the rows; rows per warp halved until the grid
// covers the SMs twice over (small feature maps are latency-bound: short CTAs, many of them)
struct Plan {
  int wc, rsw, htiles, wtiles, chunks;
  size_t smem_fwd, smem_wgrad;
};
static Plan make_plan(int B, int H, int W, int C, int max_rsw) {
  Plan p;
  p.wc = W > 8 ? 4 : (W > 4 ? 2 : 1);
  const int wr = NWARP / p.wc;
  p.chunks = (C + 63) / 64;
  p.wtiles = (W + TW * p.wc - 1) / (TW * p.wc);
  int rsw = max_rsw;
  while (rsw > (H + wr - 1) / wr && rsw > 1) rsw = (rsw + 1) / 2;
  const long long target = 2LL * sm_count();  // one full wave of 2 CTAs per SM; small maps are latency-bound
  while (rsw > 2 && (long long)B * ((H + wr * rsw - 1) / (wr * rsw)) * p.wtiles * p.chunks < target) rsw = (rsw + 1) / 2;
  p.rsw = rsw;
  p.htiles = (H + wr * rsw - 1) / (wr * rsw);
  const int tc = TW * p.wc + 6;
  p.smem_fwd = (size_t)(wr * rsw + 6) * tc * 128;
  p.smem_wgrad = p.smem_fwd + (size_t)(wr * rsw) * (TW * p.wc) * 128;
  if (p.smem_wgrad < (size_t)RED_BYTES) p.smem_wgrad = RED_BYTES;
  return p;
}

constexpr int MAX_SMEM = 112 * 1024;  // two CTAs per SM
template <typename K>
static int set_smem(K kern, PerDeviceOnce& once) {
  const int dev = PerDeviceOnce::device();
  if (once.need(dev)) {  // opt-in above 48 KB, once per kernel and device
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);
    if (e != cudaSuccess) return fail(VB200_ERR_CUDA, "dwconv7 smem attribute: %s", cudaGetErrorString(e));
    once.done(dev);
  }
  return VB200_OK;
}

template <bool BF, int WC, bool ADD>
static int launch_fwd_k(const void* x, const float* wt, const float* bias, const void* add, void* y, const Geom& g,
                        const Plan& p, cudaStream_t st) {
  auto kern = dwconv7_tile_kernel<BF, WC, ADD>;
  static PerDeviceOnce once;
  if (int rc = set_smem(kern, once)) return rc;
  const size_t smem = ADD ? p.smem_wgrad : p.smem_fwd;  // data gradient: the addend's tile rides along
  if (smem > (size_t)MAX_SMEM) return fail(VB200_ERR_UNSUPPORTED, "dwconv7 tile needs %zu B of shared memory", smem);
  dim3 grid((unsigned)(g.B * p.htiles * p.wtiles), (unsigned)p.chunks);
  kern<<<grid, NWARP * 32, smem, st>>>((const uint4*)x, wt, bias, (const uint4*)add, (uint32_t*)y, g);
  return VB200_OK;
}
template <bool BF, int WC>
static int launch_fwd(const void* x, const float* wt, const float* bias, const void* add, void* y, const Geom& g,
                      const Plan& p, cudaStream_t st) {
  return add != nullptr ? launch_fwd_k<BF, WC, true>(x, wt, bias, add, y, g, p, st)
                        : launch_fwd_k<BF, WC, false>(x, wt, bias, add, y, g, p, st);
}

template <bool BF, int WC>
static int launch_wgrad(const void* x, const void* dy, float* dwt, float* db, const Geom& g, const Plan& p,
                        cudaStream_t st) {
  auto kern = dwconv7_wgrad_tile_kernel<BF, WC>;
  static PerDeviceOnce once;
  if (int rc = set_smem(kern, once)) return rc;
  if (p.smem_wgrad > (size_t)MAX_SMEM) return fail(VB200_ERR_UNSUPPORTED, "dwconv7 wgrad tile needs %zu B of shared memory", p.smem_wgrad);
  dim3 grid((unsigned)(g.B * p.htiles * p.wtiles), (unsigned)p.chunks);
  kern<<<grid, NWARP * 32, p.smem_wgrad, st>>>((const uint4*)x, (const uint4*)dy, dwt, db, g);
  return VB200_OK;
}

}  // namespace dw
}  // namespace vb

using namespace vb;

#define DW_DISPATCH(dtype, wc, CALL)                                                   \
  do {                                                                                 \
    int rc_ = VB200_OK;                                                                \
    if ((dtype) == VB200_BF16) {                                                       \
      constexpr bool BF = true;                                                        \
      if ((wc) == 4) { constexpr int WC = 4; rc_ = CALL; }                             \
      else if ((wc) == 2) { constexpr int WC = 2; rc_ = CALL; }                        \
      else { constexpr int WC = 1; rc_ = CALL; }                                       \
    } else if ((dtype) == VB200_FP16) {                                                \
      constexpr bool BF = false;                                                       \
      if ((wc) == 4) { constexpr int WC = 4; rc_ = CALL; }                             \
      else if ((wc) == 2) { constexpr int WC = 2; rc_ = CALL; }                        \
      else { constexpr int WC = 1; rc_ = CALL; }                                       \
    } else {                                                                           \
      return vb::fail(VB200_ERR_UNSUPPORTED, "dtype %d", (int)(dtype));                \
    }                                                                                  \
    if (rc_) return rc_;                                                               \
  } while (0)

extern "C" int vb200_dwconv7(const void* x, const float* wt, const float* bias, const void* add, void* y, int B, int H,
                             int W, int C, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && wt && y, "null pointer");
  VB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "bad extent");
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 8 != 0) return dwconv7_legacy(x, wt, bias, add, y, B, H, W, C, dtype, st);
  VB_SUPPORTED((long long)B * H * W * C < (1LL << 31), "tensor must have < 2^31 elements");
  const dw::Plan p = dw::make_plan(B, H, W, C, 16);
  dw::Geom g{B, H, W, C / 2, p.rsw, p.htiles, p.wtiles};
  DW_DISPATCH(dtype, p.wc, (dw::launch_fwd<BF, WC>(x, wt, bias, add, y, g, p, st)));
  return check_launch("vb200_dwconv7");
}

extern "C" int vb200_dwconv7_wgrad(const void* x, const void* dy, float* dwt, float* db, int B, int H, int W, int C,
                                   int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && dy && dwt, "null pointer");
  VB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "bad extent");
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 8 != 0) return dwconv7_wgrad_legacy(x, dy, dwt, db, B, H, W, C, dtype, st);
  VB_SUPPORTED((long long)B * H * W * C < (1LL << 31), "tensor must have < 2^31 elements");
  const dw::Plan p = dw::make_plan(B, H, W, C, 16);
  dw::Geom g{B, H, W, C / 2, p.rsw, p.htiles, p.wtiles};
  DW_DISPATCH(dtype, p.wc, (dw::launch_wgrad<BF, WC>(x, dy, dwt, db, g, p, st)));
  return check_launch("vb200_dwconv7_wgrad");
}
