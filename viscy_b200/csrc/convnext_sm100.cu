// HBM-bound kernels of the ConvNeXt / ConvNeXt-V2 block on channels-last 16-bit activations:
// depthwise 7x7 (fwd, dgrad, wgrad), LayerNorm over C (fwd, bwd), GELU + Global Response Norm
// (fwd two-phase, bwd two-phase), column sums (bias gradients).
// Reference semantics: timm ConvNeXtBlock / LayerNorm2d / GlobalResponseNormMlp as composed by
// VM/unet/unext2.py:40-49 and VM/components/blocks.py:54-74.
#include "common.cuh"

namespace vb {

int dwconv7_legacy(const void* x, const float* wt, const float* bias, const void* add, void* y, int B, int H, int W,
                   int C, int dtype, cudaStream_t st);
int dwconv7_wgrad_legacy(const void* x, const void* dy, float* dwt, float* db, int B, int H, int W, int C, int dtype,
                         cudaStream_t st);

// ------------------------------------------------------------------------------ depthwise 7x7
// x [B,H,W,C] 16-bit, wt [49][C] fp32 (tap-major so that channel loads coalesce), bias [C] or null,
// add [B,H,W,C] 16-bit or null (residual gradient folded into the dgrad call).
// One thread = 2 adjacent channels x (DW_TH x DW_TW) output pixels, accumulators as float2 (channel pair) so every
// multiply-add is one packed FFMA2 (sm_100 fma.rn.f32x2).  The thread walks the DW_TH+6 input rows once, fully
// unrolled (static tap indices, loads of the next row overlap the FMAs of the current one); each input row
// (DW_TW+6 values) feeds up to DW_TH output rows.  The 49 taps of the block's 128 channel pairs sit in shared
// memory (conflict-free 8-byte reads).  All offsets are 32-bit (host checks numel < 2^31).
constexpr int DW_SMEM = 49 * 128 * 8;

template <bool BF16>
__device__ __forceinline__ float2 ldpair(const uint32_t* p, int off, bool ok) {
  uint32_t raw = 0;
  if (ok) raw = __ldg(p + off);
  return H16<BF16>::unpack(raw);
}

template <bool BF16, int DW_TH, int DW_TW, bool SMEM_W>
__global__ void __launch_bounds__(128)
dwconv7_kernel(const uint32_t* __restrict__ x, const float* __restrict__ wt,
               const float* __restrict__ bias, const uint32_t* __restrict__ add,
               uint32_t* __restrict__ y, int B, int H, int W, int C2) {
  extern __shared__ float2 sw[];  // [49][128]
  const int cp0 = blockIdx.y * 128;
  const int cp = cp0 + threadIdx.x;
  const int C = C2 * 2;
  if constexpr (SMEM_W) {  // big tiles: stage the block's 49 x 128 taps once (50 KB)
    for (int i = threadIdx.x; i < 49 * 128; i += 128) {
      const int tap = i >> 7, c = cp0 + (i & 127);
      sw[i] = c < C2 ? __ldg(reinterpret_cast<const float2*>(wt + tap * C) + c) : make_float2(0.f, 0.f);
    }
    __syncthreads();
  }
  if (cp >= C2) return;
  const int wtiles = (W + DW_TW - 1) / DW_TW, htiles = (H + DW_TH - 1) / DW_TH;
  int t = blockIdx.x;
  const int w0 = (t % wtiles) * DW_TW;
  t /= wtiles;
  const int h0 = (t % htiles) * DW_TH;
  const int n = t / htiles;
  const int pix0 = n * H * W;  // pixel index of the image origin
  const uint32_t* xb = x + cp;
  int coff[DW_TW + 6];
  bool cok[DW_TW + 6];
#pragma unroll
  for (int j = 0; j < DW_TW + 6; ++j) {
    const int iw = w0 + j - 3;
    cok[j] = iw >= 0 && iw < W;
    coff[j] = (pix0 + iw) * C2;
  }
  float2 acc[DW_TH][DW_TW];
  float2 b2 = make_float2(0.f, 0.f);
  if (bias != nullptr) b2 = *reinterpret_cast<const float2*>(bias + 2 * cp);
#pragma unroll
  for (int r = 0; r < DW_TH; ++r)
#pragma unroll
    for (int j = 0; j < DW_TW; ++j) acc[r][j] = b2;
  const float2* swl = sw + threadIdx.x;
#pragma unroll
  for (int ir = 0; ir < DW_TH + 6; ++ir) {
    const int ih = h0 + ir - 3;
    const bool rok = ih >= 0 && ih < H;
    const int roff = ih * W * C2;
    float2 in[DW_TW + 6];
#pragma unroll
    for (int j = 0; j < DW_TW + 6; ++j) in[j] = ldpair<BF16>(xb, roff + coff[j], rok && cok[j]);
#pragma unroll
    for (int r = 0; r < DW_TH; ++r) {
      const int kh = ir - r;  // static after unrolling
      if (kh >= 0 && kh < 7) {
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
          // small tiles read the taps through L1 instead: staging 50 KB per block would dominate
          const float2 wv = SMEM_W ? swl[(kh * 7 + kw) * 128]
                                   : __ldg(reinterpret_cast<const float2*>(wt + (kh * 7 + kw) * C) + cp);
#pragma unroll
          for (int j = 0; j < DW_TW; ++j) acc[r][j] = __ffma2_rn(in[j + kw], wv, acc[r][j]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < DW_TH; ++r) {
    const int oh = h0 + r;
    if (oh < H) {
#pragma unroll
      for (int j = 0; j < DW_TW; ++j) {
        const int ow = w0 + j;
        if (ow < W) {
          const int o = ((pix0 + oh * W) + ow) * C2 + cp;
          float2 v = acc[r][j];
          if (add != nullptr) v = __fadd2_rn(v, H16<BF16>::unpack(__ldg(add + o)));
          y[o] = H16<BF16>::pack(v.x, v.y);
        }
      }
    }
  }
}

// wgrad: dwt[tap][c] += sum_{pixels} dy[p][c] * x[p + tap offset][c];  db[c] += sum dy[p][c].
// One thread = 2 channels with all 49 float2 accumulators in registers; it walks a strip of rows of one image
// two output rows x 8 pixels at a time: the 8 x-rows h-3..h+4 (14 values each) feed both output rows
// (98 x 8 packed FFMA2 per 16 + 112 loads).
template <bool BF16>
__global__ void __launch_bounds__(128)
dwconv7_wgrad_kernel(const uint32_t* __restrict__ x, const uint32_t* __restrict__ dy,
                     float* __restrict__ dwt, float* __restrict__ db, int B, int H, int W, int C2,
                     int rows_per_block) {
  const int cp = blockIdx.y * 128 + threadIdx.x;
  if (cp >= C2) return;
  const int strips = (H + rows_per_block - 1) / rows_per_block;
  const int n = blockIdx.x / strips;
  const int h0 = (blockIdx.x % strips) * rows_per_block;
  const int h1 = min(H, h0 + rows_per_block);
  const int C = 2 * C2;
  const int pix0 = n * H * W;
  const uint32_t* xb = x + cp;
  const uint32_t* gb = dy + cp;
  float2 acc[7][7];
#pragma unroll
  for (int a = 0; a < 7; ++a)
#pragma unroll
    for (int b = 0; b < 7; ++b) acc[a][b] = make_float2(0.f, 0.f);
  float2 bsum = make_float2(0.f, 0.f);
  for (int h = h0; h < h1; h += 2) {
    const bool row1 = h + 1 < h1;
    for (int w0 = 0; w0 < W; w0 += 8) {
      float2 g0[8], g1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = w0 + j < W;
        const int off = (pix0 + h * W + w0 + j) * C2;
        g0[j] = ldpair<BF16>(gb, off, ok);
        g1[j] = ldpair<BF16>(gb, off + W * C2, ok && row1);
        bsum = __fadd2_rn(bsum, __fadd2_rn(g0[j], g1[j]));
      }
#pragma unroll
      for (int xr = 0; xr < 8; ++xr) {
        const int ih = h - 3 + xr;
        const bool rok = ih >= 0 && ih < H;
        const int roff = (pix0 + ih * W + w0 - 3) * C2;
        float2 in[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) {
          const int iw = w0 + j - 3;
          in[j] = ldpair<BF16>(xb, roff + j * C2, rok && iw >= 0 && iw < W);
        }
        if (xr < 7) {
#pragma unroll
          for (int kw = 0; kw < 7; ++kw)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[xr][kw] = __ffma2_rn(g0[j], in[j + kw], acc[xr][kw]);
        }
        if (xr >= 1) {
#pragma unroll
          for (int kw = 0; kw < 7; ++kw)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[xr - 1][kw] = __ffma2_rn(g1[j], in[j + kw], acc[xr - 1][kw]);
        }
      }
    }
  }
#pragma unroll
  for (int kh = 0; kh < 7; ++kh)
#pragma unroll
    for (int kw = 0; kw < 7; ++kw) {
      atomicAdd(dwt + (kh * 7 + kw) * C + 2 * cp, acc[kh][kw].x);
      atomicAdd(dwt + (kh * 7 + kw) * C + 2 * cp + 1, acc[kh][kw].y);
    }
  if (db != nullptr) {
    atomicAdd(db + 2 * cp, bsum.x);
    atomicAdd(db + 2 * cp + 1, bsum.y);
  }
}

// ------------------------------------------------------------------------------ LayerNorm over C
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row; x, y [M, C] 16-bit (ld = C); stats fp32 [M]
template <bool BF16>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const uint32_t* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, uint32_t* __restrict__ y,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, long long M, int C2,
                     float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const uint32_t* xr = x + row * C2;
  float s = 0.f;
  for (int i = lane; i < C2; i += 32) {
    const float2 v = H16<BF16>::unpack(__ldg(xr + i));
    s += v.x + v.y;
  }
  const float mean = warp_sum(s) / (2.0f * C2);
  float q = 0.f;
  for (int i = lane; i < C2; i += 32) {
    const float2 v = H16<BF16>::unpack(__ldg(xr + i));
    const float a = v.x - mean, b = v.y - mean;
    q += a * a + b * b;
  }
  const float rstd = rsqrtf(warp_sum(q) / (2.0f * C2) + eps);
  for (int i = lane; i < C2; i += 32) {
    const float2 v = H16<BF16>::unpack(__ldg(xr + i));
    const float2 g = __ldg(reinterpret_cast<const float2*>(gamma) + i);
    const float2 b = __ldg(reinterpret_cast<const float2*>(beta) + i);
    y[row * C2 + i] = H16<BF16>::pack((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y);
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}


// One warp per row with the row held in registers (NV 16-byte vectors per lane): one HBM read, exact two-pass statistics.
// y has row pitch ldy8 (in 16-byte vectors).  `ones`: the 8 columns [C, C+8) of y are written {1,0,...,0} -- the weight
// gradient GEMM dW1 = dh^T [l | 1] then delivers the fc1 bias gradient as its extra column (no column-sum pass over the
// hidden tensor).  ones2 (row pitch ld2_8, vector column c2_8): a second matrix that receives the same group per row
// (the block's GELU output buffer: its ones column makes the fc2 weight-gradient GEMM produce the fc2 bias gradient).
template <bool BF16, int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_rows_kernel(const uint4* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                          uint4* __restrict__ y, long long ldy8, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                          long long M, int C8, float eps, int ones, uint4* __restrict__ ones2, long long ld2_8, int c2_8) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < C8) {
      const uint4 q = __ldg(x + row * C8 + i);
      const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = H16<BF16>::unpack(w4[j]);
        v[k][2 * j] = f.x;
        v[k][2 * j + 1] = f.y;
        s += f.x + f.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
    }
  }
  const float inv_c = 1.0f / (8.0f * C8);
  const float mean = warp_sum(s) * inv_c;
  float qs = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (lane + 32 * k < C8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = v[k][j] - mean;
        qs = fmaf(a, a, qs);
      }
    }
  const float rstd = rsqrtf(warp_sum(qs) * inv_c + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < C8) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * i);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * i + 1);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf((v[k][j] - mean) * rstd, gm[j], bt[j]);
      y[row * ldy8 + i] = make_uint4(H16<BF16>::pack(o[0], o[1]), H16<BF16>::pack(o[2], o[3]),
                                     H16<BF16>::pack(o[4], o[5]), H16<BF16>::pack(o[6], o[7]));
    }
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
    // `ones` 16-byte groups behind the row: {1, 0 x 7}, then zeros (16 columns keep the row pitch a multiple of 32 B, so
    // that the 128-byte TMA box rows of the weight-gradient GEMM stay sector-aligned)
    const uint4 one = make_uint4(H16<BF16>::pack(1.0f, 0.0f), 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
    for (int k = 0; k < ones; ++k) {
      y[row * ldy8 + C8 + k] = k == 0 ? one : zero;
      if (ones2 != nullptr) ones2[row * ld2_8 + c2_8 + k] = k == 0 ? one : zero;
    }
  }
}

// LayerNorm backward, split in two HBM-friendly kernels:
//  rows:    dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma   (one warp per row, 16-byte loads)
//  columns: dgamma[c] += sum_rows dy * xhat, dbeta[c] += sum_rows dy             (128 x 4 thread blocks, smem combine)
template <bool BF16, int NV>  // NV = 16-byte vectors per lane (C <= 256 * NV)
__global__ void __launch_bounds__(256)
layernorm_bwd_rows_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ gamma, uint4* __restrict__ dx,
                          long long M, int C8) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
  float g[NV][8], xh[NV][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = lane + 32 * v;
#pragma unroll
    for (int k = 0; k < 8; ++k) g[v][k] = xh[v][k] = 0.f;
    if (i < C8) {
      const uint4 qd = __ldg(dy + row * C8 + i), qx = __ldg(x + row * C8 + i);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i + 1);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 d = H16<BF16>::unpack(wd[k]), xv = H16<BF16>::unpack(wx[k]);
        xh[v][2 * k] = (xv.x - mu) * rs;
        xh[v][2 * k + 1] = (xv.y - mu) * rs;
        g[v][2 * k] = d.x * gm[2 * k];
        g[v][2 * k + 1] = d.y * gm[2 * k + 1];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s1 += g[v][k];
        s2 = fmaf(g[v][k], xh[v][k], s2);
      }
    }
  }
  const float inv_c = 1.0f / (8.0f * C8);
  s1 = warp_sum(s1) * inv_c;
  s2 = warp_sum(s2) * inv_c;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = lane + 32 * v;
    if (i < C8) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rs * (g[v][k] - s1 - xh[v][k] * s2);
      dx[row * C8 + i] = make_uint4(H16<BF16>::pack(o[0], o[1]), H16<BF16>::pack(o[2], o[3]),
                                    H16<BF16>::pack(o[4], o[5]), H16<BF16>::pack(o[6], o[7]));
    }
  }
}

// LayerNorm backward in ONE pass over dy and x: each warp walks rows with a grid stride, writes dx for its row and keeps
// the dgamma / dbeta partial sums of the columns its lanes own in registers; warps of a block combine in shared memory,
// one atomicAdd per (block, column).  Replaces the rows + columns kernel pair (a second read of dy and x) for C <= 768.
template <bool BF16, int NV>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_fused_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const float* __restrict__ mean,
                           const float* __restrict__ rstd, const float* __restrict__ gamma, uint4* __restrict__ dx,
                           float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C8) {
  extern __shared__ float red[];  // [2][C]
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int C = 8 * C8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float pg[NV][8], pb[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int k = 0; k < 8; ++k) pg[v][k] = pb[v][k] = 0.f;
  const float inv_c = 1.0f / (float)C;
  // Rows of one warp are a serial chain (load -> two warp reductions -> store): with <= 2 vectors per lane the next row's
  // operands are requested before the current row is reduced, so its HBM latency runs under the math (PRE); with 3 vectors
  // the register file has no room for a second row.
  constexpr bool PRE = NV <= 2;
  const long long rstride = (long long)gridDim.x * warps;
  long long row = (long long)blockIdx.x * warps + (threadIdx.x >> 5);
  uint4 nd[PRE ? NV : 1], nx[PRE ? NV : 1];
  float nmu = 0.f, nrs = 0.f;
  if constexpr (PRE) {
    if (row < M) {
      nmu = __ldg(mean + row);
      nrs = __ldg(rstd + row);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (lane + 32 * v < C8) {
          nd[v] = __ldg(dy + row * C8 + lane + 32 * v);
          nx[v] = __ldg(x + row * C8 + lane + 32 * v);
        }
    }
  }
  for (; row < M; row += rstride) {
    float mu, rs;
    uint4 cd[NV], cx[NV];
    if constexpr (PRE) {
      mu = nmu;
      rs = nrs;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        cd[v] = nd[v];
        cx[v] = nx[v];
      }
      const long long rn = row + rstride;
      if (rn < M) {
        nmu = __ldg(mean + rn);
        nrs = __ldg(rstd + rn);
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (lane + 32 * v < C8) {
            nd[v] = __ldg(dy + rn * C8 + lane + 32 * v);
            nx[v] = __ldg(x + rn * C8 + lane + 32 * v);
          }
      }
    } else {
      mu = __ldg(mean + row);
      rs = __ldg(rstd + row);
    }
    float g[NV][8], xh[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int i = lane + 32 * v;
#pragma unroll
      for (int k = 0; k < 8; ++k) g[v][k] = xh[v][k] = 0.f;
      if (i < C8) {
        const uint4 qd = PRE ? cd[v] : __ldg(dy + row * C8 + i), qx = PRE ? cx[v] : __ldg(x + row * C8 + i);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i);
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i + 1);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 d = H16<BF16>::unpack(wd[k]), xv = H16<BF16>::unpack(wx[k]);
          xh[v][2 * k] = (xv.x - mu) * rs;
          xh[v][2 * k + 1] = (xv.y - mu) * rs;
          pg[v][2 * k] = fmaf(d.x, xh[v][2 * k], pg[v][2 * k]);
          pg[v][2 * k + 1] = fmaf(d.y, xh[v][2 * k + 1], pg[v][2 * k + 1]);
          pb[v][2 * k] += d.x;
          pb[v][2 * k + 1] += d.y;
          g[v][2 * k] = d.x * gm[2 * k];
          g[v][2 * k + 1] = d.y * gm[2 * k + 1];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s1 += g[v][k];
          s2 = fmaf(g[v][k], xh[v][k], s2);
        }
      }
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int i = lane + 32 * v;
      if (i < C8) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = rs * (g[v][k] - s1 - xh[v][k] * s2);
        dx[row * C8 + i] = make_uint4(H16<BF16>::pack(o[0], o[1]), H16<BF16>::pack(o[2], o[3]),
                                      H16<BF16>::pack(o[4], o[5]), H16<BF16>::pack(o[6], o[7]));
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = lane + 32 * v;
    if (i < C8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        atomicAdd(&red[i * 8 + k], pg[v][k]);
        atomicAdd(&red[C + i * 8 + k], pb[v][k]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
  }
}

template <bool BF16>
__global__ void __launch_bounds__(512)
layernorm_bwd_cols_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const float* __restrict__ mean,
                          const float* __restrict__ rstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                          long long M, int C8, int rows_per_block, int cw_log2) {
  __shared__ float red[16 * 512];
  const int cx = threadIdx.x & ((1 << cw_log2) - 1), ry = threadIdx.x >> cw_log2, RL = 512 >> cw_log2;
  const int c8 = (blockIdx.x << cw_log2) + cx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float a[2][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[0][k] = a[1][k] = 0.f;
  if (c8 < C8) {
#pragma unroll 2
    for (long long r = r0 + ry; r < r1; r += RL) {
      const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
      const uint4 qd = __ldg(dy + r * C8 + c8), qx = __ldg(x + r * C8 + c8);
      const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 d = H16<BF16>::unpack(wd[k]), xv = H16<BF16>::unpack(wx[k]);
        a[0][2 * k] = fmaf(d.x, (xv.x - mu) * rs, a[0][2 * k]);
        a[0][2 * k + 1] = fmaf(d.y, (xv.y - mu) * rs, a[0][2 * k + 1]);
        a[1][2 * k] += d.x;
        a[1][2 * k + 1] += d.y;
      }
    }
  }
  float* const outs[2] = {dgamma, dbeta};
  colred_combine<2>(a, red, cw_log2, blockIdx.x << cw_log2, C8, outs);
}

// legacy single-kernel variant (C not a multiple of 8): each warp walks rows with a grid stride and keeps its
// dgamma/dbeta partials in registers (NP channel pairs per lane)
template <bool BF16, int NP>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const uint32_t* __restrict__ dy, const uint32_t* __restrict__ x,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, uint32_t* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C2) {
  extern __shared__ float red[];  // [2][C]
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int C = 2 * C2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float2 gacc[NP], bacc[NP], gm[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    gacc[k] = make_float2(0.f, 0.f);
    bacc[k] = make_float2(0.f, 0.f);
    const int i = lane + 32 * k;
    gm[k] = i < C2 ? __ldg(reinterpret_cast<const float2*>(gamma) + i) : make_float2(0.f, 0.f);
  }
  for (long long row = (long long)blockIdx.x * warps + (threadIdx.x >> 5); row < M;
       row += (long long)gridDim.x * warps) {
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float2 g[NP], xh[NP];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int i = lane + 32 * k;
      g[k] = make_float2(0.f, 0.f);
      xh[k] = make_float2(0.f, 0.f);
      if (i < C2) {
        const float2 d = H16<BF16>::unpack(__ldg(dy + row * C2 + i));
        const float2 v = H16<BF16>::unpack(__ldg(x + row * C2 + i));
        xh[k] = make_float2((v.x - mu) * rs, (v.y - mu) * rs);
        gacc[k].x += d.x * xh[k].x;
        gacc[k].y += d.y * xh[k].y;
        bacc[k].x += d.x;
        bacc[k].y += d.y;
        g[k] = make_float2(d.x * gm[k].x, d.y * gm[k].y);
        s1 += g[k].x + g[k].y;
        s2 += g[k].x * xh[k].x + g[k].y * xh[k].y;
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int i = lane + 32 * k;
      if (i < C2)
        dx[row * C2 + i] = H16<BF16>::pack(rs * (g[k].x - s1 - xh[k].x * s2), rs * (g[k].y - s1 - xh[k].y * s2));
    }
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int i = lane + 32 * k;
    if (i < C2) {
      atomicAdd(&red[2 * i], gacc[k].x);
      atomicAdd(&red[2 * i + 1], gacc[k].y);
      atomicAdd(&red[C + 2 * i], bacc[k].x);
      atomicAdd(&red[C + 2 * i + 1], bacc[k].y);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
  }
}

// ------------------------------------------------------------------------------ GELU + GRN
// h [B, R, C] 16-bit (R = pixels per sample).  Column kernels: thread = channel pair, block walks a
// chunk of rows of one sample, one atomicAdd per (block, channel) at the end.
// mode 0: sumsq[n,c] += gelu(h)^2
// mode 1: S1[n,c] += dy * gelu(h);  sdy[c] += dy
template <bool BF16, int MODE>
__global__ void __launch_bounds__(128)
grn_reduce_kernel(const uint32_t* __restrict__ h, const uint32_t* __restrict__ dy,
                  float* __restrict__ acc0, float* __restrict__ acc1, int R, int C2,
                  int rows_per_block) {
  const int cp = blockIdx.x * 128 + threadIdx.x;
  if (cp >= C2) return;
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  const long long base = (long long)n * R;
  float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const float2 u = H16<BF16>::unpack(__ldg(h + (base + r) * C2 + cp));
    const float gx = gelu_f(u.x), gy = gelu_f(u.y);
    if (MODE == 0) {
      a.x = fmaf(gx, gx, a.x);
      a.y = fmaf(gy, gy, a.y);
    } else {
      const float2 d = H16<BF16>::unpack(__ldg(dy + (base + r) * C2 + cp));
      a.x = fmaf(d.x, gx, a.x);
      a.y = fmaf(d.y, gy, a.y);
      b.x += d.x;
      b.y += d.y;
    }
  }
  const int C = 2 * C2;
  atomicAdd(acc0 + (long long)n * C + 2 * cp, a.x);
  atomicAdd(acc0 + (long long)n * C + 2 * cp + 1, a.y);
  if (MODE == 1) {
    atomicAdd(acc1 + 2 * cp, b.x);
    atomicAdd(acc1 + 2 * cp + 1, b.y);
  }
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += scratch[i];
  return t;
}

// forward coefficients: s[n,c] = 1 + w[c] * Gx[n,c] / (mean_c Gx[n,:] + eps),  Gx = sqrt(sumsq)
// grid (sample, column chunk of blockDim.x): every block reduces the whole sample row (C values: cheap), writes its chunk
__global__ void grn_coef_fwd_kernel(const float* __restrict__ sumsq, const float* __restrict__ w,
                                    float* __restrict__ s, int C, float eps) {
  __shared__ float scratch[32];
  const int n = blockIdx.x;
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) part += sqrtf(sumsq[(long long)n * C + c]);
  const float m = block_sum(part, scratch) / C;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c < C) s[(long long)n * C + c] = 1.0f + w[c] * sqrtf(sumsq[(long long)n * C + c]) / (m + eps);
}

// backward coefficients from S1[n,c] = sum_r dy * g:
//   dNx = w * S1;  dw[c] += Nx * S1;  dGx = dNx/(m+eps) - (1/C) * sum_c'(dNx * Gx) / (m+eps)^2
//   t[n,c] = dGx / Gx   (so that dg = dy * s + g * t)
// grid (sample, column chunk of blockDim.x): every block reduces the whole sample row for the two scalars (C4 values:
// cheap) and writes its own chunk, instead of one block per sample walking all columns.
__global__ void grn_coef_bwd_kernel(const float* __restrict__ sumsq, const float* __restrict__ S1,
                                    const float* __restrict__ w, float* __restrict__ t,
                                    float* __restrict__ dw, int C, float eps) {
  __shared__ float scratch[32];
  const int n = blockIdx.x;
  float pm = 0.f, pd = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float gx = sqrtf(sumsq[(long long)n * C + c]);
    pm += gx;
    pd += w[c] * S1[(long long)n * C + c] * gx;
  }
  const float m = block_sum(pm, scratch) / C;
  const float dot = block_sum(pd, scratch);
  const float inv = 1.0f / (m + eps);
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c < C) {
    const float gx = sqrtf(sumsq[(long long)n * C + c]);
    const float s1 = S1[(long long)n * C + c];
    const float dgx = w[c] * s1 * inv - dot * inv * inv / C;
    t[(long long)n * C + c] = gx > 0.f ? dgx / gx : 0.f;
    atomicAdd(dw + c, gx * inv * s1);
  }
}

// y = gelu(h) * s[n,c] + b[c]
template <bool BF16>
__global__ void __launch_bounds__(256)
grn_apply_fwd_kernel(const uint4* __restrict__ h, const float* __restrict__ s,
                     const float* __restrict__ b, uint4* __restrict__ y, int R, int C8,
                     long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = (int)(i % C8);
  const long long row = i / C8;
  const int n = (int)(row / R);
  const uint4 q = __ldg(h + i);
  const uint32_t in[4] = {q.x, q.y, q.z, q.w};
  uint32_t o[4];
  const float* sp = s + ((long long)n * C8 + c8) * 8;
  const float* bp = b + c8 * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 u = H16<BF16>::unpack(in[k]);
    o[k] = H16<BF16>::pack(fmaf(gelu_f(u.x), __ldg(sp + 2 * k), __ldg(bp + 2 * k)),
                           fmaf(gelu_f(u.y), __ldg(sp + 2 * k + 1), __ldg(bp + 2 * k + 1)));
  }
  y[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// dh = (dy * s[n,c] + gelu(h) * t[n,c]) * gelu'(h);  dbias[c] += sum_rows dh  (fc1 bias gradient)
// thread = channel pair, block = chunk of rows of one sample (same decomposition as grn_reduce).
template <bool BF16>
__global__ void __launch_bounds__(128)
grn_apply_bwd_kernel(const uint32_t* __restrict__ h, const uint32_t* __restrict__ dy,
                     const float* __restrict__ s, const float* __restrict__ t,
                     uint32_t* __restrict__ dh, float* __restrict__ dbias, int R, int C2,
                     int rows_per_block) {
  const int cp = blockIdx.x * 128 + threadIdx.x;
  if (cp >= C2) return;
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  const long long base = (long long)n * R;
  const int C = 2 * C2;
  const float2 sv = *reinterpret_cast<const float2*>(s + (long long)n * C + 2 * cp);
  float2 tv = make_float2(0.f, 0.f);
  if (t != nullptr) tv = *reinterpret_cast<const float2*>(t + (long long)n * C + 2 * cp);
  float2 bs = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const long long o = (base + r) * C2 + cp;
    const float2 u = H16<BF16>::unpack(__ldg(h + o));
    const float2 d = H16<BF16>::unpack(__ldg(dy + o));
    const float ox = (d.x * sv.x + gelu_f(u.x) * tv.x) * dgelu_f(u.x);
    const float oy = (d.y * sv.y + gelu_f(u.y) * tv.y) * dgelu_f(u.y);
    const uint32_t packed = H16<BF16>::pack(ox, oy);
    dh[o] = packed;
    const float2 rq = H16<BF16>::unpack(packed);  // sum what the wgrad GEMM will actually see
    bs.x += rq.x;
    bs.y += rq.y;
  }
  if (dbias != nullptr) {
    atomicAdd(dbias + 2 * cp, bs.x);
    atomicAdd(dbias + 2 * cp + 1, bs.y);
  }
}

// out[c] += sum_rows x[r][c]
template <bool BF16>
__global__ void __launch_bounds__(128)
colsum_kernel(const uint32_t* __restrict__ x, float* __restrict__ out, long long M, int C2,
              int rows_per_block) {
  const int cp = blockIdx.x * 128 + threadIdx.x;
  if (cp >= C2) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float2 a = make_float2(0.f, 0.f);
#pragma unroll 4
  for (long long r = r0; r < r1; ++r) {
    const float2 v = H16<BF16>::unpack(__ldg(x + r * C2 + cp));
    a.x += v.x;
    a.y += v.y;
  }
  atomicAdd(out + 2 * cp, a.x);
  atomicAdd(out + 2 * cp + 1, a.y);
}

// ---- 16-byte vectorised column reductions (block shape: see ColRedShape in common.cuh)
// MODE 0: out[n,c] += sum x ; MODE 1: out[n,c] += sum x^2
template <bool BF16, int MODE>
__global__ void __launch_bounds__(512)
colreduce8_kernel(const uint4* __restrict__ x, float* __restrict__ out, float* __restrict__ out2, int R, int C8,
                  int rows_per_block, int cw_log2, long long ld8, const float* __restrict__ pivot) {
  constexpr int NQ = MODE == 2 ? 2 : 1;  // MODE 2: sums and sums of squares in one pass (BatchNorm statistics)
  __shared__ float red[NQ * 8 * 512];
  const int cx = threadIdx.x & ((1 << cw_log2) - 1), ry = threadIdx.x >> cw_log2, RL = 512 >> cw_log2;
  const int c8 = (blockIdx.x << cw_log2) + cx;
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  float a[NQ][8];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[q][k] = 0.f;
  if (c8 < C8) {
    const uint4* xp = x + ((long long)n * R) * ld8 + c8;
    // optional per-channel pivot subtracted before summing: one-pass variance E[(x-p)^2] - E[x-p]^2 without the
    // catastrophic cancellation of E[x^2] - E[x]^2 when |mean| >> std (BatchNorm statistics; p = running mean)
    float pv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pv[k] = pivot != nullptr ? __ldg(pivot + c8 * 8 + k) : 0.f;
#pragma unroll 8
    for (int r = r0 + ry; r < r1; r += RL) {
      const uint4 q = __ldg(xp + (long long)r * ld8);
      const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f = H16<BF16>::unpack(w4[k]);
        f.x -= pv[2 * k];
        f.y -= pv[2 * k + 1];
        if (MODE == 0 || MODE == 2) {
          a[0][2 * k] += f.x;
          a[0][2 * k + 1] += f.y;
        }
        if (MODE == 1 || MODE == 2) {
          a[NQ - 1][2 * k] = fmaf(f.x, f.x, a[NQ - 1][2 * k]);
          a[NQ - 1][2 * k + 1] = fmaf(f.y, f.y, a[NQ - 1][2 * k + 1]);
        }
      }
    }
  }
  if constexpr (NQ == 2) {
    float* const outs[2] = {out + (long long)n * C8 * 8, out2 + (long long)n * C8 * 8};
    colred_combine<2>(a, red, cw_log2, blockIdx.x << cw_log2, C8, outs);
  } else {
    float* const outs[1] = {out + (long long)n * C8 * 8};
    colred_combine<1>(a, red, cw_log2, blockIdx.x << cw_log2, C8, outs);
  }
}

// per-sample GRN-scaled fc2 weights: out[n][j][k] = W2[j][k] * s[n][k]; thread = 8 consecutive k of one row j, all n
template <bool BF16>
__global__ void __launch_bounds__(256)
grn_pack_w2_kernel(const float* __restrict__ W2, const float* __restrict__ s, uint4* __restrict__ out,
                   int nb, int C, int C48) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * C48) return;
  const int k8 = idx % C48;
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(W2) + 2 * idx);
  const float4 w1 = __ldg(reinterpret_cast<const float4*>(W2) + 2 * idx + 1);
  for (int n = 0; n < nb; ++n) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(s) + ((long long)n * C48 + k8) * 2);
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(s) + ((long long)n * C48 + k8) * 2 + 1);
    out[(long long)n * C * C48 + idx] =
        make_uint4(H16<BF16>::pack(w0.x * s0.x, w0.y * s0.y), H16<BF16>::pack(w0.z * s0.z, w0.w * s0.w),
                   H16<BF16>::pack(w1.x * s1.x, w1.y * s1.y), H16<BF16>::pack(w1.z * s1.z, w1.w * s1.w));
  }
}

// Per-sample scaled fc2 weights out[n][j][k] = W2[j][k] * s[n][k] (16-bit) and b2eff[j] = b2[j] + sum_k W2[j][k]*bgrn[k]
// in one launch: one block per GRN_JB rows of W2, thread = 8 consecutive k.
constexpr int GRN_JB = 2;
template <bool BF16>
__global__ void __launch_bounds__(256)
grn_prepare_kernel(const float* __restrict__ s, const float* __restrict__ gb, const float* __restrict__ W2,
                   const float* __restrict__ b2, uint4* __restrict__ w2s, float* __restrict__ b2e, int nb, int C,
                   int C4) {
  __shared__ float scratch[32];
  const int C48 = C4 / 8;
  for (int jj = 0; jj < GRN_JB; ++jj) {
    const int j = blockIdx.x * GRN_JB + jj;
    if (j >= C) break;
    float bacc = 0.f;
    for (int k8 = threadIdx.x; k8 < C48; k8 += blockDim.x) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W2 + (long long)j * C4) + 2 * k8);
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W2 + (long long)j * C4) + 2 * k8 + 1);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gb) + 2 * k8);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gb) + 2 * k8 + 1);
      bacc += w0.x * g0.x + w0.y * g0.y + w0.z * g0.z + w0.w * g0.w + w1.x * g1.x + w1.y * g1.y + w1.z * g1.z + w1.w * g1.w;
      for (int n = 0; n < nb; ++n) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(s + (long long)n * C4) + 2 * k8);
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(s + (long long)n * C4) + 2 * k8 + 1);
        w2s[((long long)n * C + j) * C48 + k8] =
            make_uint4(H16<BF16>::pack(w0.x * s0.x, w0.y * s0.y), H16<BF16>::pack(w0.z * s0.z, w0.w * s0.w),
                       H16<BF16>::pack(w1.x * s1.x, w1.y * s1.y), H16<BF16>::pack(w1.z * s1.z, w1.w * s1.w));
      }
    }
    const float tot = block_sum(bacc, scratch);
    if (threadIdx.x == 0) b2e[j] = b2[j] + tot;
  }
}

// GRN forward coefficients, per-sample scaled fc2 weights and the effective bias in ONE launch:
//   s[n][k] = 1 + gw[k] * Gx[n][k] / (mean_k Gx[n][:] + eps), Gx = sqrt(sumsq)      (written by block 0)
//   w2s[n][j][k] = W2[j][k] * s[n][k] (16-bit),  b2e[j] = b2[j] + sum_k W2[j][k] * gb[k]
// Every block recomputes the nb per-sample means (nb * C4 square roots: far cheaper than a kernel boundary on the
// latency-bound small stages); one block per GRN_JB rows of W2, thread = 8 consecutive k.
template <bool BF16, int NB>
__global__ void __launch_bounds__(256)
grn_prepare2_kernel(const float* __restrict__ sumsq, const float* __restrict__ gw, const float* __restrict__ gb,
                    const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ s_out,
                    uint4* __restrict__ w2s, float* __restrict__ b2e, int nb, int C, int C4, float eps) {
  __shared__ float wsum[8][NB];
  __shared__ float inv_m[NB];
  __shared__ float scratch[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float part[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) part[n] = 0.f;
  for (int k = threadIdx.x; k < C4; k += 256) {
#pragma unroll
    for (int n = 0; n < NB; ++n)
      if (n < nb) part[n] += sqrtf(__ldg(sumsq + (long long)n * C4 + k));
  }
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    const float v = warp_sum(part[n]);
    if (lane == 0) wsum[warp][n] = v;
  }
  __syncthreads();
  if (threadIdx.x < NB) {
    float m = 0.f;
    for (int w = 0; w < 8; ++w) m += wsum[w][threadIdx.x];
    inv_m[threadIdx.x] = 1.0f / (m / C4 + eps);
  }
  __syncthreads();
  const int C48 = C4 / 8;
  for (int jj = 0; jj < GRN_JB; ++jj) {
    const int j = blockIdx.x * GRN_JB + jj;
    if (j >= C) break;
    float bacc = 0.f;
    for (int k8 = threadIdx.x; k8 < C48; k8 += blockDim.x) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W2 + (long long)j * C4) + 2 * k8);
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W2 + (long long)j * C4) + 2 * k8 + 1);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gb) + 2 * k8);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gb) + 2 * k8 + 1);
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(gw) + 2 * k8);
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(gw) + 2 * k8 + 1);
      bacc += w0.x * g0.x + w0.y * g0.y + w0.z * g0.z + w0.w * g0.w + w1.x * g1.x + w1.y * g1.y + w1.z * g1.z + w1.w * g1.w;
      for (int n = 0; n < nb; ++n) {
        const float4 u0 = __ldg(reinterpret_cast<const float4*>(sumsq + (long long)n * C4) + 2 * k8);
        const float4 u1 = __ldg(reinterpret_cast<const float4*>(sumsq + (long long)n * C4) + 2 * k8 + 1);
        const float im = inv_m[n];
        float4 s0, s1;
        s0.x = fmaf(q0.x * sqrtf(u0.x), im, 1.0f); s0.y = fmaf(q0.y * sqrtf(u0.y), im, 1.0f);
        s0.z = fmaf(q0.z * sqrtf(u0.z), im, 1.0f); s0.w = fmaf(q0.w * sqrtf(u0.w), im, 1.0f);
        s1.x = fmaf(q1.x * sqrtf(u1.x), im, 1.0f); s1.y = fmaf(q1.y * sqrtf(u1.y), im, 1.0f);
        s1.z = fmaf(q1.z * sqrtf(u1.z), im, 1.0f); s1.w = fmaf(q1.w * sqrtf(u1.w), im, 1.0f);
        if (blockIdx.x == 0 && jj == 0) {
          reinterpret_cast<float4*>(s_out + (long long)n * C4)[2 * k8] = s0;
          reinterpret_cast<float4*>(s_out + (long long)n * C4)[2 * k8 + 1] = s1;
        }
        w2s[((long long)n * C + j) * C48 + k8] =
            make_uint4(H16<BF16>::pack(w0.x * s0.x, w0.y * s0.y), H16<BF16>::pack(w0.z * s0.z, w0.w * s0.w),
                       H16<BF16>::pack(w1.x * s1.x, w1.y * s1.y), H16<BF16>::pack(w1.z * s1.z, w1.w * s1.w));
      }
    }
    const float tot = block_sum(bacc, scratch);
    if (threadIdx.x == 0) b2e[j] = b2[j] + tot;
  }
}

// b2eff[j] = b2[j] + sum_k W2[j][k] * bgrn[k]   (one warp per output row)
__global__ void __launch_bounds__(256)
grn_bias_eff_kernel(const float* __restrict__ W2, const float* __restrict__ bgrn, const float* __restrict__ b2,
                    float* __restrict__ out, int C, int C4) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= C) return;
  const int lane = threadIdx.x & 31;
  float a = 0.f;
  for (int k = lane; k < C4; k += 32) a = fmaf(W2[(long long)j * C4 + k], bgrn[k], a);
  a = warp_sum(a);
  if (lane == 0) out[j] = b2[j] + a;
}

// From the per-sample wgrad partials P[n][j][k] = sum_{rows of n} dout[r][j] * g[r][k]:
//   dW2[j][k]  = sum_n s[n][k] * P[n][j][k] + bgrn[k] * db2[j]
//   S1[n][k]  += sum_j W2[j][k] * P[n][j][k]         (= sum_r dy * g with dy = dout W2)
//   dbgrn[k]  += sum_j W2[j][k] * db2[j]             (= sum_r dy)
// thread = column k, block = chunk of rows j; NB <= 16 samples per launch.
// ldp = row pitch of P (>= C4).  db2_in == NULL: the fc2 bias gradient is read from column C4 of P (the ones column of
// the GELU output buffer: P[n][j][C4] = sum_{rows of n} dout[r][j]) and written to db2_out.
template <int NB>
__global__ void __launch_bounds__(128)
grn_wgrad_finish_kernel(const float* __restrict__ P, const float* __restrict__ W2, const float* __restrict__ s,
                        const float* __restrict__ bgrn, const float* __restrict__ db2_in, float* __restrict__ dW2,
                        float* __restrict__ S1, float* __restrict__ dbgrn, float* __restrict__ db2_out, int nb, int C,
                        int C4, long long ldp, int jchunk) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= C4) return;
  const int j0 = blockIdx.y * jchunk, j1 = min(C, j0 + jchunk);
  float sv[NB], s1[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    sv[n] = n < nb ? s[(long long)n * C4 + k] : 0.f;
    s1[n] = 0.f;
  }
  const float bg = bgrn[k];
  float dbg = 0.f;
  const long long per = (long long)C * ldp;
#pragma unroll 4
  for (int j = j0; j < j1; ++j) {
    const float w = __ldg(W2 + (long long)j * C4 + k);
    float d = 0.f;
    if (db2_in != nullptr) {
      d = db2_in[j];
    } else {
#pragma unroll
      for (int n = 0; n < NB; ++n)
        if (n < nb) d += __ldg(P + n * per + (long long)j * ldp + C4);
      if (k == 0) db2_out[j] = d;
    }
    float acc = bg * d;
    dbg = fmaf(w, d, dbg);
#pragma unroll
    for (int n = 0; n < NB; ++n)
      if (n < nb) {
        const float pv = __ldg(P + n * per + (long long)j * ldp + k);
        acc = fmaf(sv[n], pv, acc);
        s1[n] = fmaf(w, pv, s1[n]);
      }
    dW2[(long long)j * C4 + k] = acc;
  }
#pragma unroll
  for (int n = 0; n < NB; ++n)
    if (n < nb) atomicAdd(S1 + (long long)n * C4 + k, s1[n]);
  atomicAdd(dbgrn + k, dbg);
}

// four columns per thread (16-byte loads of P, W2, s; 16-byte stores of dW2): C4 % 4 == 0, ldp % 4 == 0, nb <= 8.
// Block = 32 column quads x 4 row groups: the four row groups of a block combine their S1 / dbgrn partial sums through
// shared memory, so a block issues 36 x 32 atomics instead of 36 x 128 (the atomics, not the 16-byte loads, were the cost).
__global__ void __launch_bounds__(128)
grn_wgrad_finish4_kernel(const float* __restrict__ P, const float* __restrict__ W2, const float* __restrict__ s,
                         const float* __restrict__ bgrn, const float* __restrict__ db2_in, float* __restrict__ dW2,
                         float* __restrict__ S1, float* __restrict__ dbgrn, float* __restrict__ db2_out, int nb, int C, int C4,
                         long long ldp, int jchunk) {
  constexpr int NB = 8;
  __shared__ float4 red[3][9][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k = (blockIdx.x * 32 + tx) * 4;
  const bool active = k < C4;
  const int j0 = blockIdx.y * jchunk, j1 = min(C, j0 + jchunk);
  float4 sv[NB], s1[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    sv[n] = (active && n < nb) ? __ldg(reinterpret_cast<const float4*>(s + (long long)n * C4 + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s1[n] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 bg = active ? __ldg(reinterpret_cast<const float4*>(bgrn + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dbg = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long per = (long long)C * ldp;
  if (active) {
#pragma unroll 2
    for (int j = j0 + ty; j < j1; j += 4) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(W2 + (long long)j * C4 + k));
      float4 pv[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n)
        pv[n] = n < nb ? __ldg(reinterpret_cast<const float4*>(P + n * per + (long long)j * ldp + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float d = 0.f;
      if (db2_in != nullptr) {
        d = db2_in[j];
      } else {
#pragma unroll
        for (int n = 0; n < NB; ++n)
          if (n < nb) d += __ldg(P + n * per + (long long)j * ldp + C4);
        if (k == 0) db2_out[j] = d;
      }
      float4 acc = make_float4(bg.x * d, bg.y * d, bg.z * d, bg.w * d);
      dbg.x = fmaf(w.x, d, dbg.x); dbg.y = fmaf(w.y, d, dbg.y); dbg.z = fmaf(w.z, d, dbg.z); dbg.w = fmaf(w.w, d, dbg.w);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        acc.x = fmaf(sv[n].x, pv[n].x, acc.x); acc.y = fmaf(sv[n].y, pv[n].y, acc.y);
        acc.z = fmaf(sv[n].z, pv[n].z, acc.z); acc.w = fmaf(sv[n].w, pv[n].w, acc.w);
        s1[n].x = fmaf(w.x, pv[n].x, s1[n].x); s1[n].y = fmaf(w.y, pv[n].y, s1[n].y);
        s1[n].z = fmaf(w.z, pv[n].z, s1[n].z); s1[n].w = fmaf(w.w, pv[n].w, s1[n].w);
      }
      *reinterpret_cast<float4*>(dW2 + (long long)j * C4 + k) = acc;
    }
  }
  if (ty > 0) {
#pragma unroll
    for (int n = 0; n < NB; ++n) red[ty - 1][n][tx] = s1[n];
    red[ty - 1][8][tx] = dbg;
  }
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float4 v = red[g][n][tx];
        s1[n].x += v.x; s1[n].y += v.y; s1[n].z += v.z; s1[n].w += v.w;
      }
      const float4 v = red[g][8][tx];
      dbg.x += v.x; dbg.y += v.y; dbg.z += v.z; dbg.w += v.w;
    }
#pragma unroll
    for (int n = 0; n < NB; ++n)
      if (n < nb) {
        float* o = S1 + (long long)n * C4 + k;
        atomicAdd(o, s1[n].x); atomicAdd(o + 1, s1[n].y); atomicAdd(o + 2, s1[n].z); atomicAdd(o + 3, s1[n].w);
      }
    atomicAdd(dbgrn + k, dbg.x); atomicAdd(dbgrn + k + 1, dbg.y); atomicAdd(dbgrn + k + 2, dbg.z); atomicAdd(dbgrn + k + 3, dbg.w);
  }
}

static int rows_per_block_for(long long rows, int col_blocks, int samples) {
  // aim for ~8 blocks per SM in total, at least 16 rows per block
  const long long target_blocks = 148LL * 8;
  long long per = (rows * col_blocks * samples + target_blocks - 1) / target_blocks;
  if (per < 16) per = 16;
  if (per > rows) per = rows;
  return (int)per;
}

}  // namespace vb

using namespace vb;

#define DISPATCH_DT(dtype, ...)                                              \
  do {                                                                       \
    if ((dtype) == VB200_BF16) { constexpr bool BF = true; __VA_ARGS__; }    \
    else if ((dtype) == VB200_FP16) { constexpr bool BF = false; __VA_ARGS__; } \
    else return vb::fail(VB200_ERR_UNSUPPORTED, "dtype %d", (int)(dtype));   \
  } while (0)

template <bool BF, int TH, int TW, bool SW>
static void dwconv7_launch(const void* x, const float* wt, const float* bias, const void* add, void* y, int B, int H,
                           int W, int C2, cudaStream_t st) {
  static PerDeviceOnce once;
  const int dev = PerDeviceOnce::device();
  if (once.need(dev)) {
    cudaFuncSetAttribute(dwconv7_kernel<BF, TH, TW, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM);
    once.done(dev);
  }
  dim3 grid((unsigned)(((W + TW - 1) / TW) * ((H + TH - 1) / TH) * B), (unsigned)((C2 + 127) / 128));
  dwconv7_kernel<BF, TH, TW, SW><<<grid, 128, SW ? DW_SMEM : 0, st>>>((const uint32_t*)x, wt, bias, (const uint32_t*)add,
                                                        (uint32_t*)y, B, H, W, C2);
}

// register-tile kernels for channel counts the shared-memory tile kernels (dwconv_sm100.cu) do not take (C % 8 != 0)
int vb::dwconv7_legacy(const void* x, const float* wt, const float* bias, const void* add, void* y, int B, int H, int W,
                       int C, int dtype, cudaStream_t st) {
  VB_SUPPORTED(C % 2 == 0 && (long long)B * H * W * C < (1LL << 31), "C (%d) must be even, tensor < 2^31 elements", C);
  const int C2 = C / 2;
  // small feature maps: 2x4-pixel tiles so that enough blocks exist to fill the SMs
  const long long big_blocks = (long long)((W + 7) / 8) * ((H + 3) / 4) * B * ((C2 + 127) / 128);
  if (big_blocks >= 2 * 148) {
    DISPATCH_DT(dtype, (dwconv7_launch<BF, 4, 8, true>(x, wt, bias, add, y, B, H, W, C2, st)));
  } else {
    DISPATCH_DT(dtype, (dwconv7_launch<BF, 2, 4, false>(x, wt, bias, add, y, B, H, W, C2, st)));
  }
  return check_launch("vb200_dwconv7");
}

int vb::dwconv7_wgrad_legacy(const void* x, const void* dy, float* dwt, float* db, int B, int H, int W, int C,
                             int dtype, cudaStream_t st) {
  VB_SUPPORTED(C % 2 == 0 && (long long)B * H * W * C < (1LL << 31), "C (%d) must be even, tensor < 2^31 elements", C);
  const int C2 = C / 2;
  const int colb = (C2 + 127) / 128;
  int rpb = (int)(((long long)H * B * colb + 148 * 4 - 1) / (148 * 4));
  rpb = (rpb + 1) & ~1;  // the kernel walks two rows at a time
  if (rpb < 2) rpb = 2;
  if (rpb > H) rpb = H;
  const int strips = (H + rpb - 1) / rpb;
  dim3 grid((unsigned)(B * strips), (unsigned)colb);
  DISPATCH_DT(dtype, dwconv7_wgrad_kernel<BF><<<grid, 128, 0, st>>>((const uint32_t*)x, (const uint32_t*)dy, dwt,
                                                                   db, B, H, W, C2, rpb));
  return check_launch("vb200_dwconv7_wgrad");
}

template <bool BF, int NV>
static void ln_fwd_rows_launch(const void* x, const float* gamma, const float* beta, void* y, long long ldy8, float* mean,
                               float* rstd, int64_t M, int C8, float eps, int ones, void* ones2, long long ld2_8, int c2_8,
                               cudaStream_t st) {
  layernorm_fwd_rows_kernel<BF, NV><<<(unsigned)((M + 7) / 8), 256, 0, st>>>(
      (const uint4*)x, gamma, beta, (uint4*)y, ldy8, mean, rstd, M, C8, eps, ones, (uint4*)ones2, ld2_8, c2_8);
}

/* y row pitch ldy (elements, % 8); ones = n > 0: y[:, C:C+8n] = {1,0,...,0} (needs ldy >= C + 8 n); ones2 != NULL: the
 * same columns written at ones2[row * ld2 + col2 ..) */
extern "C" int vb200_layernorm_fwd_ld(const void* x, const float* gamma, const float* beta, void* y, int64_t ldy, int ones,
                                      void* ones2, int64_t ld2, int col2, float* mean, float* rstd, int64_t M, int C,
                                      float eps, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && gamma && beta && y && mean && rstd, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool plain = ldy == C && !ones && ones2 == nullptr;
  if (C % 8 == 0 && C <= 2048 && ldy % 8 == 0 && ld2 % 8 == 0 && col2 % 8 == 0) {
    VB_REQUIRE(ones >= 0 && ones <= 8 && ldy >= C + 8 * ones, "ldy (%lld) too small", (long long)ldy);
    const int C8 = C / 8, nv = (C8 + 31) / 32;
#define LN_FWD(NV) DISPATCH_DT(dtype, (ln_fwd_rows_launch<BF, NV>(x, gamma, beta, y, ldy / 8, mean, rstd, M, C8, eps, ones, ones2, ld2 / 8, col2 / 8, st)))
    if (nv <= 1) LN_FWD(1);
    else if (nv <= 2) LN_FWD(2);
    else if (nv <= 3) LN_FWD(3);
    else if (nv <= 4) LN_FWD(4);
    else LN_FWD(8);
#undef LN_FWD
    return check_launch("vb200_layernorm_fwd");
  }
  VB_SUPPORTED(plain, "layernorm_fwd: pitch / ones columns need C %% 8 == 0 and C <= 2048 (C=%d)", C);
  VB_SUPPORTED(C % 2 == 0, "C (%d) must be even", C);
  const unsigned grid = (unsigned)((M + 7) / 8);
  DISPATCH_DT(dtype, layernorm_fwd_kernel<BF><<<grid, 256, 0, st>>>((const uint32_t*)x, gamma, beta, (uint32_t*)y,
                                                                   mean, rstd, M, C / 2, eps));
  return check_launch("vb200_layernorm_fwd");
}
extern "C" int vb200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                                   float* mean, float* rstd, int64_t M, int C, float eps, int dtype,
                                   vb200_stream_t stream) {
  return vb200_layernorm_fwd_ld(x, gamma, beta, y, C, 0, nullptr, 8, 0, mean, rstd, M, C, eps, dtype, stream);
}

template <bool BF, int NP>
static void ln_bwd_launch(const void* dy, const void* x, const float* mean, const float* rstd,
                          const float* gamma, void* dx, float* dgamma, float* dbeta, int64_t M, int C,
                          cudaStream_t st) {
  long long blocks = (M + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  layernorm_bwd_kernel<BF, NP><<<(unsigned)blocks, 256, 2 * C * sizeof(float), st>>>(
      (const uint32_t*)dy, (const uint32_t*)x, mean, rstd, gamma, (uint32_t*)dx, dgamma, dbeta, M, C / 2);
}

template <bool BF, int NV>
static void ln_bwd_rows_launch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                               void* dx, int64_t M, int C8, cudaStream_t st) {
  layernorm_bwd_rows_kernel<BF, NV><<<(unsigned)((M + 7) / 8), 256, 0, st>>>((const uint4*)dy, (const uint4*)x, mean, rstd,
                                                                             gamma, (uint4*)dx, M, C8);
}

extern "C" int vb200_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                                   const float* gamma, void* dx, float* dgamma, float* dbeta, int64_t M,
                                   int C, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta, "null pointer");
  VB_SUPPORTED(C % 2 == 0 && C <= 3072, "C (%d) must be even and <= 3072", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 8 == 0 && C <= 768) {  // up to 3 vectors per lane stay in registers
    const int C8 = C / 8, nv = (C8 + 31) / 32;
    long long blocks = (M + 7) / 8;
    const long long cap = 148LL * 2;  // two resident 256-thread blocks per SM; rows beyond that by grid stride
    if (blocks > cap) blocks = cap;
    const size_t smem = 2 * (size_t)C * sizeof(float);
#define LN_FUSED(NV)                                                                                               \
  DISPATCH_DT(dtype, (layernorm_bwd_fused_kernel<BF, NV><<<(unsigned)blocks, 256, smem, st>>>(                     \
                         (const uint4*)dy, (const uint4*)x, mean, rstd, gamma, (uint4*)dx, dgamma, dbeta, M, C8)))
    if (nv <= 1) LN_FUSED(1);
    else if (nv <= 2) LN_FUSED(2);
    else LN_FUSED(3);
#undef LN_FUSED
    return check_launch("vb200_layernorm_bwd(fused)");
  }
  if (C % 8 == 0 && C <= 2048) {
    const int C8 = C / 8, nv = (C8 + 31) / 32;
#define LN_ROWS(NV) DISPATCH_DT(dtype, (ln_bwd_rows_launch<BF, NV>(dy, x, mean, rstd, gamma, dx, M, C8, st)))
    if (nv <= 1) LN_ROWS(1);
    else if (nv <= 2) LN_ROWS(2);
    else if (nv <= 3) LN_ROWS(3);
    else if (nv <= 4) LN_ROWS(4);
    else LN_ROWS(8);
#undef LN_ROWS
    if (int rc = check_launch("vb200_layernorm_bwd(rows)")) return rc;
    const ColRedShape sh = ColRedShape::make(C8);
    long long rpb = (M * sh.colb + 148 * 2 - 1) / (148 * 2);
    const long long min_rows = 4LL * (512 >> sh.cw_log2);
    if (rpb < min_rows) rpb = min_rows;
    if (rpb > M) rpb = M;
    dim3 grid(sh.colb, (unsigned)((M + rpb - 1) / rpb));
    DISPATCH_DT(dtype, layernorm_bwd_cols_kernel<BF><<<grid, 512, 0, st>>>((const uint4*)dy, (const uint4*)x, mean, rstd,
                                                                          dgamma, dbeta, M, C8, (int)rpb, sh.cw_log2));
    return check_launch("vb200_layernorm_bwd(cols)");
  }
  const int np = (C / 2 + 31) / 32;
#define LN_BWD(NP) DISPATCH_DT(dtype, (ln_bwd_launch<BF, NP>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, C, st)))
  if (np <= 3) LN_BWD(3);
  else if (np <= 6) LN_BWD(6);
  else if (np <= 12) LN_BWD(12);
  else if (np <= 24) LN_BWD(24);
  else LN_BWD(48);
#undef LN_BWD
  return check_launch("vb200_layernorm_bwd");
}

extern "C" int vb200_grn_sumsq(const void* h, float* sumsq, int B, int R, int C, int dtype,
                               vb200_stream_t stream) {
  VB_REQUIRE(h && sumsq, "null pointer");
  VB_SUPPORTED(C % 2 == 0, "C even");
  const int C2 = C / 2, colb = (C2 + 127) / 128;
  const int rpb = rows_per_block_for(R, colb, B);
  dim3 grid(colb, (R + rpb - 1) / rpb, B);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, grn_reduce_kernel<BF, 0><<<grid, 128, 0, st>>>((const uint32_t*)h, nullptr, sumsq, nullptr, R,
                                                                   C2, rpb));
  return check_launch("vb200_grn_sumsq");
}

extern "C" int vb200_grn_coef_fwd(const float* sumsq, const float* w, float* s, int B, int C, float eps,
                                  vb200_stream_t stream) {
  VB_REQUIRE(sumsq && w && s, "null pointer");
  grn_coef_fwd_kernel<<<dim3(B, (C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sumsq, w, s, C, eps);
  return check_launch("vb200_grn_coef_fwd");
}

extern "C" int vb200_grn_apply_fwd(const void* h, const float* s, const float* b, void* y, int B, int R,
                                   int C, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(h && s && b && y, "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total8 = (long long)B * R * C / 8;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, grn_apply_fwd_kernel<BF><<<(unsigned)((total8 + 255) / 256), 256, 0, st>>>(
                         (const uint4*)h, s, b, (uint4*)y, R, C / 8, total8));
  return check_launch("vb200_grn_apply_fwd");
}

extern "C" int vb200_grn_bwd_reduce(const void* h, const void* dy, float* S1, float* sdy, int B, int R,
                                    int C, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(h && dy && S1 && sdy, "null pointer");
  VB_SUPPORTED(C % 2 == 0, "C even");
  const int C2 = C / 2, colb = (C2 + 127) / 128;
  const int rpb = rows_per_block_for(R, colb, B);
  dim3 grid(colb, (R + rpb - 1) / rpb, B);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, grn_reduce_kernel<BF, 1><<<grid, 128, 0, st>>>((const uint32_t*)h, (const uint32_t*)dy, S1, sdy,
                                                                   R, C2, rpb));
  return check_launch("vb200_grn_bwd_reduce");
}

extern "C" int vb200_grn_coef_bwd(const float* sumsq, const float* S1, const float* w, float* t, float* dw,
                                  int B, int C, float eps, vb200_stream_t stream) {
  VB_REQUIRE(sumsq && S1 && w && t && dw, "null pointer");
  grn_coef_bwd_kernel<<<dim3(B, (C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sumsq, S1, w, t, dw, C, eps);
  return check_launch("vb200_grn_coef_bwd");
}

extern "C" int vb200_grn_apply_bwd(const void* h, const void* dy, const float* s, const float* t, void* dh,
                                   float* dbias, int B, int R, int C, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(h && dy && s && dh, "null pointer");
  VB_SUPPORTED(C % 2 == 0, "C even");
  const int C2 = C / 2, colb = (C2 + 127) / 128;
  const int rpb = rows_per_block_for(R, colb, B);
  dim3 grid(colb, (R + rpb - 1) / rpb, B);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, grn_apply_bwd_kernel<BF><<<grid, 128, 0, st>>>((const uint32_t*)h, (const uint32_t*)dy, s, t,
                                                                   (uint32_t*)dh, dbias, R, C2, rpb));
  return check_launch("vb200_grn_apply_bwd");
}

extern "C" int vb200_colsum(const void* x, float* out, int64_t M, int C, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && out, "null pointer");
  VB_SUPPORTED(C % 2 == 0, "C even");
  const int C2 = C / 2, colb = (C2 + 127) / 128;
  const int rpb = rows_per_block_for(M, colb, 1);
  dim3 grid(colb, (unsigned)((M + rpb - 1) / rpb));
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, colsum_kernel<BF><<<grid, 128, 0, st>>>((const uint32_t*)x, out, M, C2, rpb));
  return check_launch("vb200_colsum");
}

/* MODE 0: out[n,c] += sum_r x[n,r,c];  MODE 1: out[n,c] += sum_r x[n,r,c]^2;  MODE 2: both in one pass, sums into
 * out[0:B*C], sums of squares into out[B*C:2*B*C]   (x [B,R,C] with row pitch ld >= C, C % 8 == 0, out pre-zeroed) */
extern "C" int vb200_colreduce_ld(const void* x, float* out, int B, int64_t R, int C, int64_t ld, int mode,
                                  const float* pivot, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && out, "null pointer");
  VB_SUPPORTED(C % 8 == 0 && ld % 8 == 0 && ld >= C && R < (1LL << 31), "C (%d) %% 8, ld (%lld) %% 8", C, (long long)ld);
  const int C8 = C / 8;
  const long long ld8 = ld / 8;
  const ColRedShape sh = ColRedShape::make(C8);
  // ONE wave of 2 blocks of 512 threads per SM: the row split is rounded DOWN so that colb * row blocks * B never exceeds the
  // resident slots (13 row blocks x 24 gave 312 blocks on 296 slots = a second, almost empty wave); at least 4 rows per lane
  long long ny = (148 * 2) / ((long long)sh.colb * B);
  if (ny < 1) ny = 1;
  long long rpb = (R + ny - 1) / ny;
  const long long min_rows = 4LL * (512 >> sh.cw_log2);
  if (rpb < min_rows) rpb = min_rows;
  if (rpb > R) rpb = R;
  dim3 grid(sh.colb, (unsigned)((R + rpb - 1) / rpb), B);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0)
    DISPATCH_DT(dtype, colreduce8_kernel<BF, 0><<<grid, 512, 0, st>>>((const uint4*)x, out, nullptr, (int)R, C8, (int)rpb, sh.cw_log2, ld8, pivot));
  else if (mode == 1)
    DISPATCH_DT(dtype, colreduce8_kernel<BF, 1><<<grid, 512, 0, st>>>((const uint4*)x, out, nullptr, (int)R, C8, (int)rpb, sh.cw_log2, ld8, pivot));
  else if (mode == 2)
    DISPATCH_DT(dtype, colreduce8_kernel<BF, 2><<<grid, 512, 0, st>>>((const uint4*)x, out, out + (long long)B * C, (int)R, C8, (int)rpb, sh.cw_log2, ld8, pivot));
  else
    return fail(VB200_ERR_INVALID, "colreduce mode %d", mode);
  return check_launch("vb200_colreduce");
}
extern "C" int vb200_colreduce(const void* x, float* out, int B, int64_t R, int C, int mode, int dtype,
                               vb200_stream_t stream) {
  return vb200_colreduce_ld(x, out, B, R, C, C, mode, nullptr, dtype, stream);
}

extern "C" int vb200_grn_pack_w2(const float* W2, const float* s, void* out, int nb, int C, int C4, int dtype,
                                 vb200_stream_t stream) {
  VB_REQUIRE(W2 && s && out, "null pointer");
  VB_SUPPORTED(C4 % 8 == 0, "C4 (%d) %% 8", C4);
  const long long total = (long long)C * (C4 / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, grn_pack_w2_kernel<BF><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W2, s, (uint4*)out, nb, C,
                                                                                           C4 / 8));
  return check_launch("vb200_grn_pack_w2");
}

extern "C" int vb200_grn_bias_eff(const float* W2, const float* bgrn, const float* b2, float* out, int C, int C4,
                                  vb200_stream_t stream) {
  VB_REQUIRE(W2 && bgrn && b2 && out, "null pointer");
  grn_bias_eff_kernel<<<(C + 7) / 8, 256, 0, (cudaStream_t)stream>>>(W2, bgrn, b2, out, C, C4);
  return check_launch("vb200_grn_bias_eff");
}

extern "C" int vb200_grn_wgrad_finish_ld(const float* P, int64_t ldp, const float* W2, const float* s,
                                         const float* bgrn, const float* db2_in, float* dW2, float* S1, float* dbgrn,
                                         float* db2_out, int nb, int C, int C4, vb200_stream_t stream) {
  VB_REQUIRE(P && W2 && s && bgrn && dW2 && S1 && dbgrn, "null pointer");
  VB_REQUIRE(db2_in != nullptr || (db2_out != nullptr && ldp >= C4 + 1), "db2 from the ones column needs db2_out and ldp > C4");
  VB_SUPPORTED(nb <= 16, "at most 16 samples per call (nb=%d)", nb);
  cudaStream_t st = (cudaStream_t)stream;
  if (nb <= 8 && C4 % 4 == 0 && ldp % 4 == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(W2) & 15) == 0 && (reinterpret_cast<uintptr_t>(dW2) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(s) & 15) == 0 && (reinterpret_cast<uintptr_t>(bgrn) & 15) == 0) {
    const int colb4 = (C4 / 4 + 31) / 32;  // 32 column quads per block, its four warps take rows j0 + {0..3} + 4 i
    int jc = (C * colb4 + 148 * 4 - 1) / (148 * 4);
    jc = (jc + 3) / 4 * 4;
    if (jc < 16) jc = 16;
    dim3 grid4(colb4, (C + jc - 1) / jc);
    grn_wgrad_finish4_kernel<<<grid4, 128, 0, st>>>(P, W2, s, bgrn, db2_in, dW2, S1, dbgrn, db2_out, nb, C, C4, ldp, jc);
    return check_launch("vb200_grn_wgrad_finish");
  }
  const int colb = (C4 + 127) / 128;
  int jchunk = (C * colb + 148 * 4 - 1) / (148 * 4);
  if (jchunk < 8) jchunk = 8;
  dim3 grid(colb, (C + jchunk - 1) / jchunk);
  if (nb <= 8)
    grn_wgrad_finish_kernel<8><<<grid, 128, 0, st>>>(P, W2, s, bgrn, db2_in, dW2, S1, dbgrn, db2_out, nb, C, C4, ldp, jchunk);
  else
    grn_wgrad_finish_kernel<16><<<grid, 128, 0, st>>>(P, W2, s, bgrn, db2_in, dW2, S1, dbgrn, db2_out, nb, C, C4, ldp, jchunk);
  return check_launch("vb200_grn_wgrad_finish");
}
extern "C" int vb200_grn_wgrad_finish(const float* P, const float* W2, const float* s, const float* bgrn,
                                      const float* db2, float* dW2, float* S1, float* dbgrn, int nb, int C, int C4,
                                      vb200_stream_t stream) {
  VB_REQUIRE(db2 != nullptr, "null pointer");
  return vb200_grn_wgrad_finish_ld(P, C4, W2, s, bgrn, db2, dW2, S1, dbgrn, nullptr, nb, C, C4, stream);
}

extern "C" int vb200_grn_prepare2(const float* sumsq, const float* gw, const float* gb, const float* W2, const float* b2,
                                  float* s_out, void* w2s, float* b2e, int nb, int C, int C4, float eps, int dtype,
                                  vb200_stream_t stream) {
  VB_REQUIRE(sumsq && gw && gb && W2 && b2 && s_out && w2s && b2e, "null pointer");
  VB_SUPPORTED(C4 % 8 == 0 && nb <= 16, "grn_prepare2: C4 (%d) %% 8, nb (%d) <= 16", C4, nb);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((C + GRN_JB - 1) / GRN_JB);
  if (nb <= 8)
    DISPATCH_DT(dtype, (grn_prepare2_kernel<BF, 8><<<grid, 256, 0, st>>>(sumsq, gw, gb, W2, b2, s_out, (uint4*)w2s, b2e, nb, C, C4, eps)));
  else
    DISPATCH_DT(dtype, (grn_prepare2_kernel<BF, 16><<<grid, 256, 0, st>>>(sumsq, gw, gb, W2, b2, s_out, (uint4*)w2s, b2e, nb, C, C4, eps)));
  return check_launch("vb200_grn_prepare2");
}

extern "C" int vb200_grn_prepare(const float* s, const float* gb, const float* W2, const float* b2, void* w2s,
                                 float* b2e, int nb, int C, int C4, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(s && gb && W2 && b2 && w2s && b2e, "null pointer");
  VB_SUPPORTED(C4 % 8 == 0, "grn_prepare: C4 (%d) %% 8", C4);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((C + GRN_JB - 1) / GRN_JB);
  DISPATCH_DT(dtype, grn_prepare_kernel<BF><<<grid, 256, 0, st>>>(s, gb, W2, b2, (uint4*)w2s, b2e, nb, C, C4));
  return check_launch("vb200_grn_prepare");
}

// ---- ConvNeXt-V1 layer scale, backward side (timm ConvNeXtBlock: out = gamma * (y W2^T + b2) + x) in ONE launch:
//   G [C, ldg] fp32 = dout^T [y | 1] (the weight-gradient GEMM without the layer scale; its ones column, when present, is
//   passed separately as db_raw), W2 [C, C4], gamma, b2 [C]  ->
//     dW2 = G * gamma[:, None],  db2 = db_raw * gamma,  dgamma = sum_j G * W2 + db_raw * b2   (dgamma pre-zeroed),
//     w2t16 [C4, C] = 16-bit (W2 * gamma / max|gamma|)^T  (the fc2 data-gradient operand; the epilogue multiplies by sv),
//     sv [C4] = max|gamma|.
// Tile = 32 output channels x 64 hidden columns, transposed through shared memory so that both W2 reads and w2t16 writes
// are row-contiguous.  Replaces ~12 torch elementwise / reduction launches per block.
namespace vb {
template <bool BF16>
__global__ void __launch_bounds__(256)
layerscale_bwd_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ W2, const float* __restrict__ gamma,
                      const float* __restrict__ b2, const float* __restrict__ db_raw, uint16_t* __restrict__ w2t,
                      float* __restrict__ dgamma, float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ sv,
                      int C, int C4) {
  __shared__ float gmax_s;
  __shared__ float tile[32][65];
  __shared__ float part[32][9];
  // max |gamma| (C <= a few thousand fp32 values out of L2: every block recomputes it)
  float m = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(gamma + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) part[0][threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t = fmaxf(t, part[0][w]);
    gmax_s = fmaxf(t, 1e-30f);
  }
  __syncthreads();
  const float gmax = gmax_s;
  const int c0 = blockIdx.y * 32, j0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 columns x 4 row lanes
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 4 * i, j = j0 + tx;
    float g = 0.f, w = 0.f, gam = 0.f;
    if (c < C && j < C4) {
      g = __ldg(G + (long long)c * ldg + j);
      w = __ldg(W2 + (long long)c * C4 + j);
      gam = __ldg(gamma + c);
      dW2[(long long)c * C4 + j] = g * gam;
    }
    acc[i] = g * w;
    tile[ty + 4 * i][tx] = w * (gam / gmax);
  }
  // dgamma: sum over the tile's 64 columns (two warps per row lane)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if ((threadIdx.x & 31) == 0) {
      const int c = c0 + ty + 4 * i;
      if (c < C) atomicAdd(dgamma + c, acc[i]);
    }
  }
  __syncthreads();
  // transposed 16-bit store: w2t[j, c0 .. c0 + 31]
  for (int e = threadIdx.x; e < 64 * 32; e += blockDim.x) {
    const int jj = e >> 5, cc = e & 31;
    const int j = j0 + jj, c = c0 + cc;
    if (j < C4 && c < C) {
      const typename H16<BF16>::T v = H16<BF16>::from_f(tile[cc][jj]);
      w2t[(long long)j * C + c] = *reinterpret_cast<const uint16_t*>(&v);
    }
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
      const int c = c0 + i;
      if (c < C) {
        const float dr = db_raw ? __ldg(db_raw + c) : 0.f, gam = __ldg(gamma + c);
        if (db2) db2[c] = dr * gam;
        if (db_raw) atomicAdd(dgamma + c, dr * __ldg(b2 + c));
      }
    }
  }
  if (blockIdx.y == 0) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x)
      if (j0 + i < C4) sv[j0 + i] = gmax;
  }
}
}  // namespace vb

extern "C" int vb200_layerscale_bwd(const float* G, int64_t ldg, const float* W2, const float* gamma, const float* b2,
                                    const float* db_raw, void* w2t, float* dgamma, float* dW2, float* db2, float* sv, int C,
                                    int C4, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(G && W2 && gamma && b2 && w2t && dgamma && dW2 && sv, "null pointer");
  VB_REQUIRE(C > 0 && C4 > 0 && ldg >= C4, "shape %d x %d (ld %lld)", C, C4, (long long)ldg);
  dim3 grid((C4 + 63) / 64, (C + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, layerscale_bwd_kernel<BF><<<grid, 256, 0, st>>>(G, ldg, W2, gamma, b2, db_raw, (uint16_t*)w2t, dgamma,
                                                                     dW2, db2, sv, C, C4));
  return check_launch("vb200_layerscale_bwd");
}
