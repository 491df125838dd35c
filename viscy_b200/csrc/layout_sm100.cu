// Data-movement kernels on channels-last 16-bit activations: pixel-shuffle + skip concat (fwd/bwd),
// 2x2 patchify for the k2s2 downsample convs (fwd/bwd), stem patchify (NCDHW -> GEMM rows),
// generic 3-D im2col / col2im for 3x3x3-style convolutions (v1 lowering of the 3-D convs).
// Reference semantics: monai SubpixelUpsample(pre_conv=None) + torch.cat (VM/components/blocks.py:137-172),
// timm ConvNeXtStage.downsample Conv2d(k=2,s=2), UNeXt2Stem / StemDepthtoChannels (VM/components/stems.py:26-50,117-134).
#include "common.cuh"

namespace vb {

// out[n, 2h+i, 2w+j, c'] = prev[n,h,w, c'*4 + i*2 + j]  (c' < Cq = Cp/4)   else skip[n,2h+i,2w+j,c'-Cq]
__global__ void __launch_bounds__(256)
pixshuf_cat_fwd_kernel(const uint16_t* __restrict__ prev, const uint16_t* __restrict__ skip,
                       uint16_t* __restrict__ out, int h, int w, int Cp, int Cs, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Cq = Cp / 4, Co = Cq + Cs;
  const int c = (int)(idx % Co);
  long long pix = idx / Co;  // (n, Y, X) of the 2h x 2w grid
  const int X = (int)(pix % (2 * w));
  const long long t = pix / (2 * w);
  const int Y = (int)(t % (2 * h));
  const long long n = t / (2 * h);
  uint16_t v;
  if (c < Cq)
    v = prev[((n * h + (Y >> 1)) * w + (X >> 1)) * Cp + c * 4 + (Y & 1) * 2 + (X & 1)];
  else
    v = skip[pix * Cs + (c - Cq)];
  out[idx] = v;
}

__global__ void __launch_bounds__(256)
pixshuf_cat_bwd_kernel(const uint16_t* __restrict__ dout, uint16_t* __restrict__ dprev,
                       uint16_t* __restrict__ dskip, int h, int w, int Cp, int Cs, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Cq = Cp / 4, Co = Cq + Cs;
  const int c = (int)(idx % Co);
  long long pix = idx / Co;
  const int X = (int)(pix % (2 * w));
  const long long t = pix / (2 * w);
  const int Y = (int)(t % (2 * h));
  const long long n = t / (2 * h);
  const uint16_t v = dout[idx];
  if (c < Cq)
    dprev[((n * h + (Y >> 1)) * w + (X >> 1)) * Cp + c * 4 + (Y & 1) * 2 + (X & 1)] = v;
  else
    dskip[pix * Cs + (c - Cq)] = v;
}

// out[(n,oh,ow), (kh,kw,c)] = x[n, 2oh+kh, 2ow+kw, c]; 8 channels (16 B) per thread. inverse=1 scatters back.
__global__ void __launch_bounds__(256)
patchify2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int W, int C8,
                 long long total, int inverse) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx enumerates the patchified tensor [(n,oh,ow)][kh][kw][c8]
  const int c8 = (int)(idx % C8);
  long long t = idx / C8;
  const int kw = (int)(t & 1);
  t >>= 1;
  const int kh = (int)(t & 1);
  t >>= 1;
  const int ow = (int)(t % (W / 2));
  t /= (W / 2);
  const int oh = (int)(t % (H / 2));
  const long long n = t / (H / 2);
  const long long xi = ((n * H + 2 * oh + kh) * W + 2 * ow + kw) * C8 + c8;
  if (!inverse)
    dst[idx] = src[xi];
  else
    dst[xi] = src[idx];
}

// Stem rows: A[(n,oh,ow), k] with k = ((c*D + z)*kH + kh)*kW + kw  <- x[n,c,z, oh*kH+kh, ow*kW+kw]
// (the whole Z extent goes into K; the block-diagonal weight built on the host side reproduces
//  Conv3d(kernel == stride) followed by the reference's (B,C,D,H,W)->(B,C*D,H,W) reshape.)
template <typename TIN, bool BF16>
__global__ void __launch_bounds__(256)
stem_patchify_kernel(const TIN* __restrict__ x, uint16_t* __restrict__ A, int Cin, int D, int H, int W,
                     int kH, int kW, int Kpad, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx enumerates (n, c, z, ih, ow) ; each thread moves kW contiguous input values
  const int OW = W / kW, OH = H / kH;
  const int ow = (int)(idx % OW);
  long long t = idx / OW;
  const int ih = (int)(t % H);
  t /= H;
  const int z = (int)(t % D);
  t /= D;
  const int c = (int)(t % Cin);
  const long long n = t / Cin;
  const int oh = ih / kH, kh = ih % kH;
  const TIN* src = x + (((n * Cin + c) * D + z) * H + ih) * (long long)W + (long long)ow * kW;
  uint16_t* dst = A + ((n * OH + oh) * OW + ow) * (long long)Kpad + ((c * D + z) * kH + kh) * kW;
  for (int j = 0; j < kW; ++j) {
    const float v = static_cast<float>(src[j]);
    typename H16<BF16>::T hv = H16<BF16>::from_f(v);
    dst[j] = *reinterpret_cast<uint16_t*>(&hv);
  }
}

// Tiled form of the same lowering: the block of one (sample, patch row) reads its Cin*D*kH input rows (W contiguous values
// each) into shared memory as 16-bit, then writes the OW matrix rows [K] with 16-byte stores - both sides coalesced.  (The
// element-wise kernel above writes kW * 2 bytes per thread into OW different matrix rows: 0.5 TB/s on the 385 MB input of
// the contrastive stem.)
template <typename TIN, bool BF16>
__global__ void __launch_bounds__(256)
stem_patchify_tile_kernel(const TIN* __restrict__ x, uint16_t* __restrict__ A, int Cin, int D, int H, int W, int kH, int kW,
                          int Kpad) {
  extern __shared__ uint16_t slab[];  // [Cin*D*kH][W]
  const int OH = H / kH, OW = W / kW;
  const int oh = blockIdx.x % OH;
  const long long n = blockIdx.x / OH;
  const int R = Cin * D * kH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const bool vec4 = (W & 3) == 0;  // rows are 16-byte (fp32) / 8-byte (16-bit) aligned: four values per load
  for (int r = warp; r < R; r += nwarp) {  // one input row per warp and step: divisions per row, not per element
    const int kh = r % kH, cz = r / kH;    // cz = c * D + z
    const TIN* src = x + ((n * Cin * D + cz) * H + oh * kH + kh) * (long long)W;
    if (vec4) {
      for (int w4 = lane; w4 < W / 4; w4 += 32) {
        float v[4];
        if constexpr (sizeof(TIN) == 4) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(src) + w4);
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(src) + w4);
          const float2 a = H16<BF16>::unpack(u.x), c = H16<BF16>::unpack(u.y);
          v[0] = a.x; v[1] = a.y; v[2] = c.x; v[3] = c.y;
        }
        *reinterpret_cast<uint2*>(slab + r * W + w4 * 4) = make_uint2(H16<BF16>::pack(v[0], v[1]), H16<BF16>::pack(v[2], v[3]));
      }
    } else {
      for (int w = lane; w < W; w += 32) {
        typename H16<BF16>::T hv = H16<BF16>::from_f(static_cast<float>(src[w]));
        slab[r * W + w] = *reinterpret_cast<uint16_t*>(&hv);
      }
    }
  }
  __syncthreads();
  const int K = R * kW, K8 = Kpad / 8;
  uint4* out = reinterpret_cast<uint4*>(A + ((n * OH + oh) * (long long)OW) * Kpad);
  if (kW == 4 && vec4 && K == Kpad) {
    // 8 consecutive K entries = the 4-wide patch rows r0 = 2 k8, r0 + 1 at this patch column: two 8-byte shared loads
    for (int i = threadIdx.x; i < OW * K8; i += blockDim.x) {
      const int ow = i / K8, k8 = i - ow * K8;
      const uint2 lo = *reinterpret_cast<const uint2*>(slab + (2 * k8) * W + ow * 4);
      const uint2 hi = *reinterpret_cast<const uint2*>(slab + (2 * k8 + 1) * W + ow * 4);
      out[(long long)ow * K8 + k8] = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
    return;
  }
  for (int i = threadIdx.x; i < OW * K8; i += blockDim.x) {
    const int ow = i / K8, k8 = i - ow * K8;
    uint32_t q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t pair = 0;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int k = k8 * 8 + e * 2 + h2;
        const int r = k / kW, j = k - r * kW;
        const uint32_t v = k < K ? slab[r * W + ow * kW + j] : 0u;
        pair |= v << (16 * h2);
      }
      q[e] = pair;
    }
    out[(long long)ow * K8 + k8] = make_uint4(q[0], q[1], q[2], q[3]);
  }
}

// col[(n,oz,oy,ox), (kd,kh,kw,c)] = u[n, oz*sd+kd-pd, oy*sh+kh-ph, ox*sw+kw-pw, c]  (0 outside); 8 ch / thread
struct Conv3dGeom {
  int N, D, H, W, C;        // input NDHWC
  int OD, OH, OW;           // output extent
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
};

__global__ void __launch_bounds__(256)
im2col3d_kernel(const uint4* __restrict__ u, uint4* __restrict__ col, Conv3dGeom g, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int C8 = g.C / 8;
  const int c8 = (int)(idx % C8);
  long long t = idx / C8;
  const int taps = g.kd * g.kh * g.kw;
  const int tap = (int)(t % taps);
  long long row = t / taps;
  const int kw = tap % g.kw, kh = (tap / g.kw) % g.kh, kd = tap / (g.kw * g.kh);
  const int ox = (int)(row % g.OW);
  long long r = row / g.OW;
  const int oy = (int)(r % g.OH);
  r /= g.OH;
  const int oz = (int)(r % g.OD);
  const long long n = r / g.OD;
  const int iz = oz * g.sd + kd - g.pd, iy = oy * g.sh + kh - g.ph, ix = ox * g.sw + kw - g.pw;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (iz >= 0 && iz < g.D && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
    v = __ldg(u + (((n * g.D + iz) * g.H + iy) * (long long)g.W + ix) * C8 + c8);
  col[idx] = v;
}

// du[n,z,y,x,c] = sum over taps of dcol[(n,oz,oy,ox),(tap,c)] with oz*sd + kd - pd == z etc. (gather, no atomics)
template <bool BF16>
__global__ void __launch_bounds__(256)
col2im3d_kernel(const uint4* __restrict__ dcol, uint4* __restrict__ du, Conv3dGeom g, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int C8 = g.C / 8;
  const int c8 = (int)(idx % C8);
  long long t = idx / C8;
  const int x = (int)(t % g.W);
  t /= g.W;
  const int y = (int)(t % g.H);
  t /= g.H;
  const int z = (int)(t % g.D);
  const long long n = t / g.D;
  const int taps = g.kd * g.kh * g.kw;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int kd = 0; kd < g.kd; ++kd) {
    const int zz = z + g.pd - kd;
    if (zz < 0 || zz % g.sd != 0) continue;
    const int oz = zz / g.sd;
    if (oz >= g.OD) continue;
    for (int kh = 0; kh < g.kh; ++kh) {
      const int yy = y + g.ph - kh;
      if (yy < 0 || yy % g.sh != 0) continue;
      const int oy = yy / g.sh;
      if (oy >= g.OH) continue;
      for (int kw = 0; kw < g.kw; ++kw) {
        const int xx = x + g.pw - kw;
        if (xx < 0 || xx % g.sw != 0) continue;
        const int ox = xx / g.sw;
        if (ox >= g.OW) continue;
        const long long row = ((n * g.OD + oz) * g.OH + oy) * (long long)g.OW + ox;
        const int tap = (kd * g.kh + kh) * g.kw + kw;
        const uint4 q = __ldg(dcol + (row * taps + tap) * C8 + c8);
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = H16<BF16>::unpack(w4[k]);
          acc[2 * k] += f.x;
          acc[2 * k + 1] += f.y;
        }
      }
    }
  }
  du[idx] = make_uint4(H16<BF16>::pack(acc[0], acc[1]), H16<BF16>::pack(acc[2], acc[3]),
                       H16<BF16>::pack(acc[4], acc[5]), H16<BF16>::pack(acc[6], acc[7]));
}

// fp32 -> 16-bit cast (weight packing), optional transpose of a [R, Cc] matrix to [Cc, R]
template <bool BF16>
__global__ void __launch_bounds__(256)
cast_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long R, long long Cc,
            int transpose) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * Cc) return;
  long long s = idx;
  if (transpose) {
    const long long c = idx / R, r = idx % R;  // dst is [Cc, R]
    s = r * Cc + c;
  }
  typename H16<BF16>::T hv = H16<BF16>::from_f(src[s]);
  dst[idx] = *reinterpret_cast<uint16_t*>(&hv);
}

// conv_dw.weight [C][49] fp32 -> tap-major [49][C] and its spatially flipped copy (dgrad taps), one launch
__global__ void __launch_bounds__(256)
dw_pack_kernel(const float* __restrict__ w, float* __restrict__ wt, float* __restrict__ wtf, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 49 * C) return;
  const int tap = idx / C, c = idx % C;
  const float v = w[c * 49 + tap];
  wt[idx] = v;
  wtf[(48 - tap) * C + c] = v;
}

static inline unsigned blocks_for(long long total, int bs = 256) { return (unsigned)((total + bs - 1) / bs); }

}  // namespace vb

using namespace vb;

extern "C" int vb200_pixshuf_cat_fwd(const void* prev, const void* skip, void* out, int B, int h, int w,
                                     int Cp, int Cs, vb200_stream_t stream) {
  VB_REQUIRE(prev && out && (skip || Cs == 0), "null pointer");
  VB_REQUIRE(Cp % 4 == 0, "pixel shuffle needs Cp %% 4 == 0 (Cp=%d)", Cp);
  const long long total = (long long)B * 4 * h * w * (Cp / 4 + Cs);
  pixshuf_cat_fwd_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)prev, (const uint16_t*)skip, (uint16_t*)out, h, w, Cp, Cs, total);
  return check_launch("vb200_pixshuf_cat_fwd");
}

extern "C" int vb200_pixshuf_cat_bwd(const void* dout, void* dprev, void* dskip, int B, int h, int w, int Cp,
                                     int Cs, vb200_stream_t stream) {
  VB_REQUIRE(dout && dprev && (dskip || Cs == 0), "null pointer");
  VB_REQUIRE(Cp % 4 == 0, "pixel shuffle needs Cp %% 4 == 0 (Cp=%d)", Cp);
  const long long total = (long long)B * 4 * h * w * (Cp / 4 + Cs);
  pixshuf_cat_bwd_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)dout, (uint16_t*)dprev, (uint16_t*)dskip, h, w, Cp, Cs, total);
  return check_launch("vb200_pixshuf_cat_bwd");
}

extern "C" int vb200_patchify2(const void* src, void* dst, int B, int H, int W, int C, int inverse,
                               vb200_stream_t stream) {
  VB_REQUIRE(src && dst, "null pointer");
  VB_SUPPORTED(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "patchify2 needs C %% 8 == 0 and even H, W");
  const long long total = (long long)B * H * W * (C / 8);
  patchify2_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, H, W,
                                                                        C / 8, total, inverse);
  return check_launch("vb200_patchify2");
}

extern "C" int vb200_stem_patchify(const void* x, int x_dtype /*0 bf16, 1 fp16, 2 fp32*/, void* A, int B,
                                   int Cin, int D, int H, int W, int kH, int kW, int Kpad, int dtype,
                                   vb200_stream_t stream) {
  VB_REQUIRE(x && A, "null pointer");
  VB_REQUIRE(H % kH == 0 && W % kW == 0 && Kpad >= Cin * D * kH * kW, "bad stem geometry");
  const long long total = (long long)B * Cin * D * H * (W / kW);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = blocks_for(total);
  const size_t slab = (size_t)Cin * D * kH * W * 2;
  const bool tiled = slab <= 200 * 1024 && Kpad % 8 == 0 && (long long)B * (H / kH) < (1LL << 31);
  const unsigned tgrid = (unsigned)(B * (H / kH));
#define STEM(TIN, BF)                                                                                                   \
  do {                                                                                                                  \
    if (tiled) {                                                                                                        \
      static PerDeviceOnce once;                                                                                        \
      const int dev = PerDeviceOnce::device();                                                                          \
      if (once.need(dev)) {                                                                                             \
        cudaFuncSetAttribute(stem_patchify_tile_kernel<TIN, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
        once.done(dev);                                                                                                 \
      }                                                                                                                 \
      stem_patchify_tile_kernel<TIN, BF><<<tgrid, 256, slab, st>>>((const TIN*)x, (uint16_t*)A, Cin, D, H, W, kH, kW, Kpad); \
    } else {                                                                                                            \
      stem_patchify_kernel<TIN, BF><<<grid, 256, 0, st>>>((const TIN*)x, (uint16_t*)A, Cin, D, H, W, kH, kW, Kpad, total); \
    }                                                                                                                   \
  } while (0)
  if (dtype == VB200_BF16) {
    if (x_dtype == 2) STEM(float, true);
    else if (x_dtype == 0) STEM(__nv_bfloat16, true);
    else return fail(VB200_ERR_UNSUPPORTED, "stem input dtype %d for bf16 path", x_dtype);
  } else if (dtype == VB200_FP16) {
    if (x_dtype == 2) STEM(float, false);
    else if (x_dtype == 1) STEM(__half, false);
    else return fail(VB200_ERR_UNSUPPORTED, "stem input dtype %d for fp16 path", x_dtype);
  } else {
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  }
#undef STEM
  return check_launch("vb200_stem_patchify");
}

static int geom_from(const int32_t* g, Conv3dGeom* o) {
  o->N = g[0]; o->D = g[1]; o->H = g[2]; o->W = g[3]; o->C = g[4];
  o->kd = g[5]; o->kh = g[6]; o->kw = g[7];
  o->sd = g[8]; o->sh = g[9]; o->sw = g[10];
  o->pd = g[11]; o->ph = g[12]; o->pw = g[13];
  o->OD = g[14]; o->OH = g[15]; o->OW = g[16];
  if (o->C % 8 != 0) return fail(VB200_ERR_UNSUPPORTED, "conv3d lowering needs C %% 8 == 0 (C=%d)", o->C);
  return VB200_OK;
}

/* geom = {N,D,H,W,C, kd,kh,kw, sd,sh,sw, pd,ph,pw, OD,OH,OW} */
extern "C" int vb200_im2col3d(const void* u, void* col, const int32_t* geom, vb200_stream_t stream) {
  VB_REQUIRE(u && col && geom, "null pointer");
  Conv3dGeom g;
  if (int rc = geom_from(geom, &g)) return rc;
  const long long total = (long long)g.N * g.OD * g.OH * g.OW * g.kd * g.kh * g.kw * (g.C / 8);
  im2col3d_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)u, (uint4*)col, g, total);
  return check_launch("vb200_im2col3d");
}

extern "C" int vb200_col2im3d(const void* dcol, void* du, const int32_t* geom, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(dcol && du && geom, "null pointer");
  Conv3dGeom g;
  if (int rc = geom_from(geom, &g)) return rc;
  const long long total = (long long)g.N * g.D * g.H * g.W * (g.C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB200_BF16)
    col2im3d_kernel<true><<<blocks_for(total), 256, 0, st>>>((const uint4*)dcol, (uint4*)du, g, total);
  else if (dtype == VB200_FP16)
    col2im3d_kernel<false><<<blocks_for(total), 256, 0, st>>>((const uint4*)dcol, (uint4*)du, g, total);
  else
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  return check_launch("vb200_col2im3d");
}

extern "C" int vb200_cast_pack(const float* src, void* dst, int64_t R, int64_t Cc, int transpose, int dtype,
                               vb200_stream_t stream) {
  VB_REQUIRE(src && dst, "null pointer");
  const long long total = R * Cc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB200_BF16)
    cast_kernel<true><<<blocks_for(total), 256, 0, st>>>(src, (uint16_t*)dst, R, Cc, transpose);
  else if (dtype == VB200_FP16)
    cast_kernel<false><<<blocks_for(total), 256, 0, st>>>(src, (uint16_t*)dst, R, Cc, transpose);
  else
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  return check_launch("vb200_cast_pack");
}

// nn.Conv3d weight [Co, Ci, KD, KH, KW] fp32 -> 16-bit GEMM rows in one launch (was permute + pad + contiguous + cast,
// plus a flip for the data gradient: 3-4 ATen launches per convolution and step).
//   flipped == 0: out [cout_pad][(kd, kh, kw, cin_pad)]            = W[co][ci][kd][kh][kw]       (forward / wgrad order)
//   flipped != 0: out [cin_pad][(kd, kh, kw, cout_pad)]            = W[co][ci][KD-1-kd][KH-1-kh][KW-1-kw]  (data gradient)
// One block per (output row, chunk of 32 inner channels): the 32 x T source floats are read in runs of T (contiguous per
// channel pair), transposed through shared memory, and written as 64-byte runs per tap; padding rows / channels are zeros.
namespace vb {
template <bool BF16>
__global__ void __launch_bounds__(128)
conv_weight_rows_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Co, int Ci, int KD, int KH, int KW,
                        int cin_pad, int cout_pad, int flipped) {
  __shared__ float tile[32][28];
  const int T = KD * KH * KW;
  const int n_outer = flipped ? Ci : Co, n_inner = flipped ? Co : Ci, inner_pad = flipped ? cout_pad : cin_pad;
  const int o = blockIdx.x, i0 = blockIdx.y * 32;
  const long long so = flipped ? (long long)T : (long long)Ci * T, si = flipped ? (long long)Ci * T : (long long)T;
  for (int idx = threadIdx.x; idx < 32 * T; idx += 128) {
    const int i = idx / T, t = idx - i * T;
    float v = 0.f;
    if (o < n_outer && i0 + i < n_inner) v = __ldg(w + o * so + (i0 + i) * si + t);
    tile[i][t] = v;
  }
  __syncthreads();
  using H = typename H16<BF16>::T;
  H* orow = reinterpret_cast<H*>(out) + (long long)o * T * inner_pad;
  for (int idx = threadIdx.x; idx < 32 * T; idx += 128) {
    const int t = idx >> 5, i = idx & 31;
    if (i0 + i >= inner_pad) continue;
    int ts = t;
    if (flipped) {
      const int kw = t % KW, kh = (t / KW) % KH, kd = t / (KW * KH);
      ts = ((KD - 1 - kd) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw);
    }
    orow[(long long)t * inner_pad + i0 + i] = H16<BF16>::from_f(tile[i][ts]);
  }
}
}  // namespace vb

extern "C" int vb200_conv_weight_rows(const float* w, void* out, int Co, int Ci, int KD, int KH, int KW, int cin_pad,
                                      int cout_pad, int flipped, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(w && out, "null pointer");
  VB_REQUIRE(Co > 0 && Ci > 0 && cin_pad >= Ci && cout_pad >= Co, "padded channel counts (%d, %d) below (%d, %d)", cin_pad, cout_pad, Ci, Co);
  VB_SUPPORTED(KD * KH * KW <= 27 && KD > 0 && KH > 0 && KW > 0, "up to 27 taps (%d x %d x %d)", KD, KH, KW);
  const int rows = flipped ? cin_pad : cout_pad, inner_pad = flipped ? cout_pad : cin_pad;
  dim3 grid(rows, (inner_pad + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB200_BF16)
    vb::conv_weight_rows_kernel<true><<<grid, 128, 0, st>>>(w, (uint16_t*)out, Co, Ci, KD, KH, KW, cin_pad, cout_pad, flipped);
  else if (dtype == VB200_FP16)
    vb::conv_weight_rows_kernel<false><<<grid, 128, 0, st>>>(w, (uint16_t*)out, Co, Ci, KD, KH, KW, cin_pad, cout_pad, flipped);
  else
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  return check_launch("vb200_conv_weight_rows");
}

extern "C" int vb200_dw_pack(const float* w, float* wt, float* wtf, int C, vb200_stream_t stream) {
  VB_REQUIRE(w && wt && wtf, "null pointer");
  dw_pack_kernel<<<blocks_for(49LL * C), 256, 0, (cudaStream_t)stream>>>(w, wt, wtf, C);
  return check_launch("vb200_dw_pack");
}

// ------------------------------------------------------------------------------ small-row helpers (projection head)
namespace vb {

// BatchNorm1d over rows of x [B, C] (16-bit): one thread per column.
// training: batch statistics (biased variance for normalisation), mean / rstd / unbiased var written out;
// eval: use the provided running mean / var.  Optional fused ReLU.
template <bool BF16>
__global__ void __launch_bounds__(128)
bn_rows_fwd_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ run_mean, const float* __restrict__ run_var, uint16_t* __restrict__ y,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ var_unbiased,
                   int B, int C, float eps, int training, int relu) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  using T = typename H16<BF16>::T;
  const T* xp = reinterpret_cast<const T*>(x);
  float m, var;
  if (training) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += H16<BF16>::to_f(xp[(long long)b * C + c]);
    m = s / B;
    float q = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = H16<BF16>::to_f(xp[(long long)b * C + c]) - m;
      q = fmaf(d, d, q);
    }
    var = q / B;
    if (var_unbiased != nullptr) var_unbiased[c] = B > 1 ? q / (B - 1) : q;
  } else {
    m = run_mean[c];
    var = run_var[c];
  }
  const float rs = rsqrtf(var + eps);
  mean_out[c] = m;
  rstd_out[c] = rs;
  const float g = gamma[c], bt = beta[c];
  T* yp = reinterpret_cast<T*>(y);
  for (int b = 0; b < B; ++b) {
    float v = (H16<BF16>::to_f(xp[(long long)b * C + c]) - m) * rs * g + bt;
    if (relu) v = fmaxf(v, 0.f);
    yp[(long long)b * C + c] = H16<BF16>::from_f(v);
  }
}

// backward of the above (y is the forward output, used for the ReLU mask)
template <bool BF16>
__global__ void __launch_bounds__(128)
bn_rows_bwd_kernel(const uint16_t* __restrict__ dy, const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                   const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                   uint16_t* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C,
                   int training, int relu) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  using T = typename H16<BF16>::T;
  const T* dyp = reinterpret_cast<const T*>(dy);
  const T* xp = reinterpret_cast<const T*>(x);
  const T* yp = reinterpret_cast<const T*>(y);
  const float m = mean[c], rs = rstd[c], g = gamma[c];
  float s1 = 0.f, s2 = 0.f;
  for (int b = 0; b < B; ++b) {
    float d = H16<BF16>::to_f(dyp[(long long)b * C + c]);
    if (relu && !(H16<BF16>::to_f(yp[(long long)b * C + c]) > 0.f)) d = 0.f;
    const float xh = (H16<BF16>::to_f(xp[(long long)b * C + c]) - m) * rs;
    s1 += d;
    s2 = fmaf(d, xh, s2);
  }
  dgamma[c] = s2;
  dbeta[c] = s1;
  T* dxp = reinterpret_cast<T*>(dx);
  for (int b = 0; b < B; ++b) {
    float d = H16<BF16>::to_f(dyp[(long long)b * C + c]);
    if (relu && !(H16<BF16>::to_f(yp[(long long)b * C + c]) > 0.f)) d = 0.f;
    const float xh = (H16<BF16>::to_f(xp[(long long)b * C + c]) - m) * rs;
    const float v = training ? g * rs * (d - s1 / B - xh * s2 / B) : g * rs * d;
    dxp[(long long)b * C + c] = H16<BF16>::from_f(v);
  }
}

// out[b, r, c] = src[b, c] * scale   (gradient of the global average pool), 8 channels per thread
template <bool BF16>
__global__ void __launch_bounds__(256)
bcast_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ out, int R, int C8, float scale, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  const long long b = idx / ((long long)R * C8);
  const uint4 q = __ldg(src + b * C8 + c8);
  const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = H16<BF16>::unpack(w4[k]);
    o[k] = H16<BF16>::pack(f.x * scale, f.y * scale);
  }
  out[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace vb

extern "C" int vb200_bn_rows_fwd(const void* x, const float* gamma, const float* beta, const float* run_mean,
                                 const float* run_var, void* y, float* mean, float* rstd, float* var_unbiased, int B,
                                 int C, float eps, int training, int relu, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(x && gamma && beta && y && mean && rstd && (training || (run_mean && run_var)), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((C + 127) / 128);
  if (dtype == VB200_BF16)
    bn_rows_fwd_kernel<true><<<grid, 128, 0, st>>>((const uint16_t*)x, gamma, beta, run_mean, run_var, (uint16_t*)y, mean, rstd, var_unbiased, B, C, eps, training, relu);
  else if (dtype == VB200_FP16)
    bn_rows_fwd_kernel<false><<<grid, 128, 0, st>>>((const uint16_t*)x, gamma, beta, run_mean, run_var, (uint16_t*)y, mean, rstd, var_unbiased, B, C, eps, training, relu);
  else
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  return check_launch("vb200_bn_rows_fwd");
}

extern "C" int vb200_bn_rows_bwd(const void* dy, const void* x, const void* y, const float* gamma, const float* mean,
                                 const float* rstd, void* dx, float* dgamma, float* dbeta, int B, int C, int training,
                                 int relu, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(dy && x && y && gamma && mean && rstd && dx && dgamma && dbeta, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((C + 127) / 128);
  if (dtype == VB200_BF16)
    bn_rows_bwd_kernel<true><<<grid, 128, 0, st>>>((const uint16_t*)dy, (const uint16_t*)x, (const uint16_t*)y, gamma, mean, rstd, (uint16_t*)dx, dgamma, dbeta, B, C, training, relu);
  else if (dtype == VB200_FP16)
    bn_rows_bwd_kernel<false><<<grid, 128, 0, st>>>((const uint16_t*)dy, (const uint16_t*)x, (const uint16_t*)y, gamma, mean, rstd, (uint16_t*)dx, dgamma, dbeta, B, C, training, relu);
  else
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  return check_launch("vb200_bn_rows_bwd");
}

extern "C" int vb200_bcast_rows(const void* src, void* out, int B, int R, int C, float scale, int dtype,
                                vb200_stream_t stream) {
  VB_REQUIRE(src && out, "null pointer");
  VB_SUPPORTED(C % 8 == 0, "C (%d) %% 8", C);
  const long long total = (long long)B * R * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB200_BF16)
    bcast_rows_kernel<true><<<blocks_for(total), 256, 0, st>>>((const uint4*)src, (uint4*)out, R, C / 8, scale, total);
  else if (dtype == VB200_FP16)
    bcast_rows_kernel<false><<<blocks_for(total), 256, 0, st>>>((const uint4*)src, (uint4*)out, R, C / 8, scale, total);
  else
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  return check_launch("vb200_bcast_rows");
}

// ------------------------------------------------------------------------------ per-step weight packing, one launch
// table[item] = {src, dst, dst2, R, Cc, kind, first_block}:
//   kind 0 / 1 (legacy, 1024 elements per block): cast [R,Cc] / cast + transpose -> [Cc,R];
//   kind 2: depthwise taps src [C=R][49] -> dst fp32 [49][C] and flipped dst2 fp32 [49][C] (1024 elements per block);
//   kind 3: one 64 x 64 tile of src [R,Cc] per block, read once (coalesced rows), written as BOTH 16-bit operand copies:
//           dst [R,Cc] (forward, K-major) and dst2 [Cc,R] (data gradient) through a padded shared-memory transpose.
namespace vb {
constexpr int PM_FIELDS = 7;
constexpr int PM_PER_BLOCK = 1024;
constexpr int PM_TILE = 64;
template <bool BF16>
__global__ void __launch_bounds__(256)
pack_multi_kernel(const long long* __restrict__ table, int n_items) {
  __shared__ int s_item;
  __shared__ float tile[PM_TILE][PM_TILE + 1];
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_items - 1;  // last item whose first_block <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid * PM_FIELDS + 6] <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_item = lo;
  }
  __syncthreads();
  const long long* it = table + s_item * PM_FIELDS;
  const float* src = reinterpret_cast<const float*>(it[0]);
  const long long R = it[3], Cc = it[4];
  const int kind = (int)it[5];
  if (kind == 3) {
    const int tiles_c = (int)((Cc + PM_TILE - 1) / PM_TILE);
    const int tidx = (int)((long long)blockIdx.x - it[6]);
    const long long r0 = (long long)(tidx / tiles_c) * PM_TILE, c0 = (long long)(tidx % tiles_c) * PM_TILE;
    const bool vec = (R % 8 == 0) && (Cc % 8 == 0);
    uint16_t* dn = reinterpret_cast<uint16_t*>(it[1]);
    uint16_t* dt = reinterpret_cast<uint16_t*>(it[2]);
    if (vec) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // 64 rows x 16 float4
        const int id = threadIdx.x + 256 * k;
        const int r = id >> 4, c4 = (id & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < R && c0 + c4 < Cc) v = __ldg(reinterpret_cast<const float4*>(src + (r0 + r) * Cc + c0 + c4));
        tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 2; ++k) {  // 64 rows x 8 groups of 8 columns
        const int id = threadIdx.x + 256 * k;
        const int r = id >> 3, c8 = (id & 7) * 8;
        if (r0 + r < R && c0 + c8 < Cc) {
          const float* t = &tile[r][c8];
          *reinterpret_cast<uint4*>(dn + (r0 + r) * Cc + c0 + c8) =
              make_uint4(H16<BF16>::pack(t[0], t[1]), H16<BF16>::pack(t[2], t[3]), H16<BF16>::pack(t[4], t[5]),
                         H16<BF16>::pack(t[6], t[7]));
        }
        const int c = id >> 3, r8 = (id & 7) * 8;  // transposed copy: row c of dst2 holds 8 consecutive r
        if (c0 + c < Cc && r0 + r8 < R) {
          *reinterpret_cast<uint4*>(dt + (c0 + c) * R + r0 + r8) =
              make_uint4(H16<BF16>::pack(tile[r8][c], tile[r8 + 1][c]), H16<BF16>::pack(tile[r8 + 2][c], tile[r8 + 3][c]),
                         H16<BF16>::pack(tile[r8 + 4][c], tile[r8 + 5][c]), H16<BF16>::pack(tile[r8 + 6][c], tile[r8 + 7][c]));
        }
      }
    } else {
      for (int id = threadIdx.x; id < PM_TILE * PM_TILE; id += 256) {
        const int r = id >> 6, c = id & 63;
        tile[r][c] = (r0 + r < R && c0 + c < Cc) ? __ldg(src + (r0 + r) * Cc + c0 + c) : 0.f;
      }
      __syncthreads();
      for (int id = threadIdx.x; id < PM_TILE * PM_TILE; id += 256) {
        const int r = id >> 6, c = id & 63;
        if (r0 + r < R && c0 + c < Cc) {
          typename H16<BF16>::T hv = H16<BF16>::from_f(tile[r][c]);
          dn[(r0 + r) * Cc + c0 + c] = *reinterpret_cast<uint16_t*>(&hv);
        }
        const int c2 = id >> 6, r2 = id & 63;
        if (r0 + r2 < R && c0 + c2 < Cc) {
          typename H16<BF16>::T hv = H16<BF16>::from_f(tile[r2][c2]);
          dt[(c0 + c2) * R + r0 + r2] = *reinterpret_cast<uint16_t*>(&hv);
        }
      }
    }
    return;
  }
  const long long base = ((long long)blockIdx.x - it[6]) * PM_PER_BLOCK;
  const long long total = kind == 2 ? 49 * R : R * Cc;
#pragma unroll
  for (int k = 0; k < PM_PER_BLOCK / 256; ++k) {
    const long long idx = base + k * 256 + threadIdx.x;
    if (idx >= total) break;
    if (kind == 2) {
      const long long tap = idx / R, c = idx % R;
      const float v = src[c * 49 + tap];
      reinterpret_cast<float*>(it[1])[idx] = v;
      reinterpret_cast<float*>(it[2])[(48 - tap) * R + c] = v;
    } else {
      long long s = idx;
      if (kind == 1) {
        const long long c = idx / R, r = idx % R;
        s = r * Cc + c;
      }
      typename H16<BF16>::T hv = H16<BF16>::from_f(src[s]);
      reinterpret_cast<uint16_t*>(it[1])[idx] = *reinterpret_cast<uint16_t*>(&hv);
    }
  }
}
}  // namespace vb

extern "C" int vb200_pack_multi(const void* table, int n_items, int64_t total_blocks, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(table && n_items > 0 && total_blocks > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB200_BF16)
    pack_multi_kernel<true><<<(unsigned)total_blocks, 256, 0, st>>>((const long long*)table, n_items);
  else if (dtype == VB200_FP16)
    pack_multi_kernel<false><<<(unsigned)total_blocks, 256, 0, st>>>((const long long*)table, n_items);
  else
    return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  return check_launch("vb200_pack_multi");
}
