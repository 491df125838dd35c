// PixelToVoxelHead (VM/components/heads.py:594-641) on channels-last 16-bit activations:
//   pixelshuffle(2) [+ ConstantPad2d((1,0,1,0)) + AvgPool2d(2, stride 1)] + (c//d, d) un-fold  -> u [B,Dz,2h,2w,Cu]
//   (Conv3d k3 pad (0,1,1) runs through the im2col + tcgen05 GEMM path)
//   InstanceNorm3d(affine=False, eps 1e-5) + PReLU + Conv3d(k=1) + transpose + PixelShuffle(2) + transpose -> NCDHW
// and the matching backward kernels.
#include "common.cuh"

namespace vb {

// S[n, cm, 2h+i, 2w+j] = dec[n,h,w, cm*4 + i*2 + j];  P = pool ? avg2x2(pad_left_top(S)) : S
// u[n, dz, Y, X, c] = P[n, c*Dz + dz, Y, X]  (c < Cc; zero for Cc <= c < Cu)
//
// One block = one decoder row h (output rows 2h, 2h+1) x 16 decoder pixels (32 output columns).
// Phase 1: thread = (decoder pixel, cm): four 8-byte loads (the 2x2 sub-pixels of cm at (h-1..h, w-1..w)) give the four
//          pooled outputs of that pixel; results go to shared memory [row][x][cm].
// Phase 2: thread = (row, dz, x): gathers the Cu channels c*Dz+dz from shared memory and writes one 16-byte voxel.
__device__ __forceinline__ uint32_t lo16(uint32_t v) { return v & 0xffffu; }
__device__ __forceinline__ uint32_t hi16(uint32_t v) { return v >> 16; }

template <bool BF16>
__device__ __forceinline__ float h2f(uint32_t raw16) {
  const uint16_t r = static_cast<uint16_t>(raw16);
  return H16<BF16>::to_f(*reinterpret_cast<const typename H16<BF16>::T*>(&r));
}
template <bool BF16>
__device__ __forceinline__ uint16_t f2h(float v) {
  typename H16<BF16>::T hv = H16<BF16>::from_f(v);
  return *reinterpret_cast<uint16_t*>(&hv);
}

template <bool BF16>
__global__ void __launch_bounds__(256)
head_shuffle_pool_fwd_kernel(const uint16_t* __restrict__ dec, uint16_t* __restrict__ u, int h, int w,
                             int Cm, int Dz, int Cc, int Cu, int pool) {
  extern __shared__ uint16_t sm[];  // [2][32][Cm]
  const int w0 = blockIdx.x * 16, hh = blockIdx.y;
  const long long n = blockIdx.z;
  const int Hs = 2 * h, Ws = 2 * w;
  const uint2* dec2 = reinterpret_cast<const uint2*>(dec);  // one uint2 = the 4 sub-pixels of one cm
  for (int idx = threadIdx.x; idx < 16 * Cm; idx += blockDim.x) {
    const int wl = idx / Cm, cm = idx - wl * Cm;
    const int ww = w0 + wl;
    float p00 = 0.f, p01 = 0.f, p10 = 0.f, p11 = 0.f;  // P[2h+i][2w+j]
    if (ww < w) {
      auto ld = [&](int a, int b) -> uint2 {  // decoder pixel (hh - a, ww - b) or zeros outside
        if (hh - a < 0 || ww - b < 0) return make_uint2(0u, 0u);
        return __ldg(dec2 + ((n * h + hh - a) * w + ww - b) * Cm + cm);
      };
      const uint2 c = ld(0, 0);
      const float s00 = h2f<BF16>(lo16(c.x)), s01 = h2f<BF16>(hi16(c.x)), s10 = h2f<BF16>(lo16(c.y)), s11 = h2f<BF16>(hi16(c.y));
      if (pool) {
        const uint2 l = ld(0, 1), t = ld(1, 0), tl = ld(1, 1);
        const float l01 = h2f<BF16>(hi16(l.x)), l11 = h2f<BF16>(hi16(l.y));      // S[2h+i][2w-1]
        const float t10 = h2f<BF16>(lo16(t.y)), t11 = h2f<BF16>(hi16(t.y));      // S[2h-1][2w+j]
        const float tl11 = h2f<BF16>(hi16(tl.y));                                // S[2h-1][2w-1]
        p00 = 0.25f * (s00 + l01 + t10 + tl11);
        p01 = 0.25f * (s01 + s00 + t11 + t10);
        p10 = 0.25f * (s10 + l11 + s00 + l01);
        p11 = 0.25f * (s11 + s10 + s01 + s00);
      } else {
        p00 = s00; p01 = s01; p10 = s10; p11 = s11;
      }
    }
    sm[(0 * 32 + 2 * wl) * Cm + cm] = f2h<BF16>(p00);
    sm[(0 * 32 + 2 * wl + 1) * Cm + cm] = f2h<BF16>(p01);
    sm[(1 * 32 + 2 * wl) * Cm + cm] = f2h<BF16>(p10);
    sm[(1 * 32 + 2 * wl + 1) * Cm + cm] = f2h<BF16>(p11);
  }
  __syncthreads();
  const int C8 = Cu / 8;
  for (int idx = threadIdx.x; idx < 2 * Dz * 32 * C8; idx += blockDim.x) {
    const int c8 = idx % C8;
    int t = idx / C8;
    const int xl = t & 31;
    t >>= 5;
    const int dz = t % Dz, r = t / Dz;
    const int X = 2 * w0 + xl, Y = 2 * hh + r;
    if (X < Ws && Y < Hs) {
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ca = c8 * 8 + 2 * k, cb = ca + 1;
        const uint32_t va = ca < Cc ? sm[(r * 32 + xl) * Cm + ca * Dz + dz] : 0u;
        const uint32_t vb = cb < Cc ? sm[(r * 32 + xl) * Cm + cb * Dz + dz] : 0u;
        o[k] = va | (vb << 16);
      }
      *reinterpret_cast<uint4*>(u + ((((n * Dz + dz) * Hs + Y) * Ws + X) * (long long)Cu + c8 * 8)) =
          make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ddec[n,h,w, cm*4+i*2+j] = dS[n,cm,2h+i,2w+j];  dS[Y,X] = pool ? 0.25*(dP[Y,X]+dP[Y,X+1]+dP[Y+1,X]+dP[Y+1,X+1]) : dP
// Phase 1: 16-byte loads of du rows 2h..2h+2, columns 2w0..2w0+32 into shared memory [r][x][cm];
// Phase 2: thread = (decoder pixel, cm): 9 shared reads -> the 4 sub-pixel gradients -> one 8-byte store.
template <bool BF16>
__global__ void __launch_bounds__(256)
head_shuffle_pool_bwd_kernel(const uint16_t* __restrict__ du, uint16_t* __restrict__ ddec, int h, int w,
                             int Cm, int Dz, int Cc, int Cu, int pool) {
  extern __shared__ uint16_t sm[];  // [3][33][Cm]
  const int w0 = blockIdx.x * 16, hh = blockIdx.y;
  const long long n = blockIdx.z;
  const int Hs = 2 * h, Ws = 2 * w;
  const int C8 = Cu / 8;
  for (int idx = threadIdx.x; idx < 3 * Dz * 33 * C8; idx += blockDim.x) {
    const int c8 = idx % C8;
    int t = idx / C8;
    const int xl = t % 33;
    t /= 33;
    const int dz = t % Dz, r = t / Dz;
    const int Y = 2 * hh + r, X = 2 * w0 + xl;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (Y < Hs && X < Ws) q = __ldg(reinterpret_cast<const uint4*>(du + ((((n * Dz + dz) * Hs + Y) * Ws + X) * (long long)Cu + c8 * 8)));
    const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ca = c8 * 8 + 2 * k, cb = ca + 1;
      if (ca < Cc) sm[(r * 33 + xl) * Cm + ca * Dz + dz] = static_cast<uint16_t>(lo16(w4[k]));
      if (cb < Cc) sm[(r * 33 + xl) * Cm + cb * Dz + dz] = static_cast<uint16_t>(hi16(w4[k]));
    }
  }
  __syncthreads();
  uint2* out2 = reinterpret_cast<uint2*>(ddec);
  for (int idx = threadIdx.x; idx < 16 * Cm; idx += blockDim.x) {
    const int wl = idx / Cm, cm = idx - wl * Cm;
    const int ww = w0 + wl;
    if (ww >= w) continue;
    auto at = [&](int rr, int xx) { return h2f<BF16>(sm[(rr * 33 + xx) * Cm + cm]); };
    const int x0 = 2 * wl;
    float g00, g01, g10, g11;
    if (pool) {
      const float a00 = at(0, x0), a01 = at(0, x0 + 1), a02 = at(0, x0 + 2);
      const float a10 = at(1, x0), a11 = at(1, x0 + 1), a12 = at(1, x0 + 2);
      const float a20 = at(2, x0), a21 = at(2, x0 + 1), a22 = at(2, x0 + 2);
      g00 = 0.25f * (a00 + a01 + a10 + a11);
      g01 = 0.25f * (a01 + a02 + a11 + a12);
      g10 = 0.25f * (a10 + a11 + a20 + a21);
      g11 = 0.25f * (a11 + a12 + a21 + a22);
    } else {
      g00 = at(0, x0); g01 = at(0, x0 + 1); g10 = at(1, x0); g11 = at(1, x0 + 1);
    }
    const uint32_t lo = f2h<BF16>(g00) | (static_cast<uint32_t>(f2h<BF16>(g01)) << 16);
    const uint32_t hi = f2h<BF16>(g10) | (static_cast<uint32_t>(f2h<BF16>(g11)) << 16);
    out2[((n * h + hh) * w + ww) * Cm + cm] = make_uint2(lo, hi);
  }
}

// per-sample, per-channel sum and sum of squares of z [B, R, C] (C % 8 == 0, C <= 256)
template <bool BF16>
__global__ void __launch_bounds__(256)
instnorm_stats_kernel(const uint4* __restrict__ z, float* __restrict__ sum, float* __restrict__ sumsq,
                      int R, int C8, int rows_per_block) {
  __shared__ float red[2 * 256];
  const int n = blockIdx.y;
  const int v = threadIdx.x % C8, rl = threadIdx.x / C8, rstep = blockDim.x / C8;
  const int C = C8 * 8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < rstep) {
    for (int r = r0 + rl; r < r1; r += rstep) {
      const uint4 t = __ldg(z + ((long long)n * R + r) * C8 + v);
      const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = H16<BF16>::unpack(w4[k]);
        s[2 * k] += f.x;
        s[2 * k + 1] += f.y;
        q[2 * k] = fmaf(f.x, f.x, q[2 * k]);
        q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&red[v * 8 + k], s[k]);
      atomicAdd(&red[C + v * 8 + k], q[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(sum + (long long)n * C + i, red[i]);
    atomicAdd(sumsq + (long long)n * C + i, red[C + i]);
  }
}

__global__ void instnorm_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq,
                                         float* __restrict__ mean, float* __restrict__ rstd, int total,
                                         float inv_count, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float m = sum[i] * inv_count;
  const float var = fmaxf(sumsq[i] * inv_count - m * m, 0.f);
  mean[i] = m;
  rstd[i] = rsqrtf(var + eps);
}


// out[n, o>>2, dz, 2y + ((o>>1)&1), 2x + (o&1)] = b1[o] + sum_c W1[o][c] * prelu((z - mean) * rstd)
// one thread per row (n, dz, y, x); z [B, R, Cmid], R = Dz*H*W
template <bool BF16, int HT_MAXO>
__global__ void __launch_bounds__(128)
head_tail_fwd_kernel(const uint4* __restrict__ z, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ alpha, int alpha_n,
                     const float* __restrict__ W1, const float* __restrict__ b1,
                     uint16_t* __restrict__ out, int Dz, int H, int W, int Cmid, int Co4) {
  extern __shared__ float smf[];  // W1 [Co4][Cmid], b1[Co4], mean[Cmid], rstd[Cmid], alpha[Cmid]
  float* sW = smf;
  float* sb = sW + Co4 * Cmid;
  float* sm_ = sb + Co4;
  float* sr = sm_ + Cmid;
  float* sa = sr + Cmid;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < Co4 * Cmid; i += blockDim.x) sW[i] = W1[i];
  for (int i = threadIdx.x; i < Co4; i += blockDim.x) sb[i] = b1[i];
  for (int i = threadIdx.x; i < Cmid; i += blockDim.x) {
    sm_[i] = mean[(long long)n * Cmid + i];
    sr[i] = rstd[(long long)n * Cmid + i];
    sa[i] = alpha[alpha_n == 1 ? 0 : i];
  }
  __syncthreads();
  const int R = Dz * H * W;  // host guarantees per-sample extents < 2^31
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= R) return;
  const int x = row % W;
  const int y = (row / W) % H;
  const int dz = row / (W * H);
  z += (long long)n * R * (Cmid / 8);
  out += (long long)n * (Co4 / 4) * Dz * 4 * H * W;
  float acc[HT_MAXO];
#pragma unroll
  for (int o = 0; o < HT_MAXO; ++o) acc[o] = o < Co4 ? sb[o] : 0.f;
  const int C8 = Cmid / 8;
  for (int v = 0; v < C8; ++v) {
    const uint4 t = __ldg(z + row * C8 + v);
    const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
    float a[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = H16<BF16>::unpack(w4[k]);
      a[2 * k] = f.x;
      a[2 * k + 1] = f.y;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = v * 8 + k;
      const float p = (a[k] - sm_[c]) * sr[c];
      a[k] = p > 0.f ? p : p * sa[c];
    }
#pragma unroll
    for (int o = 0; o < HT_MAXO; ++o)
      if (o < Co4) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[o] = fmaf(sW[o * Cmid + v * 8 + k], a[k], acc[o]);
      }
  }
  const int Co = Co4 / 4;
#pragma unroll
  for (int o = 0; o < HT_MAXO; o += 2)
    if (o < Co4) {
      const int co = o >> 2, i = (o >> 1) & 1;
      const int oi = ((co * Dz + dz) * (2 * H) + 2 * y + i) * (2 * W) + 2 * x;
      *reinterpret_cast<uint32_t*>(out + oi) = H16<BF16>::pack(acc[o], acc[o + 1]);
    }
}

// Backward of the fused tail.  thread = (voxel row, 8-channel chunk v); W1 sits in shared memory as [c][MAXO].
// MODE 0: reductions sdp[n,c] = sum dpre, sdpx[n,c] = sum dpre*xhat, db1[o] = sum dt, dalpha = sum da*xhat*(xhat<=0);
//         also materialises act = prelu(xhat) [M,Cmid] and dt (the un-shuffled output gradient) [M,ldt] so that
//         dW1 = dt^T act runs on the tensor cores (MN-major tcgen05 GEMM).
// MODE 1: dz = rstd * (dpre - sdp/R - xhat * sdpx/R), dbz[c] += sum dz
template <bool BF16, int MODE, int MAXO>
__global__ void __launch_bounds__(256, 2)
head_tail_bwd_kernel(const uint4* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ alpha, int alpha_n, const float* __restrict__ W1,
                     const uint16_t* __restrict__ dout, float* __restrict__ sdp, float* __restrict__ sdpx,
                     float* __restrict__ db1, float* __restrict__ dalpha, uint4* __restrict__ act_out,
                     uint16_t* __restrict__ dt_out, uint4* __restrict__ dz, float* __restrict__ dbz, int Dz, int H,
                     int W, int Cmid, int Co4, int ldt, int rows_per_block) {
  extern __shared__ float smf[];
  float* sW = smf;                 // [Cmid][MAXO]
  float* red = sW + Cmid * MAXO;   // MODE 0: sdp[Cmid] sdpx[Cmid] dalpha[Cmid] db1[MAXO];  MODE 1: dbz[Cmid]
  const int n = blockIdx.y;
  const int C8 = Cmid / 8;
  const int R = Dz * H * W;
  const int nred = MODE == 0 ? 3 * Cmid + MAXO : Cmid;
  for (int i = threadIdx.x; i < Cmid * MAXO; i += blockDim.x) {
    const int c = i / MAXO, o = i % MAXO;
    sW[i] = o < Co4 ? W1[o * Cmid + c] : 0.f;
  }
  for (int i = threadIdx.x; i < nred; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int v = threadIdx.x % C8, rl = threadIdx.x / C8, rstep = blockDim.x / C8;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  const int Co = Co4 / 4;
  z += (long long)n * R * C8;
  dout += (long long)n * Co * Dz * 4 * H * W;
  if (MODE == 0) {
    act_out += (long long)n * R * C8;
    dt_out += (long long)n * R * ldt;
  } else {
    dz += (long long)n * R * C8;
  }
  const int HW = H * W;
  float mu8[8], rs8[8], al8[8], m1[8], m2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k;
    mu8[k] = mean[(long long)n * Cmid + c];
    rs8[k] = rstd[(long long)n * Cmid + c];
    al8[k] = alpha[alpha_n == 1 ? 0 : c];
    m1[k] = MODE == 1 ? sdp[(long long)n * Cmid + c] / (float)R : 0.f;
    m2[k] = MODE == 1 ? sdpx[(long long)n * Cmid + c] / (float)R : 0.f;
  }
  const float4* sw4 = reinterpret_cast<const float4*>(sW + v * 8 * MAXO);
  float acc0[8], acc1[8], acc2[8], accb[MAXO];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc0[k] = acc1[k] = acc2[k] = 0.f;
#pragma unroll
  for (int o = 0; o < MAXO; ++o) accb[o] = 0.f;

#pragma unroll 2
  for (int row = r0 + rl; row < r1; row += rstep) {
    const int dzi = row / HW;
    const int rem = row - dzi * HW;
    const int y = rem / W;
    const int x = rem - y * W;
    const uint4 t = __ldg(z + row * C8 + v);
    float dt[MAXO];
#pragma unroll
    for (int o = 0; o < MAXO; o += 2) {
      dt[o] = dt[o + 1] = 0.f;
      if (o < Co4) {
        const int co = o >> 2, i = (o >> 1) & 1;
        const int oi = ((co * Dz + dzi) * (2 * H) + 2 * y + i) * (2 * W) + 2 * x;
        const float2 f = H16<BF16>::unpack(__ldg(reinterpret_cast<const uint32_t*>(dout + oi)));
        dt[o] = f.x;
        dt[o + 1] = f.y;
      }
    }
    const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
    float xh[8], o8[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = H16<BF16>::unpack(w4[k]);
      xh[2 * k] = (f.x - mu8[2 * k]) * rs8[2 * k];
      xh[2 * k + 1] = (f.y - mu8[2 * k + 1]) * rs8[2 * k + 1];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float da = 0.f;
#pragma unroll
      for (int o4 = 0; o4 < MAXO / 4; ++o4) {
        const float4 wv = sw4[k * (MAXO / 4) + o4];
        da = fmaf(wv.x, dt[o4 * 4], da);
        da = fmaf(wv.y, dt[o4 * 4 + 1], da);
        da = fmaf(wv.z, dt[o4 * 4 + 2], da);
        da = fmaf(wv.w, dt[o4 * 4 + 3], da);
      }
      const bool pos = xh[k] > 0.f;
      const float dpre = pos ? da : da * al8[k];
      if (MODE == 0) {
        o8[k] = pos ? xh[k] : xh[k] * al8[k];  // act
        acc0[k] += dpre;
        acc1[k] = fmaf(dpre, xh[k], acc1[k]);
        acc2[k] += pos ? 0.f : da * xh[k];
      } else {
        o8[k] = rs8[k] * (dpre - m1[k] - xh[k] * m2[k]);
      }
    }
    const uint4 q = make_uint4(H16<BF16>::pack(o8[0], o8[1]), H16<BF16>::pack(o8[2], o8[3]),
                               H16<BF16>::pack(o8[4], o8[5]), H16<BF16>::pack(o8[6], o8[7]));
    if (MODE == 0) {
      act_out[row * C8 + v] = q;
      if (v == 0) {
#pragma unroll
        for (int o = 0; o < MAXO; ++o) accb[o] += dt[o];
#pragma unroll
        for (int o8i = 0; o8i < MAXO / 8; ++o8i) {
          if (o8i * 8 < ldt)
            *reinterpret_cast<uint4*>(dt_out + (long long)row * ldt + o8i * 8) =
                make_uint4(H16<BF16>::pack(dt[o8i * 8], dt[o8i * 8 + 1]), H16<BF16>::pack(dt[o8i * 8 + 2], dt[o8i * 8 + 3]),
                           H16<BF16>::pack(dt[o8i * 8 + 4], dt[o8i * 8 + 5]), H16<BF16>::pack(dt[o8i * 8 + 6], dt[o8i * 8 + 7]));
        }
      }
    } else {
      dz[row * C8 + v] = q;
      const uint32_t w4o[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = H16<BF16>::unpack(w4o[k]);
        acc0[2 * k] += f.x;
        acc0[2 * k + 1] += f.y;
      }
    }
  }
  // combine the lanes that share a channel chunk (lane % C8), then one shared atomic per warp and value
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    for (int off = 16; off >= C8; off >>= 1) {
      acc0[k] += __shfl_xor_sync(0xffffffffu, acc0[k], off);
      if (MODE == 0) {
        acc1[k] += __shfl_xor_sync(0xffffffffu, acc1[k], off);
        acc2[k] += __shfl_xor_sync(0xffffffffu, acc2[k], off);
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int o = 0; o < MAXO; ++o)
      for (int off = 16; off >= C8; off >>= 1) accb[o] += __shfl_xor_sync(0xffffffffu, accb[o], off);
  }
  if (lane < C8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = v * 8 + k;
      atomicAdd(&red[c], acc0[k]);
      if (MODE == 0) {
        atomicAdd(&red[Cmid + c], acc1[k]);
        atomicAdd(&red[2 * Cmid + c], acc2[k]);
      }
    }
    if (MODE == 0 && v == 0) {
#pragma unroll
      for (int o = 0; o < MAXO; ++o) atomicAdd(&red[3 * Cmid + o], accb[o]);
    }
  }
  __syncthreads();
  if (MODE == 0) {
    for (int i = threadIdx.x; i < Cmid; i += blockDim.x) {
      atomicAdd(sdp + (long long)n * Cmid + i, red[i]);
      atomicAdd(sdpx + (long long)n * Cmid + i, red[Cmid + i]);
      atomicAdd(dalpha + (alpha_n == 1 ? 0 : i), red[2 * Cmid + i]);
    }
    for (int i = threadIdx.x; i < Co4; i += blockDim.x) atomicAdd(db1 + i, red[3 * Cmid + i]);
  } else {
    for (int i = threadIdx.x; i < Cmid; i += blockDim.x) atomicAdd(dbz + i, red[i]);
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_head_shuffle_pool(const void* src, void* dst, int B, int h, int w, int Cm, int Dz,
                                       int Cu, int pool, int backward, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(src && dst, "null pointer");
  VB_REQUIRE(Cm % Dz == 0, "mid channels %d not divisible by depth %d", Cm, Dz);
  const int Cc = Cm / Dz;
  VB_REQUIRE(Cu >= Cc && Cu % 8 == 0, "Cu (%d) must be a multiple of 8 >= %d", Cu, Cc);
  cudaStream_t st = (cudaStream_t)stream;
  if (!backward) {
    dim3 grid((w + 15) / 16, h, B);
    const size_t smem = (size_t)2 * 32 * Cm * 2;
    VB_SUPPORTED(smem <= 48 * 1024, "head: Cm (%d) too large", Cm);
    if (dtype == VB200_BF16)
      head_shuffle_pool_fwd_kernel<true><<<grid, 256, smem, st>>>((const uint16_t*)src, (uint16_t*)dst, h, w, Cm, Dz, Cc, Cu, pool);
    else
      head_shuffle_pool_fwd_kernel<false><<<grid, 256, smem, st>>>((const uint16_t*)src, (uint16_t*)dst, h, w, Cm, Dz, Cc, Cu, pool);
  } else {
    dim3 grid((w + 15) / 16, h, B);
    const size_t smem = (size_t)3 * 33 * Cm * 2;
    VB_SUPPORTED(smem <= 48 * 1024, "head: Cm (%d) too large", Cm);
    if (dtype == VB200_BF16)
      head_shuffle_pool_bwd_kernel<true><<<grid, 256, smem, st>>>((const uint16_t*)src, (uint16_t*)dst, h, w, Cm, Dz, Cc, Cu, pool);
    else
      head_shuffle_pool_bwd_kernel<false><<<grid, 256, smem, st>>>((const uint16_t*)src, (uint16_t*)dst, h, w, Cm, Dz, Cc, Cu, pool);
  }
  return check_launch("vb200_head_shuffle_pool");
}

extern "C" int vb200_instnorm_stats(const void* z, float* sum, float* sumsq, float* mean, float* rstd, int B,
                                    int64_t R, int C, float eps, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(z && sum && sumsq && mean && rstd, "null pointer");
  VB_SUPPORTED(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0, "instnorm C (%d) must be 8*2^k <= 256", C);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(sum, 0, sizeof(float) * B * C, st);
  cudaMemsetAsync(sumsq, 0, sizeof(float) * B * C, st);
  long long rpb = (R * B + 148 * 8 - 1) / (148 * 8);
  if (rpb < 64) rpb = 64;
  if (rpb > R) rpb = R;
  dim3 grid((unsigned)((R + rpb - 1) / rpb), B);
  if (dtype == VB200_BF16)
    instnorm_stats_kernel<true><<<grid, 256, 0, st>>>((const uint4*)z, sum, sumsq, (int)R, C / 8, (int)rpb);
  else
    instnorm_stats_kernel<false><<<grid, 256, 0, st>>>((const uint4*)z, sum, sumsq, (int)R, C / 8, (int)rpb);
  if (int rc = check_launch("vb200_instnorm_stats")) return rc;
  instnorm_finalize_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(sum, sumsq, mean, rstd, B * C, 1.0f / (float)R, eps);
  return check_launch("vb200_instnorm_finalize");
}

extern "C" int vb200_head_tail_fwd(const void* z, const float* mean, const float* rstd, const float* alpha,
                                   int alpha_n, const float* W1, const float* b1, void* out, int B, int Dz,
                                   int H, int W, int Cmid, int Co4, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(z && mean && rstd && alpha && W1 && b1 && out, "null pointer");
  VB_SUPPORTED(Cmid % 8 == 0 && Co4 % 4 == 0 && Co4 <= 16, "head tail: Cmid %d / Co4 %d", Cmid, Co4);
  VB_REQUIRE(alpha_n == 1 || alpha_n == Cmid, "PReLU parameter count %d", alpha_n);
  const long long R = (long long)Dz * H * W;
  VB_SUPPORTED(R * Cmid < (1LL << 31) && R * Co4 < (1LL << 31), "head: per-sample extent too large");
  dim3 grid((unsigned)((R + 127) / 128), B);
  const size_t smem = sizeof(float) * (Co4 * Cmid + Co4 + 3 * Cmid);
  cudaStream_t st = (cudaStream_t)stream;
#define HTF(BF, MO) head_tail_fwd_kernel<BF, MO><<<grid, 128, smem, st>>>((const uint4*)z, mean, rstd, alpha, alpha_n, W1, b1, (uint16_t*)out, Dz, H, W, Cmid, Co4)
  if (dtype == VB200_BF16) { if (Co4 <= 4) HTF(true, 4); else if (Co4 <= 8) HTF(true, 8); else HTF(true, 16); }
  else if (dtype == VB200_FP16) { if (Co4 <= 4) HTF(false, 4); else if (Co4 <= 8) HTF(false, 8); else HTF(false, 16); }
  else return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
#undef HTF
  return check_launch("vb200_head_tail_fwd");
}

/* phase 0: reductions into sdp, sdpx [B,Cmid], db1 [Co4], dalpha [alpha_n] (all pre-zeroed) and the materialised
 *          act [B,R,Cmid] / dt [B,R,ldt] operands of the dW1 GEMM (ldt = Co4 rounded up to 8, zero padded)
 * phase 1: dz [B,R,Cmid] and dbz [Cmid] (pre-zeroed) */
extern "C" int vb200_head_tail_bwd(int phase, const void* z, const float* mean, const float* rstd,
                                   const float* alpha, int alpha_n, const float* W1, const void* dout,
                                   float* sdp, float* sdpx, float* db1, float* dalpha, void* act_out, void* dt_out,
                                   void* dz, float* dbz, int B, int Dz, int H, int W, int Cmid, int Co4, int dtype,
                                   vb200_stream_t stream) {
  VB_REQUIRE(z && mean && rstd && alpha && W1 && dout && sdp && sdpx, "null pointer");
  VB_SUPPORTED(Cmid % 8 == 0 && Co4 % 4 == 0 && Co4 <= 16 && 32 % (Cmid / 8) == 0 && Cmid <= 256,
               "head tail: Cmid %d / Co4 %d", Cmid, Co4);
  const long long R = (long long)Dz * H * W;
  VB_SUPPORTED(R * Cmid < (1LL << 31) && R * 16 < (1LL << 31), "head: per-sample extent too large");
  long long rpb = (R * B + 148 * 16 - 1) / (148 * 16);
  if (rpb < 256) rpb = 256;
  if (rpb > R) rpb = R;
  dim3 grid((unsigned)((R + rpb - 1) / rpb), B);
  cudaStream_t st = (cudaStream_t)stream;
  const int maxo = Co4 <= 8 ? 8 : 16;
  const int ldt = (Co4 + 7) / 8 * 8;
  const size_t smem = sizeof(float) * (Cmid * maxo + 3 * Cmid + maxo);
#define HT2(BF, MODE, MO)                                                                                             \
  head_tail_bwd_kernel<BF, MODE, MO><<<grid, 256, smem, st>>>((const uint4*)z, mean, rstd, alpha, alpha_n, W1,        \
                                                             (const uint16_t*)dout, sdp, sdpx, db1, dalpha,          \
                                                             (uint4*)act_out, (uint16_t*)dt_out, (uint4*)dz, dbz, Dz, \
                                                             H, W, Cmid, Co4, ldt, (int)rpb)
#define HT(BF, MODE) do { if (maxo == 8) HT2(BF, MODE, 8); else HT2(BF, MODE, 16); } while (0)
  if (phase == 0) {
    VB_REQUIRE(db1 && dalpha && act_out && dt_out, "null pointer");
    if (dtype == VB200_BF16) HT(true, 0); else if (dtype == VB200_FP16) HT(false, 0); else return fail(VB200_ERR_UNSUPPORTED, "dtype");
  } else {
    VB_REQUIRE(dz && dbz, "null pointer");
    if (dtype == VB200_BF16) HT(true, 1); else if (dtype == VB200_FP16) HT(false, 1); else return fail(VB200_ERR_UNSUPPORTED, "dtype");
  }
#undef HT
#undef HT2
  return check_launch("vb200_head_tail_bwd");
}
