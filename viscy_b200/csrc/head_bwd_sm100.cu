// UNeXt2 head tail backward, streaming form (PixelToVoxelHead: InstanceNorm3d + PReLU + Conv3d(k=1) + PixelShuffle(2),
// VM/components/heads.py:607-641) for the BASELINE head geometry (Cmid = 32, 4*Cout = 8).
//
// The generic two-phase kernels in head_sm100.cu materialise act = prelu(xhat) (176 MB at config 2) and the un-shuffled
// output gradient so that dW1 = dt^T act can run as a GEMM, and they fetch dout with scattered 4-byte loads.  Here one
// elected thread streams 256-row tiles of z (one contiguous 16 KB bulk copy) and the matching pixel-shuffled dout lines
// (4 planes x lines, 4W bytes each) through a 3-stage mbarrier ring; the 256 threads (thread = row x 8-channel
// chunk, W1 slice in registers) form everything in registers:
//   phase 0: sdp[n,c] = sum dpre, sdpx[n,c] = sum dpre*xhat, dalpha, db1[o] = sum dt[o], dW1[o][c] = sum dt[o]*act[c]
//   phase 1: dz = rstd * (dpre - sdp/R - xhat * sdpx/R) (coalesced 16-byte stores), dbz[c] = sum dz
// Algorithmic traffic: phase 0 reads z + dout (220 MB at config 2), phase 1 reads the same and writes dz (176 MB).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vb {
namespace hb {

constexpr int TR = 256, STAGES = 3, CM = 32, CO4 = 8, NCOMP = 256;
constexpr int ZB = TR * CM * 2, DB = TR * CO4 * 2;  // bytes per stage: z tile, dout tile

struct Params {
  const uint8_t* z;      // [B, R, 32] 16-bit
  const uint8_t* dout;   // [B, 2, Dz, 2H, 2W] 16-bit
  const float* mean;     // [B, 32]
  const float* rstd;
  const float* alpha;
  const float* W1;       // [8][32]
  float* sdp;            // [B, 32]  (phase 0: out; phase 1: in)
  float* sdpx;
  float* db1;            // [8]
  float* dalpha;         // [alpha_n]
  float* dW1;            // [8][32]
  uint4* dz;             // [B, R, 32] 16-bit
  float* dbz;            // [32]
  int alpha_n, Dz, H, W, R, tiles_per_sample, n_tiles;
};

__device__ __forceinline__ void named_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <bool BF16, int MODE>
__global__ void __launch_bounds__(NCOMP, 1) head_bwd_stream_kernel(const Params p) {
  using H = H16<BF16>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* zs = smem;                          // [STAGES][ZB]
  uint8_t* ds = smem + STAGES * ZB;            // [STAGES][DB]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (ZB + DB));
  uint64_t* empty = full + STAGES;
  float* red = reinterpret_cast<float*>(empty + STAGES);  // phase 0: 360 floats, phase 1: 32
  constexpr int NRED = MODE == 0 ? 3 * CM + CO4 * CM + CO4 : CM;

  const int t0 = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
  const int t1 = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCOMP);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NRED; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  // thread 0 doubles as the producer: it issues the bulk copies of tile it + STAGES - 1 before working on tile it
  // (256 threads = 8 warps: a ninth warp would cost a fourth of the register file to the 4-warp allocation granule)
  auto produce = [&](int it) {
    const int t = t0 + it;
    if (t >= t1) return;
    const int s = it % STAGES;
    mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
    mbar_expect_tx(&full[s], ZB + DB);
    const int n = t / p.tiles_per_sample, tt = t - n * p.tiles_per_sample;
    const long long row0 = (long long)tt * TR;
    bulk_load_1d(zs + s * ZB, p.z + ((long long)n * p.R + row0) * (CM * 2), ZB, &full[s]);
    const int L = TR / p.W, line0 = (int)(row0 / p.W);
    const uint32_t seg = (uint32_t)p.W * 4;  // 2W elements of 2 bytes
    for (int l = 0; l < L; ++l) {
      const int ln = line0 + l, dzi = ln / p.H, y = ln - dzi * p.H;
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) {
        const int co = pl >> 1, i = pl & 1;
        const long long off = ((((long long)n * 2 + co) * p.Dz + dzi) * (2 * p.H) + 2 * y + i) * (2LL * p.W);
        bulk_load_1d(ds + s * DB + (l * 4 + pl) * seg, p.dout + off * 2, seg, &full[s]);
      }
    }
  };
  if (threadIdx.x == 0) {
    for (int it = 0; it < STAGES - 1; ++it) produce(it);
  }
  // ------------------------------------------------------------------ consumers
  const int v = threadIdx.x & 3, rl = threadIdx.x >> 2;  // channel chunk, row within a 64-row pass
  float2 w2[4][8];  // w2[kp][o] = (W1[o][v*8 + 2kp], W1[o][v*8 + 2kp + 1]): channel pairs ride the packed fp32x2 pipe
#pragma unroll
  for (int kp = 0; kp < 4; ++kp)
#pragma unroll
    for (int o = 0; o < 8; ++o)
      w2[kp][o] = make_float2(__ldg(p.W1 + o * CM + v * 8 + 2 * kp), __ldg(p.W1 + o * CM + v * 8 + 2 * kp + 1));
  float al[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) al[k] = __ldg(p.alpha + (p.alpha_n == 1 ? 0 : v * 8 + k));
  float mu[8], rs[8], m1[8], m2[8];
  float a_dp[8], a_dpx[8], a_al[8], a_db[8];
  float2 a_dw2[8][4];  // [o][channel pair]
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a_dp[k] = a_dpx[k] = a_al[k] = a_db[k] = 0.f;
#pragma unroll
    for (int o = 0; o < 8; ++o) a_dw2[o][k >> 1] = make_float2(0.f, 0.f);
  }
  int cur_n = -1;
  const int lane = threadIdx.x & 31;

  auto flush = [&](int n) {
    // lanes sharing a chunk (lane % 4) -> lanes 0..3, then shared atomics, then one global atomic per value
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int off = 16; off >= 4; off >>= 1) {
        a_dp[k] += __shfl_xor_sync(0xffffffffu, a_dp[k], off);
        if (MODE == 0) {
          a_dpx[k] += __shfl_xor_sync(0xffffffffu, a_dpx[k], off);
          a_al[k] += __shfl_xor_sync(0xffffffffu, a_al[k], off);
          a_db[k] += __shfl_xor_sync(0xffffffffu, a_db[k], off);
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            float& e = (k & 1) ? a_dw2[o][k >> 1].y : a_dw2[o][k >> 1].x;
            e += __shfl_xor_sync(0xffffffffu, e, off);
          }
        }
      }
    }
    if (lane < 4) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = v * 8 + k;
        atomicAdd(&red[c], a_dp[k]);
        if (MODE == 0) {
          atomicAdd(&red[CM + c], a_dpx[k]);
          atomicAdd(&red[2 * CM + c], a_al[k]);
#pragma unroll
          for (int o = 0; o < 8; ++o) atomicAdd(&red[3 * CM + o * CM + c], (k & 1) ? a_dw2[o][k >> 1].y : a_dw2[o][k >> 1].x);
        }
      }
      if (MODE == 0 && v == 0) {
#pragma unroll
        for (int o = 0; o < 8; ++o) atomicAdd(&red[3 * CM + CO4 * CM + o], a_db[o]);
      }
    }
    named_sync();
    for (int i = threadIdx.x; i < NRED; i += NCOMP) {
      const float val = red[i];
      if (MODE == 0) {
        if (i < CM) atomicAdd(p.sdp + (long long)n * CM + i, val);
        else if (i < 2 * CM) atomicAdd(p.sdpx + (long long)n * CM + i - CM, val);
        else if (i < 3 * CM) atomicAdd(p.dalpha + (p.alpha_n == 1 ? 0 : i - 2 * CM), val);
        else if (i < 3 * CM + CO4 * CM) atomicAdd(p.dW1 + i - 3 * CM, val);
        else atomicAdd(p.db1 + i - 3 * CM - CO4 * CM, val);
      } else {
        atomicAdd(p.dbz + i, val);
      }
      red[i] = 0.f;
    }
    named_sync();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a_dp[k] = a_dpx[k] = a_al[k] = a_db[k] = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o) a_dw2[o][k >> 1] = make_float2(0.f, 0.f);
    }
  };

  for (int t = t0, it = 0; t < t1; ++t, ++it) {
    const int s = it % STAGES;
    const int n = t / p.tiles_per_sample, tt = t - n * p.tiles_per_sample;
    if (n != cur_n) {
      if (cur_n >= 0) flush(cur_n);
      cur_n = n;
      const float invR = 1.0f / (float)p.R;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = v * 8 + k;
        mu[k] = __ldg(p.mean + (long long)n * CM + c);
        rs[k] = __ldg(p.rstd + (long long)n * CM + c);
        m1[k] = MODE == 1 ? p.sdp[(long long)n * CM + c] * invR : 0.f;
        m2[k] = MODE == 1 ? p.sdpx[(long long)n * CM + c] * invR : 0.f;
      }
    }
    if (threadIdx.x == 0) produce(it + STAGES - 1);
    __syncwarp();
    mbar_wait(&full[s], (it / STAGES) & 1);
    const uint4* zt = reinterpret_cast<const uint4*>(zs + s * ZB);
    const uint32_t* dt32 = reinterpret_cast<const uint32_t*>(ds + s * DB);
#pragma unroll 1
    for (int pass = 0; pass < TR / 64; ++pass) {
      const int r = pass * 64 + rl;
      const uint4 zq = zt[r * 4 + v];
      const int l = r / p.W, x = r - l * p.W;
      float dt[8];
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) {
        const float2 f = H::unpack(dt32[(l * 4 + pl) * p.W + x]);
        dt[2 * pl] = f.x;
        dt[2 * pl + 1] = f.y;
      }
      const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
      float xh[8], o8[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = H::unpack(zw[k]);
        xh[2 * k] = (f.x - mu[2 * k]) * rs[2 * k];
        xh[2 * k + 1] = (f.y - mu[2 * k + 1]) * rs[2 * k + 1];
      }
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        float2 da2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int o = 0; o < 8; ++o) da2 = __ffma2_rn(w2[kp][o], make_float2(dt[o], dt[o]), da2);
        float2 act2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = 2 * kp + h;
          const float da = h ? da2.y : da2.x;
          const bool pos = xh[k] > 0.f;
          const float dpre = pos ? da : da * al[k];
          if (MODE == 0) {
            (h ? act2.y : act2.x) = pos ? xh[k] : xh[k] * al[k];
            a_dp[k] += dpre;
            a_dpx[k] = fmaf(dpre, xh[k], a_dpx[k]);
            a_al[k] += pos ? 0.f : da * xh[k];
          } else {
            o8[k] = rs[k] * (dpre - m1[k] - xh[k] * m2[k]);
          }
        }
        if (MODE == 0) {
#pragma unroll
          for (int o = 0; o < 8; ++o) a_dw2[o][kp] = __ffma2_rn(make_float2(dt[o], dt[o]), act2, a_dw2[o][kp]);
        }
      }
      if (MODE == 0) {
#pragma unroll
        for (int o = 0; o < 8; ++o) a_db[o] += dt[o];
      } else {
        const uint4 q = make_uint4(H::pack(o8[0], o8[1]), H::pack(o8[2], o8[3]), H::pack(o8[4], o8[5]), H::pack(o8[6], o8[7]));
        p.dz[((long long)n * p.R + (long long)tt * TR + r) * 4 + v] = q;
        const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = H::unpack(qw[k]);
          a_dp[2 * k] += f.x;
          a_dp[2 * k + 1] += f.y;
        }
      }
    }
    mbar_arrive(&empty[s]);
  }
  if (cur_n >= 0) flush(cur_n);
}

}  // namespace hb
}  // namespace vb

using namespace vb;

// Streaming head-tail backward.  phase 0: sums (sdp, sdpx [B,32]; db1 [8]; dalpha; dW1 [8,32], all pre-zeroed);
// phase 1: dz [B,R,32] and dbz [32] (pre-zeroed) from the finished sums.  Returns VB200_ERR_UNSUPPORTED for any other
// geometry (the caller then uses vb200_head_tail_bwd).
extern "C" int vb200_head_tail_bwd_stream(int phase, const void* z, const float* mean, const float* rstd, const float* alpha,
                                          int alpha_n, const float* W1, const void* dout, float* sdp, float* sdpx, float* db1,
                                          float* dalpha, float* dW1, void* dz, float* dbz, int B, int Dz, int H, int W, int Cmid,
                                          int Co4, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(z && mean && rstd && alpha && W1 && dout && sdp && sdpx, "null pointer");
  const long long R = (long long)Dz * H * W;
  VB_SUPPORTED(Cmid == hb::CM && Co4 == hb::CO4 && W >= 4 && hb::TR % W == 0 && R % hb::TR == 0 && R < (1LL << 31) &&
                   (R / hb::TR) * B < (1LL << 31) && (alpha_n == 1 || alpha_n == Cmid),
               "head tail stream: Cmid %d / Co4 %d / W %d / R %lld", Cmid, Co4, W, R);
  VB_SUPPORTED(dtype == VB200_BF16 || dtype == VB200_FP16, "dtype %d", dtype);
  if (phase == 0) VB_REQUIRE(db1 && dalpha && dW1, "null pointer");
  else VB_REQUIRE(dz && dbz, "null pointer");
  hb::Params p{};
  p.z = (const uint8_t*)z; p.dout = (const uint8_t*)dout; p.mean = mean; p.rstd = rstd; p.alpha = alpha; p.W1 = W1;
  p.sdp = sdp; p.sdpx = sdpx; p.db1 = db1; p.dalpha = dalpha; p.dW1 = dW1; p.dz = (uint4*)dz; p.dbz = dbz;
  p.alpha_n = alpha_n; p.Dz = Dz; p.H = H; p.W = W; p.R = (int)R; p.tiles_per_sample = (int)(R / hb::TR);
  p.n_tiles = p.tiles_per_sample * B;
  const size_t smem = hb::STAGES * (hb::ZB + hb::DB) + 2 * hb::STAGES * sizeof(uint64_t) + 368 * sizeof(float);
  const void* fns[4] = {(const void*)hb::head_bwd_stream_kernel<true, 0>, (const void*)hb::head_bwd_stream_kernel<true, 1>,
                        (const void*)hb::head_bwd_stream_kernel<false, 0>, (const void*)hb::head_bwd_stream_kernel<false, 1>};
  static PerDeviceOnce once[4];
  const int slot = (dtype == VB200_BF16 ? 0 : 2) + (phase ? 1 : 0);
  const int dev = PerDeviceOnce::device();
  if (once[slot].need(dev)) {
    cudaFuncSetAttribute(fns[slot], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    once[slot].done(dev);
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (slot == 0) hb::head_bwd_stream_kernel<true, 0><<<grid, hb::NCOMP, smem, st>>>(p);
  else if (slot == 1) hb::head_bwd_stream_kernel<true, 1><<<grid, hb::NCOMP, smem, st>>>(p);
  else if (slot == 2) hb::head_bwd_stream_kernel<false, 0><<<grid, hb::NCOMP, smem, st>>>(p);
  else hb::head_bwd_stream_kernel<false, 1><<<grid, hb::NCOMP, smem, st>>>(p);
  return check_launch("vb200_head_tail_bwd_stream");
}
