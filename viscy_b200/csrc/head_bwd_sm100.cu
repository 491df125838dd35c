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
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vb {
namespace hb {

constexpr int TR = 256, STAGES = 3, MSTAGES = 5, CM = 32, CO4 = 8, NCOMP = 256;  // MSTAGES: ring depth of the MMA form (2 CTAs / SM)
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
  int wshift;  // log2(W): 256 % W == 0 makes W a power of two
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

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core form of the same two phases (legacy warp-level mma.sync: these 32 x 8 contractions are far too small for a
// tcgen05 tile).  A warp owns 32 rows of the staged tile; everything is kept TRANSPOSED ([channel][row]) so that one fragment
// layout serves all three roles (the flash-attention C-fragment -> A-fragment trick):
//   z^T      : ldmatrix.x4.trans of the row-major z tile            -> [ch g / g+8][rows 2q, 2q+1 (+8)] bf16 pairs
//   da^T     = W1^T [16 ch x 8 o] . dt^T [8 o x 8 rows]   (m16n8k8)  -> C fragment in the same [ch][row pair] layout
//   dW1^T   += act^T [16 ch x 16 rows] . dt [16 rows x 8 o] (m16n8k16), act^T packed straight from the element-wise results
// ~6 MMAs + 16 element-wise channel-row values per thread and 16 rows, instead of 128 FFMA2 per row chunk.
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void stsm_x4_trans(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
template <bool BF16>
__device__ __forceinline__ void mma_k8_zero(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  if constexpr (BF16)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %7, %7, %7};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
  else
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %7, %7, %7};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
}
template <bool BF16>
__device__ __forceinline__ void mma_k8_acc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  if constexpr (BF16)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  else
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
template <bool BF16>
__device__ __forceinline__ void mma_k16_acc(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool BF16, int MODE>
__global__ void __launch_bounds__(NCOMP, 2) head_bwd_mma_kernel(const Params p) {
  using H = H16<BF16>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* zs = smem;                          // [MSTAGES][ZB]
  uint8_t* ds = smem + MSTAGES * ZB;            // [MSTAGES][DB]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + MSTAGES * (ZB + DB));
  uint64_t* empty = full + MSTAGES;
  float* red = reinterpret_cast<float*>(empty + MSTAGES);
  constexpr int NRED = MODE == 0 ? 3 * CM + CO4 * CM + CO4 : CM;

  const int t0 = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
  const int t1 = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
  if (threadIdx.x == 0) {
    for (int s = 0; s < MSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCOMP);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NRED; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  auto produce = [&](int it) {
    const int t = t0 + it;
    if (t >= t1) return;
    const int s = it % MSTAGES;
    mbar_wait(&empty[s], ((it / MSTAGES) & 1) ^ 1);
    mbar_expect_tx(&full[s], ZB + DB);
    const int n = t / p.tiles_per_sample, tt = t - n * p.tiles_per_sample;
    const long long row0 = (long long)tt * TR;
    bulk_load_1d(zs + s * ZB, p.z + ((long long)n * p.R + row0) * (CM * 2), ZB, &full[s]);
    const int L = TR / p.W, line0 = (int)(row0 / p.W);
    const uint32_t seg = (uint32_t)p.W * 4;
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
    for (int it = 0; it < MSTAGES - 1; ++it) produce(it);
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  // this thread's four channels: ch(mt, hi) = 16 mt + g + 8 hi
  // W1^T A fragments of the da MMA (rows = channels, k = output o = 2q, 2q + 1), split into a 16-bit head and a 16-bit
  // remainder so that the fp32 weights enter the tensor-core product at ~16 mantissa bits (two MMAs per fragment)
  uint32_t wa[2][2], wb[2][2];
  float al[4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int ch = 16 * mt + g + 8 * hi;
      const float w0 = __ldg(p.W1 + (2 * q) * CM + ch), w1 = __ldg(p.W1 + (2 * q + 1) * CM + ch);
      wa[mt][hi] = H::pack(w0, w1);
      const float2 back = H::unpack(wa[mt][hi]);
      wb[mt][hi] = H::pack(w0 - back.x, w1 - back.y);
      al[mt * 2 + hi] = __ldg(p.alpha + (p.alpha_n == 1 ? 0 : ch));
    }
  float mu[4], rs[4], m1[4], m2[4];
  float a_dp[4], a_dpx[4], a_al[4], a_db[2], a_dw[2][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) a_dp[k] = a_dpx[k] = a_al[k] = a_dw[0][k] = a_dw[1][k] = 0.f;
  a_db[0] = a_db[1] = 0.f;
  int cur_n = -1;

  auto flush = [&](int n) {
    // per-channel sums: the four lanes of a quad (q = 0..3) hold different rows of the same channels
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int off = 1; off <= 2; off <<= 1) {
        a_dp[k] += __shfl_xor_sync(0xffffffffu, a_dp[k], off);
        if (MODE == 0) {
          a_dpx[k] += __shfl_xor_sync(0xffffffffu, a_dpx[k], off);
          a_al[k] += __shfl_xor_sync(0xffffffffu, a_al[k], off);
        }
      }
    }
    if (MODE == 0) {  // db1[o = 2q, 2q + 1]: the eight lanes with the same q hold different rows
#pragma unroll
      for (int off = 4; off <= 16; off <<= 1) {
        a_db[0] += __shfl_xor_sync(0xffffffffu, a_db[0], off);
        a_db[1] += __shfl_xor_sync(0xffffffffu, a_db[1], off);
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int ch = 16 * mt + g + 8 * hi, k = mt * 2 + hi;
        if (q == 0) {
          atomicAdd(&red[ch], a_dp[k]);
          if (MODE == 0) {
            atomicAdd(&red[CM + ch], a_dpx[k]);
            atomicAdd(&red[2 * CM + ch], a_al[k]);
          }
        }
        if (MODE == 0) {  // dW1[o][ch]: every lane owns its own (ch, o) entries
          atomicAdd(&red[3 * CM + (2 * q) * CM + ch], a_dw[mt][hi * 2]);
          atomicAdd(&red[3 * CM + (2 * q + 1) * CM + ch], a_dw[mt][hi * 2 + 1]);
        }
      }
    if (MODE == 0 && g == 0) {
      atomicAdd(&red[3 * CM + CO4 * CM + 2 * q], a_db[0]);
      atomicAdd(&red[3 * CM + CO4 * CM + 2 * q + 1], a_db[1]);
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
    for (int k = 0; k < 4; ++k) a_dp[k] = a_dpx[k] = a_al[k] = a_dw[0][k] = a_dw[1][k] = 0.f;
    a_db[0] = a_db[1] = 0.f;
  };

  for (int t = t0, it = 0; t < t1; ++t, ++it) {
    const int s = it % MSTAGES;
    const int n = t / p.tiles_per_sample, tt = t - n * p.tiles_per_sample;
    if (n != cur_n) {
      if (cur_n >= 0) flush(cur_n);
      cur_n = n;
      const float invR = 1.0f / (float)p.R;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const int ch = 16 * mt + g + 8 * hi, k = mt * 2 + hi;
          mu[k] = __ldg(p.mean + (long long)n * CM + ch);
          rs[k] = __ldg(p.rstd + (long long)n * CM + ch);
          m1[k] = MODE == 1 ? p.sdp[(long long)n * CM + ch] * invR : 0.f;
          m2[k] = MODE == 1 ? p.sdpx[(long long)n * CM + ch] * invR : 0.f;
        }
    }
    if (threadIdx.x == 0) produce(it + MSTAGES - 1);
    __syncwarp();
    mbar_wait(&full[s], (it / MSTAGES) & 1);
    const uint32_t zt = smem_u32(zs + s * ZB);
    const uint32_t* dt32 = reinterpret_cast<const uint32_t*>(ds + s * DB);
#pragma unroll 1
    for (int gi = 0; gi < 2; ++gi) {
      const int rb = warp * 32 + gi * 16;  // first of this group's 16 rows
      // dt fragments: B of the da MMA (row g of each half, outputs 2q, 2q+1) and B of the dW MMA (rows 2q, 2q+1, output g)
      uint32_t bda[2], bdw[2];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int r1 = rb + hf * 8 + g;
        const int l1 = r1 >> p.wshift, x1 = r1 & (p.W - 1);
        bda[hf] = dt32[(l1 * 4 + q) * p.W + x1];
        const int r2 = rb + hf * 8 + 2 * q;
        const int l2 = r2 >> p.wshift, x2 = r2 & (p.W - 1);
        const uint2 w2 = *reinterpret_cast<const uint2*>(dt32 + (l2 * 4 + (g >> 1)) * p.W + x2);
        bdw[hf] = __byte_perm(w2.x, w2.y, (g & 1) ? 0x7632 : 0x5410);
        if (MODE == 0) {
          const float2 f = H::unpack(bda[hf]);
          a_db[0] += f.x;
          a_db[1] += f.y;
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t zf[4];  // z^T: [hf * 2 + hi] = (ch 16 mt + g + 8 hi, rows rb + 8 hf + 2q, +1)
        const uint32_t addr = zt + static_cast<uint32_t>(((rb + ((lane >> 4) << 3) + (lane & 7)) * CM + 16 * mt + ((lane >> 3) & 1) * 8) * 2);
        ldsm_x4_trans(zf, addr);
        // ldmatrix order: matrix 0 = (rows 0-7, ch 0-7), 1 = (rows 0-7, ch 8-15), 2 = (rows 8-15, ch 0-7), 3 = (rows 8-15, ch 8-15)
        uint32_t af[4];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float da[4];
          mma_k8_zero<BF16>(da, wa[mt][0], wa[mt][1], bda[hf]);
          mma_k8_acc<BF16>(da, wb[mt][0], wb[mt][1], bda[hf]);
#pragma unroll
          for (int hi = 0; hi < 2; ++hi) {
            const int k = mt * 2 + hi;
            const float2 zz = H::unpack(zf[hf * 2 + hi]);
            float o2[2];
#pragma unroll
            for (int rp = 0; rp < 2; ++rp) {
              const float xh = ((rp ? zz.y : zz.x) - mu[k]) * rs[k];
              const float d = da[hi * 2 + rp];
              const bool pos = xh > 0.f;
              const float dpre = pos ? d : d * al[k];
              if (MODE == 0) {
                o2[rp] = pos ? xh : xh * al[k];  // act
                a_dp[k] += dpre;
                a_dpx[k] = fmaf(dpre, xh, a_dpx[k]);
                a_al[k] += pos ? 0.f : d * xh;
              } else {
                o2[rp] = rs[k] * (dpre - m1[k] - xh * m2[k]);
              }
            }
            af[hf * 2 + hi] = H::pack(o2[0], o2[1]);
            if (MODE == 1) {
              const float2 rr = H::unpack(af[hf * 2 + hi]);
              a_dp[k] += rr.x + rr.y;
            }
          }
        }
        if (MODE == 0) mma_k16_acc<BF16>(a_dw[mt], af, bdw[0], bdw[1]);
        else stsm_x4_trans(addr, af);  // dz back over the z tile (row-major [row][32 ch])
      }
      if (MODE == 1) {
        __syncwarp();
        // the group's 16 rows x 64 B are contiguous in the tile and in dz: 64 16-byte vectors, two per lane
        const uint4* src = reinterpret_cast<const uint4*>(zs + s * ZB + rb * CM * 2);
        uint4* dst = p.dz + ((long long)n * p.R + (long long)tt * TR + rb) * 4;
        dst[lane] = src[lane];
        dst[lane + 32] = src[lane + 32];
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
  p.wshift = 0;
  while ((1 << p.wshift) < W) ++p.wshift;
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
  cudaStream_t st = (cudaStream_t)stream;
  static const int use_mma = [] { const char* e = getenv("VB200_HEAD_BWD_MMA"); return e ? atoi(e) : 1; }();
  if (use_mma) {
    const void* mf[4] = {(const void*)hb::head_bwd_mma_kernel<true, 0>, (const void*)hb::head_bwd_mma_kernel<true, 1>,
                         (const void*)hb::head_bwd_mma_kernel<false, 0>, (const void*)hb::head_bwd_mma_kernel<false, 1>};
    static PerDeviceOnce monce[4];
    const size_t msmem = hb::MSTAGES * (hb::ZB + hb::DB) + 2 * hb::MSTAGES * sizeof(uint64_t) + 368 * sizeof(float);
    if (monce[slot].need(dev)) {
      cudaFuncSetAttribute(mf[slot], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
      monce[slot].done(dev);
    }
    const int mgrid = p.n_tiles < 2 * sms ? p.n_tiles : 2 * sms;  // two CTAs per SM
    if (slot == 0) hb::head_bwd_mma_kernel<true, 0><<<mgrid, hb::NCOMP, msmem, st>>>(p);
    else if (slot == 1) hb::head_bwd_mma_kernel<true, 1><<<mgrid, hb::NCOMP, msmem, st>>>(p);
    else if (slot == 2) hb::head_bwd_mma_kernel<false, 0><<<mgrid, hb::NCOMP, msmem, st>>>(p);
    else hb::head_bwd_mma_kernel<false, 1><<<mgrid, hb::NCOMP, msmem, st>>>(p);
    return check_launch("vb200_head_tail_bwd_stream");
  }
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  if (slot == 0) hb::head_bwd_stream_kernel<true, 0><<<grid, hb::NCOMP, smem, st>>>(p);
  else if (slot == 1) hb::head_bwd_stream_kernel<true, 1><<<grid, hb::NCOMP, smem, st>>>(p);
  else if (slot == 2) hb::head_bwd_stream_kernel<false, 0><<<grid, hb::NCOMP, smem, st>>>(p);
  else hb::head_bwd_stream_kernel<false, 1><<<grid, hb::NCOMP, smem, st>>>(p);
  return check_launch("vb200_head_tail_bwd_stream");
}
