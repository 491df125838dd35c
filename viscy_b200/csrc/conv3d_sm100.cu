// Implicit-GEMM 3x3x3 Conv3d (stride 1) on tcgen05 for small channel counts (Cin = 8 * CCH, CCH <= 4), channels-last.
//
// im2col is folded into TMA: the input is described by ONE 5-D tensor map (C, X, Y, Z, N) and every (kd, kh, channel
// chunk) "line" of a 128-voxel output row is fetched as a box of 8 channels x 130 voxels whose out-of-bounds part is
// zero-filled by the TMA unit (= the conv padding).  In shared memory a line is 130 rows of 16 bytes, which is exactly
// the SWIZZLE_NONE K-major core-matrix layout (8 rows x 16 B contiguous); the three kw taps of a line are the SAME
// buffer viewed through descriptors whose start address is shifted by kw rows (kw * 16 B), so each input byte crosses
// L2 -> SM once per (kd, kh) instead of once per tap.  Weights stay resident in shared memory for the CTA's lifetime.
//
//   conv3d_k3_kernel      : forward, and (with flipped/transposed weights + "full" padding) the data gradient
//   conv3d_k3_wgrad_kernel: weight gradient; both operands MN-major views of the same line buffers, the whole
//                           [Cout x 27*Cin] gradient accumulates in TMEM across all tiles of a persistent CTA
//
// Replaces nn.Conv3d(k=3) of monai Convolution in PixelToVoxelHead (VM/components/heads.py:607-628) and its autograd
// dgrad / wgrad (cuDNN today).  Same kernels carry small-channel 3x3x3 convs of Unet3d / Unet25d stems.
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vb {

int sm_count();  // gemm_sm100.cu

constexpr int LINE_ROWS = 130;              // 128 output voxels + 2 halo
constexpr int LINE_BYTES = LINE_ROWS * 16;  // 2080 (TMA box bytes)
constexpr int LINE_PITCH = 2176;            // 17 * 128: TMA destinations stay 128-B aligned

// SWIZZLE_NONE shared-memory matrix descriptor (layout type 0)
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

struct ConvParams {
  int N, D, H, W;      // input extent (voxels); channels = 8 * CCH
  int OD, OH, OW;      // output extent
  int pd, ph, pw;      // padding
  int co_store;        // output channels written per voxel (row pitch of out), multiple of 8
  int bf16;
  int xtiles;          // ceil(OW / 128)
  long long tiles;     // N * OD * OH * xtiles
  const uint4* wpack;  // fwd/dgrad: [KCH][CO][8] 16-bit, K order (kd, kh, chunk, kw), zero chunk appended when odd
  const float* bias;   // [co_store] or null
  void* out;           // fwd/dgrad: [N,OD,OH,OW,co_store] 16-bit;  wgrad: fp32 [CO][9][3*8] accumulated (atomics)
};

__device__ __forceinline__ void tile_coords(const ConvParams& p, int t, int& n, int& oz, int& oy, int& x0) {
  x0 = (t % p.xtiles) * 128;  // 32-bit: the host rejects problems with >= 2^31 tiles
  t /= p.xtiles;
  oy = t % p.OH;
  t /= p.OH;
  oz = t % p.OD;
  n = t / p.OD;
}

// ------------------------------------------------------------------------------------------------ forward / dgrad
template <int CCH, int CO>
struct ConvCfg {
  static constexpr int LINES = 9 * CCH;
  static constexpr int KCH = 27 * CCH + ((27 * CCH) & 1);  // K chunks of 8, padded to even
  static constexpr int STAGE_BYTES = LINES * LINE_PITCH;
  static constexpr int STAGES = CCH == 1 ? 4 : (CCH == 2 ? 3 : 2);
  static constexpr int W_BYTES = KCH * CO * 16;
  static constexpr int TMEM_COLS = (2 * CO) < 32 ? 32 : 2 * CO;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + LINE_PITCH /*zero chunk*/ + W_BYTES + 1024 + 256;
};

template <int CCH, int CO>
__global__ void __launch_bounds__(192, 1)
conv3d_k3_kernel(const __grid_constant__ CUtensorMap tmU, const ConvParams p) {
  using C = ConvCfg<CCH, CO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* zero_line = smem + C::STAGES * C::STAGE_BYTES;
  uint8_t* wsm = zero_line + LINE_PITCH;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wsm + C::W_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // weights + zero chunk -> shared memory (generic proxy), then make them visible to the tensor-core (async) proxy
  for (int i = threadIdx.x; i < C::W_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(wsm)[i] = __ldg(p.wpack + i);
  for (int i = threadIdx.x; i < LINE_PITCH / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(zero_line)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmU);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ===================== TMA producer: 9*CCH line boxes per tile
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < (int)p.tiles; t += gridDim.x) {
        int n, oz, oy, x0;
        tile_coords(p, t, n, oz, oy, x0);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], C::LINES * LINE_BYTES);
        uint8_t* sbase = smem + stage * C::STAGE_BYTES;
#pragma unroll
        for (int l = 0; l < C::LINES; ++l) {  // fully unrolled: tap / chunk indices are compile-time constants
          const int c = l % CCH, kh = (l / CCH) % 3, kd = l / (3 * CCH);
          tma_load_5d(sbase + l * LINE_PITCH, &tmU, &full_bar[stage], c * 8, x0 - p.pw, oy + kh - p.ph,
                      oz + kd - p.pd, n);
        }
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp walks the loop (uniform control flow keeps the descriptor
    // arithmetic on the uniform datapath), one lane elected once issues; descriptors = per-stage base + constants
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(128, CO, p.bf16 != 0, false, false);
    const uint32_t w_addr = smem_u32(wsm);
    const uint32_t z_addr = smem_u32(zero_line);
    const uint32_t s0 = smem_u32(smem);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < (int)p.tiles; t += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sbase = s0 + stage * C::STAGE_BYTES;
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * CO);
      // A: 8-row groups 128 B apart (SBO), the two K chunks of an MMA (a1 - a0) apart (LBO); B: chunks CO*16 B apart
      const uint64_t db0 = make_desc_noswz(w_addr, CO * 16, 128);
      if (leader) {
#pragma unroll
        for (int i = 0; i < C::KCH / 2; ++i) {  // fully unrolled: descriptor offsets fold to constants
          // K chunk q -> (line = q / 3, kw = q % 3) with K order (kd, kh, chunk, kw); the padding chunk reads zeros
          const int q0 = 2 * i, q1 = 2 * i + 1;
          const uint32_t o0 = (q0 / 3) * LINE_PITCH + (q0 % 3) * 16;
          uint64_t da;
          if (q1 < 27 * CCH) {
            const uint32_t o1 = (q1 / 3) * LINE_PITCH + (q1 % 3) * 16;
            da = make_desc_noswz(0, o1 - o0, 128) + ((sbase + o0) >> 4);
          } else {  // zero chunk sits above the stages
            da = make_desc_noswz(sbase + o0, z_addr - (sbase + o0), 128);
          }
          tc_mma_f16(d_tmem, da, db0 + ((q0 * (CO * 16)) >> 4), idesc, i > 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);
        tc_commit(&tmem_full[acc]);
      }
      __syncwarp();
      if (++stage == C::STAGES) {
        stage = 0;
        phase ^= 1;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {  // ===================== epilogue: 4 warps, thread = output voxel (row), CO columns
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool bf16 = p.bf16 != 0;
    for (int t = blockIdx.x; t < (int)p.tiles; t += gridDim.x) {
      int n, oz, oy, x0;
      tile_coords(p, t, n, oz, oy, x0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int xr = x0 + quarter * 32 + lane;
      const uint32_t t_addr =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * CO);
      uint32_t r[32];
      if constexpr (CO == 32) {
        tmem_ld32(t_addr, r);
      } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(t_addr)
            : "memory");
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);  // accumulator is in registers: release the buffer early
      if (xr < p.OW) {
        const long long row = (((long long)n * p.OD + oz) * p.OH + oy) * p.OW + xr;
        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + row * p.co_store;
#pragma unroll
        for (int g = 0; g < CO / 8; ++g) {
          if (g * 8 < p.co_store) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[j] = __uint_as_float(r[g * 8 + j]) + (p.bias != nullptr ? __ldg(p.bias + g * 8 + j) : 0.f);
            uint4 q;
            if (bf16) {
              q = make_uint4(H16<true>::pack(v[0], v[1]), H16<true>::pack(v[2], v[3]), H16<true>::pack(v[4], v[5]),
                             H16<true>::pack(v[6], v[7]));
            } else {
              q = make_uint4(H16<false>::pack(v[0], v[1]), H16<false>::pack(v[2], v[3]), H16<false>::pack(v[4], v[5]),
                             H16<false>::pack(v[6], v[7]));
            }
            *reinterpret_cast<uint4*>(o + g * 8) = q;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad (Cin = 8)
// D[co, (line, kw, ci)] += sum_{voxels x of the row} dz[x][co] * u[x + kw - pw][ci]   for the 9 (kd, kh) lines.
// A = dz tile viewed MN-major: chunks of 8 co (one TMA box of 128 rows x 16 B each), M padded to 128 with zero chunks.
// B = line buffer viewed MN-major: N = 32 = four kw-shifted views (chunk stride 16 B; the 4th is a don't-care that
// keeps N a multiple of 16 as UMMA M=128 requires), K = rows.
constexpr int WG_A_CHUNK = 2048;  // 128 rows x 16 B
template <int COCH>               // Cout = 8 * COCH real output channels
struct WgCfg {
  static constexpr int A_BYTES = 16 * WG_A_CHUNK;  // M = 128 = 16 chunks (upper ones stay zero)
  static constexpr int STAGE_BYTES = A_BYTES + 9 * LINE_PITCH;
  static constexpr int STAGES = 3;
  static constexpr int TMEM_COLS = 512;  // 9 lines x 32 columns = 288
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int COCH>
__global__ void __launch_bounds__(192, 1)
conv3d_k3_wgrad_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmDz,
                       const ConvParams p) {
  using C = WgCfg<COCH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* done_bar = empty_bar + C::STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the padding chunks of A in every stage once (the TMA boxes only ever overwrite the first COCH chunks)
  for (int s = 0; s < C::STAGES; ++s) {
    uint4* a = reinterpret_cast<uint4*>(smem + s * C::STAGE_BYTES + COCH * WG_A_CHUNK);
    for (int i = threadIdx.x; i < (16 - COCH) * WG_A_CHUNK / 16; i += blockDim.x) a[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmDz);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < (int)p.tiles; t += gridDim.x) {
        int n, oz, oy, x0;
        tile_coords(p, t, n, oz, oy, x0);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], COCH * WG_A_CHUNK + 9 * LINE_BYTES);
        uint8_t* sa = smem + stage * C::STAGE_BYTES;
        uint8_t* sl = sa + C::A_BYTES;
#pragma unroll
        for (int c = 0; c < COCH; ++c) tma_load_5d(sa + c * WG_A_CHUNK, &tmDz, &full_bar[stage], c * 8, x0, oy, oz, n);
#pragma unroll
        for (int l = 0; l < 9; ++l)
          tma_load_5d(sl + l * LINE_PITCH, &tmU, &full_bar[stage], 0, x0 - p.pw, oy + (l % 3) - p.ph,
                      oz + (l / 3) - p.pd, n);
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer (whole warp walks the loop, one lane elected once issues; see conv3d_k3_kernel)
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(128, 32, p.bf16 != 0, true, true);
    // MN-major, no swizzle: MN chunks SBO apart, 8-row K groups LBO = 128 B apart; 16 K rows = 256 B per step
    const uint64_t da0 = make_desc_noswz(0, 128, WG_A_CHUNK), db0 = make_desc_noswz(0, 128, 16);
    const uint32_t s0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int t = blockIdx.x; t < (int)p.tiles; t += gridDim.x) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = s0 + stage * C::STAGE_BYTES;
      const uint64_t da_s = da0 + (sa >> 4), db_s = db0 + ((sa + C::A_BYTES) >> 4);
      if (leader) {
#pragma unroll
        for (int l = 0; l < 9; ++l) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            tc_mma_f16(tmem_base + l * 32, da_s + ((ks * 256) >> 4), db_s + ((l * LINE_PITCH + ks * 256) >> 4), idesc,
                       (!first || ks > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);
      }
      __syncwarp();
      first = false;
      if (++stage == C::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (leader) tc_commit(done_bar);
    __syncwarp();
  } else {
    // epilogue after the last tile: rows (co) 0 .. 8*COCH-1 of the accumulator -> global atomics
    const int quarter = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int co = quarter * 32 + lane;
    if (quarter * 32 < 8 * COCH) {  // warp-uniform
      float* o = reinterpret_cast<float*>(p.out) + (long long)co * 216;
#pragma unroll 1
      for (int l = 0; l < 9; ++l) {
#pragma unroll 1
        for (int cb = 0; cb < 24; cb += 8) {
          uint32_t r[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + l * 32 + cb)
                       : "memory");
          tmem_ld_wait();
          if (co < 8 * COCH) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(o + l * 24 + cb + j, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// channels-last 5-D map (C, X, Y, Z, N), box = 8 channels x box_x voxels, no swizzle, zero OOB fill
static int make_tmap_ndhwc(CUtensorMap* m, const void* base, int N, int D, int H, int W, int Cc, int box_x, bool bf16) {
  static EncodeTiledFn enc = nullptr;
  if (enc == nullptr) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      return fail(VB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    enc = reinterpret_cast<EncodeTiledFn>(f);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || Cc % 8 != 0)
    return fail(VB200_ERR_UNSUPPORTED, "conv3d operand needs a 16-byte aligned base and C %% 8 == 0 (C=%d)", Cc);
  cuuint64_t dims[5] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2,
                           (cuuint64_t)D * H * W * Cc * 2};
  cuuint32_t box[5] = {8, (cuuint32_t)box_x, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VB200_ERR_CUDA, "cuTensorMapEncodeTiled (5-D) failed (%d)", (int)r);
  return VB200_OK;
}

template <int CCH, int CO>
static int launch_conv(const CUtensorMap& tm, const ConvParams& p, cudaStream_t st) {
  static PerDeviceOnce once;
  const int dev = PerDeviceOnce::device();
  auto kern = conv3d_k3_kernel<CCH, CO>;
  if (once.need(dev)) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<CCH, CO>::SMEM_BYTES);
    if (e != cudaSuccess) return fail(VB200_ERR_CUDA, "conv3d smem attribute: %s", cudaGetErrorString(e));
    once.done(dev);
  }
  const int grid = (int)(p.tiles < sm_count() ? p.tiles : sm_count());
  kern<<<grid, 192, ConvCfg<CCH, CO>::SMEM_BYTES, st>>>(tm, p);
  return check_launch("vb200_conv3d_k3");
}

template <int COCH>
static int launch_wgrad(const CUtensorMap& tmU, const CUtensorMap& tmDz, const ConvParams& p, cudaStream_t st) {
  static PerDeviceOnce once;
  const int dev = PerDeviceOnce::device();
  auto kern = conv3d_k3_wgrad_kernel<COCH>;
  if (once.need(dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<COCH>::SMEM_BYTES);
    if (e != cudaSuccess) return fail(VB200_ERR_CUDA, "conv3d wgrad smem attribute: %s", cudaGetErrorString(e));
    once.done(dev);
  }
  const int grid = (int)(p.tiles < sm_count() ? p.tiles : sm_count());
  kern<<<grid, 192, WgCfg<COCH>::SMEM_BYTES, st>>>(tmU, tmDz, p);
  return check_launch("vb200_conv3d_k3_wgrad");
}

}  // namespace vb

using namespace vb;

static int fill_params(ConvParams* p, const int32_t* g, int dtype) {
  p->N = g[0];
  p->D = g[1];
  p->H = g[2];
  p->W = g[3];
  p->pd = g[4];
  p->ph = g[5];
  p->pw = g[6];
  p->OD = p->D + 2 * p->pd - 2;
  p->OH = p->H + 2 * p->ph - 2;
  p->OW = p->W + 2 * p->pw - 2;
  if (p->OD <= 0 || p->OH <= 0 || p->OW <= 0) return fail(VB200_ERR_INVALID, "conv3d: empty output");
  if (dtype != VB200_BF16 && dtype != VB200_FP16) return fail(VB200_ERR_UNSUPPORTED, "dtype %d", dtype);
  p->bf16 = dtype == VB200_BF16;
  p->xtiles = (p->OW + 127) / 128;
  p->tiles = (long long)p->N * p->OD * p->OH * p->xtiles;
  if (p->tiles >= (1LL << 31)) return fail(VB200_ERR_UNSUPPORTED, "conv3d: too many tiles");
  return VB200_OK;
}

extern "C" int vb200_conv3d_k3(const void* u, const void* wpack, const float* bias, void* out, const int32_t* geom,
                               int cin, int cout_pad, int co_store, int dtype, vb200_stream_t stream) {
  VB_REQUIRE(u && wpack && out && geom, "null pointer");
  ConvParams p;
  if (int rc = fill_params(&p, geom, dtype)) return rc;
  VB_SUPPORTED(co_store % 8 == 0 && co_store <= cout_pad, "co_store %d", co_store);
  p.co_store = co_store;
  p.wpack = reinterpret_cast<const uint4*>(wpack);
  p.bias = bias;
  p.out = out;
  CUtensorMap tm;
  if (int rc = make_tmap_ndhwc(&tm, u, p.N, p.D, p.H, p.W, cin, LINE_ROWS, p.bf16)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (cin == 8 && cout_pad == 32) return launch_conv<1, 32>(tm, p, st);
  if (cin == 8 && cout_pad == 16) return launch_conv<1, 16>(tm, p, st);
  if (cin == 16 && cout_pad == 32) return launch_conv<2, 32>(tm, p, st);
  if (cin == 16 && cout_pad == 16) return launch_conv<2, 16>(tm, p, st);
  if (cin == 32 && cout_pad == 32) return launch_conv<4, 32>(tm, p, st);
  if (cin == 32 && cout_pad == 16) return launch_conv<4, 16>(tm, p, st);
  return fail(VB200_ERR_UNSUPPORTED, "conv3d_k3: cin %d / cout_pad %d (supported: cin 8|16|32, cout_pad 16|32)", cin,
              cout_pad);
}

extern "C" int vb200_conv3d_k3_wgrad(const void* u, const void* dz, float* dw, const int32_t* geom, int cin, int cout,
                                     int dtype, vb200_stream_t stream) {
  VB_REQUIRE(u && dz && dw && geom, "null pointer");
  VB_SUPPORTED(cin == 8, "conv3d_k3_wgrad: cin %d (only 8)", cin);
  ConvParams p;
  if (int rc = fill_params(&p, geom, dtype)) return rc;
  p.co_store = cout;
  p.wpack = nullptr;
  p.bias = nullptr;
  p.out = dw;
  CUtensorMap tmU, tmDz;
  if (int rc = make_tmap_ndhwc(&tmU, u, p.N, p.D, p.H, p.W, cin, LINE_ROWS, p.bf16)) return rc;
  if (int rc = make_tmap_ndhwc(&tmDz, dz, p.N, p.OD, p.OH, p.OW, cout, 128, p.bf16)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (cout == 32) return launch_wgrad<4>(tmU, tmDz, p, st);
  if (cout == 16) return launch_wgrad<2>(tmU, tmDz, p, st);
  if (cout == 8) return launch_wgrad<1>(tmU, tmDz, p, st);
  return fail(VB200_ERR_UNSUPPORTED, "conv3d_k3_wgrad: cout %d (8|16|32)", cout);
}
