// Weight gradient of stride-1 3-D convolutions with few channels (Cin <= 64 per tile), tcgen05, channels-last.
//
//   dw[co, (kd,kh,kw), ci] += sum_v dout[v, co] * x[v + tap - pad, ci]
//
// The generic implicit-GEMM weight gradient (gemm_sm100.cu, MODE_CONVMN) fetches one box of dout and one of x per
// (tap, 64 voxels); with <= 64 channels a box row carries 64-128 bytes and the kernel is bound by the number of box rows
// the TMA unit can deliver, not by the tensor pipe.  Here a K block is an 8 x 8 voxel patch of one z-plane:
//   A = x patch with y halo   : box (64 ci, 8, 11, 1, 1) -> 88 rows, MN-major SWIZZLE_128B; a kh tap is the SAME box
//                               viewed 8 rows further down (shifts by 8 rows keep the 128-byte swizzle phase), and the two
//                               64-wide M atoms of one MMA are two consecutive kh views (LBO = one voxel row = 1024 B):
//                               M = 128 = (kh, kh+1) x 64 input channels
//   B = dout patch            : box (64 co, 8, 8, 1, 1)  -> 64 rows; N = 32 or 64 output channels
// so one unit = (co tile, kd, kw, ci tile, K split) accumulates the three kh taps in two TMEM accumulators ((0,1) and
// (2, unused)) from 152 box rows and 8 MMAs per 64 voxels instead of 3 x 128 rows and 12 MMAs with 3/4 of the M rows
// padding: measured, the tensor pipe waits on shared-memory operand reads, so bytes per useful MAC is what counts.  Replaces the autograd weight gradient of nn.Conv3d(k=3, padding=1) in Block.proj /
// ResnetBlock (VM/unet/blocks.py:88-188) and ConvBlock3D (VM/components/conv_block_3d.py:261-274) (cuDNN today).
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vb {

int sm_count();  // gemm_sm100.cu

namespace wg3 {

constexpr int A_BYTES = 88 * 128;       // x: 64 channels x (8 x 11) voxel rows
constexpr int B_BYTES = 64 * 128;       // dout: 64 channels x (8 x 8) voxel rows
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 19 KB
constexpr int STAGES = 10;
constexpr int TMEM_COLS = 256;          // 2 buffers x 2 accumulators x 64 columns
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

struct Params {
  int N, OD, OH, OW;      // output extent (= extent of dout)
  int KD, KH, KW;         // filter extent (KH == 3)
  int pd, ph, pw;
  int cin, cout;          // channels of x / dout (row pitches)
  int ctiles;             // ceil(cin / 64)
  int tiles_m;            // ceil(cout / 64): output-channel tiles (the MMA N)
  int px_n, py_n;         // patches per row / column
  long long patches;      // N * OD * py_n * px_n  (K blocks)
  int k_splits, kb_per_split;
  int bf16;
  float* dw;              // [cout, KD*KH*KW*cin] fp32, accumulated
};

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// unit -> (m tile, kd, kw, channel tile, split)
struct Unit {
  int m0, kd, kw, ci0, split;
};
__device__ __forceinline__ Unit decode(const Params& p, int unit) {
  Unit u;
  const int tiles = p.tiles_m * p.KD * p.KW * p.ctiles;
  u.split = unit / tiles;
  int t = unit - u.split * tiles;
  u.ci0 = (t % p.ctiles) * 64;
  t /= p.ctiles;
  u.kw = t % p.KW;
  t /= p.KW;
  u.kd = t % p.KD;
  u.m0 = (t / p.KD) * 64;  // first output channel of the tile
  return u;
}

__global__ void __launch_bounds__(192, 1)
conv3d_wgrad_kh3_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmX,
                        const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDz);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < STAGES; ++i) {
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
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int units = p.tiles_m * p.KD * p.KW * p.ctiles * p.k_splits;

  if (warp == 0) {
    if (lane == 0) {  // ===================== TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const Unit u = decode(p, unit);
        const long long kb0 = (long long)u.split * p.kb_per_split;
        const long long kb1 = min(p.patches, kb0 + p.kb_per_split);
        // patch walk: (n, z, py, px) with px fastest; decomposed once, then advanced with carries
        long long t = kb0;
        int px = (int)(t % p.px_n);
        t /= p.px_n;
        int py = (int)(t % p.py_n);
        t /= p.py_n;
        int z = (int)(t % p.OD);
        int n = (int)(t / p.OD);
        for (long long kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_5d(sa, &tmX, &full_bar[stage], u.ci0, px * 8 + u.kw - p.pw, py * 8 - p.ph, z + u.kd - p.pd, n);
          tma_load_5d(sb, &tmDz, &full_bar[stage], u.m0, px * 8, py * 8, z, n);
          if (++px == p.px_n) {
            px = 0;
            if (++py == p.py_n) {
              py = 0;
              if (++z == p.OD) {
                z = 0;
                ++n;
              }
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 2 kh pairs x 4 K steps per K block.  The whole warp walks the loops (uniform
    // control flow, uniform-datapath address arithmetic); one lane, elected once, issues tcgen05.mma / commit.
    // MN-major SW128 descriptors: 8-row K groups 1024 B apart (SBO), 16 K rows = 2048 B per step.  A: the second 64-wide
    // M atom is the next kh view = 8 voxel rows = 1024 B further (LBO); pair 1's second atom is unused.
    const bool leader = elect_one();
    const uint64_t desc_a0 = make_smem_desc(0, 1024, 1024), desc_b0 = make_smem_desc(0, 0, 1024);
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const Unit u = decode(p, unit);
      const long long kb0 = (long long)u.split * p.kb_per_split;
      const long long kb1 = min(p.patches, kb0 + p.kb_per_split);
      const int ncols = p.cout - u.m0 <= 32 ? 32 : 64;  // MMA N: output channels of this tile
      const uint32_t idesc = make_idesc(128, ncols, p.bf16 != 0, true, true);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + static_cast<uint32_t>(acc * 128);
      for (long long kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem0 + stage * STAGE_BYTES;
        const uint64_t da_s = desc_a0 + (sa >> 4), db_s = desc_b0 + ((sa + A_BYTES) >> 4);
        if (leader) {
#pragma unroll
          for (int pair = 0; pair < 2; ++pair) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16(d0 + pair * 64, da_s + ((pair * 2048 + k * 2048) >> 4), db_s + ((k * 2048) >> 4), idesc,
                         (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) tc_commit(&tmem_full[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {  // ===================== epilogue: 4 warps; lane = (kh of the pair, input channel), columns = output channels
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const long long taps_row = (long long)p.KD * p.KH * p.KW * p.cin;  // row pitch of dw
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const Unit u = decode(p, unit);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int ci = u.ci0 + (quarter & 1) * 32 + lane;
      const int nco = min(64, p.cout - u.m0);
#pragma unroll 1
      for (int pair = 0; pair < 2; ++pair) {
        const int kh = pair * 2 + (quarter >> 1);
        const bool live = kh < 3 && u.ci0 + (quarter & 1) * 32 < p.cin;  // warp-uniform
        const int tap = (u.kd * p.KH + kh) * p.KW + u.kw;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          if (live && c * 32 < nco) {
            uint32_t r[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 128 + pair * 64 + c * 32, r);
            tmem_ld_wait();
            if (ci < p.cin) {
              float* o = p.dw + (long long)(u.m0 + c * 32) * taps_row + (long long)tap * p.cin + ci;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c * 32 + j < nco) atomicAdd(o + j * taps_row, __uint_as_float(r[j]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// channels-last [N,D,H,W,C] as (C, X, Y, Z, N); box = 64 channels x (8, by, 1, 1) voxels, SWIZZLE_128B, zero OOB fill
static int make_tmap_patch(CUtensorMap* m, const void* base, int N, int D, int H, int W, int Cc, int by, bool bf16) {
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
    return fail(VB200_ERR_UNSUPPORTED, "conv operand needs a 16-byte aligned base and C %% 8 == 0 (C=%d)", Cc);
  cuuint64_t dims[5] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2,
                           (cuuint64_t)D * H * W * Cc * 2};
  cuuint32_t box[5] = {64, 8, (cuuint32_t)by, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VB200_ERR_CUDA, "cuTensorMapEncodeTiled (patch box) failed (%d)", (int)r);
  return VB200_OK;
}

}  // namespace wg3
}  // namespace vb

using namespace vb;

// 1 when the patch form applies: stride 1, kh == 3, output rows / columns in whole 8 x 8 patches
extern "C" int vb200_conv3d_wgrad_kh3_supported(const vb200_conv3d_desc* d) {
  if (d == nullptr || d->kh != 3 || d->kd < 1 || d->kw < 1) return 0;
  if ((d->sd > 1) || (d->sh > 1) || (d->sw > 1)) return 0;
  if (d->cin % 8 != 0 || d->cout % 8 != 0 || d->cin <= 0 || d->cout <= 0) return 0;
  if (d->dtype != VB200_BF16 && d->dtype != VB200_FP16) return 0;
  const int OD = d->D + 2 * d->pd - d->kd + 1, OH = d->H + 2 * d->ph - 2, OW = d->W + 2 * d->pw - d->kw + 1;
  if (OD <= 0 || OH <= 0 || OW <= 0 || OH % 8 != 0 || OW % 8 != 0) return 0;
  return 1;
}

extern "C" int vb200_conv3d_wgrad_kh3(const vb200_conv3d_desc* d, vb200_stream_t stream) {
  VB_REQUIRE(d != nullptr && d->x && d->dout && d->dw, "null pointer");
  VB_SUPPORTED(vb200_conv3d_wgrad_kh3_supported(d), "conv3d_wgrad_kh3: needs stride 1, kh == 3 and an output extent in 8 x 8 patches");
  wg3::Params p{};
  p.N = d->N;
  p.OD = d->D + 2 * d->pd - d->kd + 1;
  p.OH = d->H + 2 * d->ph - 2;
  p.OW = d->W + 2 * d->pw - d->kw + 1;
  p.KD = d->kd; p.KH = 3; p.KW = d->kw;
  p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
  p.cin = d->cin; p.cout = d->cout;
  p.ctiles = (d->cin + 63) / 64;
  p.tiles_m = (d->cout + 63) / 64;
  p.px_n = p.OW / 8; p.py_n = p.OH / 8;
  p.patches = (long long)p.N * p.OD * p.py_n * p.px_n;
  p.bf16 = d->dtype == VB200_BF16;
  p.dw = d->dw;
  const int sms = sm_count();
  const long long tiles = (long long)p.tiles_m * p.KD * p.KW * p.ctiles;
  long long splits = d->k_splits > 0 ? d->k_splits : (2LL * sms) / tiles;  // whole waves (see vb200_conv3d_igemm_wgrad)
  if (splits > p.patches / 8) splits = p.patches / 8;
  if (splits < 1) splits = 1;
  p.kb_per_split = (int)((p.patches + splits - 1) / splits);
  p.k_splits = (int)((p.patches + p.kb_per_split - 1) / p.kb_per_split);
  VB_SUPPORTED(tiles * p.k_splits < (1LL << 30), "conv3d_wgrad_kh3: too many units");
  CUtensorMap tmDz, tmX;
  if (int rc = wg3::make_tmap_patch(&tmDz, d->dout, p.N, p.OD, p.OH, p.OW, d->cout, 8, p.bf16 != 0)) return rc;
  if (int rc = wg3::make_tmap_patch(&tmX, d->x, d->N, d->D, d->H, d->W, d->cin, 11, p.bf16 != 0)) return rc;
  static PerDeviceOnce once;
  const int dev = PerDeviceOnce::device();
  if (once.need(dev)) {
    cudaError_t e = cudaFuncSetAttribute(wg3::conv3d_wgrad_kh3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         wg3::SMEM_BYTES);
    if (e != cudaSuccess) return fail(VB200_ERR_CUDA, "conv3d_wgrad_kh3 smem attribute: %s", cudaGetErrorString(e));
    once.done(dev);
  }
  const long long units = tiles * p.k_splits;
  const int grid = (int)(units < sms ? units : sms);
  wg3::conv3d_wgrad_kh3_kernel<<<grid, 192, wg3::SMEM_BYTES, (cudaStream_t)stream>>>(tmDz, tmX, p);
  return check_launch("vb200_conv3d_wgrad_kh3");
}
