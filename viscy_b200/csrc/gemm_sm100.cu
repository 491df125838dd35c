// tcgen05 GEMM for sm_100a: persistent, warp-specialised (TMA producer / single-thread MMA issuer /
// 8 epilogue warps), 128 x BN x 64 tiles, 4-stage (BN=256) mbarrier ring in shared memory, two
// accumulator buffers in TMEM so that the drain of tile i overlaps the MMAs of tile i+1.
//
// Two operand forms (include/viscy_b200.h : vb200_gemm_desc):
//   mn_major = 0 : D[M,N] = A[M,K] . B[N,K]^T   (forward 1x1-conv / Linear, and dgrad)
//   mn_major = 1 : D[M,N] = sum_k At[k,M] * Bt[k,N]  (wgrad; k runs over pixels, K-split with red.add)
// Replaces the cuBLAS / cuDNN calls behind nn.Linear / 1x1 nn.Conv2d / kernel==stride convs of the
// reference hot path (VM/components/blocks.py:60-74, VM/components/stems.py:26-50).
#include <cuda.h>

#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace vb {

thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = 128 B = one SWIZZLE_128B row
// Epilogue warps per CTA: 8 (two per TMEM lane quarter, half of the tile's columns each), or 16 for the epilogues that
// are bound by their own instruction latency (GELU'/GELU pair, fused GRN+GELU backward) on 256-wide tiles: twice the
// warps in flight, each with a quarter of the columns, single-buffered operands to stay under the register ceiling.
template <int BN, int EPI>
struct EpiWarps {
  static constexpr int value = (BN == 256 && (EPI == VB200_EPI_GELU_GP || EPI == VB200_EPI_DGELU_GRN)) ? 16 : 8;
  // dual-output epilogue: results leave through TMA stores (two staging tiles, two bulk groups in flight per warp)
  static constexpr bool tma_store = value == 16 && EPI == VB200_EPI_GELU_GP;
  // fused GRN+GELU backward on CTA pairs: g and gp chunks land in per-warp tiles by TMA, one chunk ahead of the math
  static constexpr bool aux_tma(bool pair) { return pair && value == 16 && EPI == VB200_EPI_DGELU_GRN; }
};

// Operand forms.  The two CONV forms are the implicit-GEMM 3-D convolution: one operand is the channels-last activation
// tensor seen through a 5-D tensor map (C, X, Y, Z, N); a K block (CONVK) / an N tile (CONVMN) belongs to one filter tap
// and its box is fetched at the tap-shifted voxel coordinate, TMA zero-filling whatever falls outside (= the padding).
// MODE_CONVKP is the patch variant of CONVK for 3-row filters (kh == 3): the M tile is a 16 x 8 voxel patch of one
// z-plane, a K block = (kd, kw, channel chunk) fetches ONE box of 16 x 10 voxels (y halo) and the three kh taps are
// that box viewed at row offsets 0 / 16 / 32 (swizzle phase survives shifts by multiples of 8 rows), each against its
// own weight sub-tile: 160 + 3 BN box rows per three taps instead of 3 (128 + BN).
// MODE_CONVKPW: CONVKP with the whole filter resident in shared memory (fetched once per CTA, <= 110 KB: Cin x Cout up to
// 64 x 32 / 32 x 64 at 27 taps), so that a K block is the 160 A rows only -- the weight sub-tiles were over half of the
// box rows the TMA unit had to deliver per K block.
// MODE_KMAJOR2: the K-major form on a CTA pair (cluster of two CTAs, tcgen05 cta_group::2): one MMA covers a 256 x BN tile,
// 128 rows per CTA, each CTA stages only its half of the B rows - half the B traffic into and out of shared memory per SM.
enum { MODE_KMAJOR = 0, MODE_MNMAJOR = 1, MODE_CONVK = 2, MODE_CONVMN = 3, MODE_CONVKP = 4, MODE_CONVKPW = 5, MODE_KMAJOR2 = 6, MODE_CONVKP2 = 7 };  // CONVKP2: the patch conv form on CTA pairs

// BKE = K elements per stage: 64 (SWIZZLE_128B rows) or, for 32-channel conv operands, 32 (SWIZZLE_64B rows)
// AUXT: the 16-bit epilogue operands (g, gp of the fused GRN+GELU backward) arrive through TMA into two more per-warp tiles
template <int BN, int BKE = BK, int EW = 8, bool TWO_TILES = false, bool PATCH = false, bool WRES = false, bool PAIR = false,
          bool AUXT = false>
struct Cfg {
  static constexpr int A_BYTES = (PATCH ? 160 : BM) * BKE * 2;
  static constexpr int B_TAP_BYTES = (PAIR ? BN / 2 : BN) * BKE * 2;
  static constexpr int B_BYTES = WRES ? 0 : (PATCH ? 3 : 1) * B_TAP_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // (the 16-warp epilogues need 32 KB of staging tiles: one operand stage less on the 256-wide tile)
  static constexpr int STAGES =
      WRES  ? (BKE == 64 ? 4 : 8)
      : PATCH ? (PAIR ? (BN == 64 ? 5 : 4) : (BKE == 64 ? (BN == 64 ? 4 : 3) : (BN == 64 ? 8 : 5)))
      : PAIR ? (EW == 16 ? (AUXT ? 3 : 4) : 6)
            : (BKE == 64 ? ((BN == 256) ? (EW == 16 ? 3 : 4) : (BN == 128 ? 6 : 8)) : ((BN == 256) ? 8 : (BN == 128 ? 10 : 12)));
  static_assert(!PATCH || BN <= 128, "patch conv form: tiles up to 128 output channels");
  static constexpr int TMEM_COLS = 2 * BN;  // 512 / 256 / 128: powers of two >= 32
  // per epilogue warp: [bias, s, t][its BN / (EW / 4) columns] fp32 (private: no CTA-wide barrier per tile)
  // (the patch conv forms carry a bias only and sit at the shared-memory limit: one vector)
  static constexpr int COLV_BYTES = EW * (PATCH ? 1 : 3) * (BN / (EW / 4)) * 4;
  // per-warp 32 rows x 64 B staging tile(s): one, or two for the 16-warp epilogues whose outputs leave through TMA stores
  static constexpr int STG_PER_WARP = AUXT ? 6144 : (TWO_TILES ? 4096 : 2048);
  static_assert(!AUXT || (EW == 16 && PAIR && STAGES <= 6), "TMA-fed epilogue operands: the 16-warp CTA-pair form");
  static constexpr int STG_BYTES = EW * STG_PER_WARP;
  static_assert(!TWO_TILES || EW == 16, "TMA-store epilogues run 16 warps");
  // The staging tiles that TMA stores read (SWIZZLE_64B output maps) must sit on the swizzle period: the hardware
  // derives the XOR pattern from absolute shared-memory address bits, the epilogue warps from the row index.
  static constexpr int STG_OFFSET = (STAGES * STAGE_BYTES + 256 + COLV_BYTES + 1023) / 1024 * 1024;
  static constexpr int SMEM_BYTES = STG_OFFSET + 1024 /*align slack*/ + STG_BYTES;
  // WRES: the resident filter follows (1024-byte aligned), its size is a launch parameter
  static constexpr int W_OFFSET = (STG_OFFSET + STG_BYTES + 1023) / 1024 * 1024;
};

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, k_splits, kb_total, kb_per_split;
  int bf16;
  int act;
  int atomic_out;
  int b_batch_rows;     // > 0: B is [nb][N][K]; tile rows [m0, m0+128) use batch m0 / b_batch_rows
  int rows_per_sample;  // EPI_DGELU_GRN: n = row / rows_per_sample
  int n_split;          // EPI_F32: columns >= n_split (multiple of 16) go to out2 (fp32, row pitch ldo2) at column - n_split
  long long ldo, ldo2, ldr, ldaux, ldaux2;
  long long split_out_stride;
  void* out;
  void* out2;
  const float* bias;
  const void* residual;
  const void* aux;
  const void* aux2;
  const float* tvec;
  const float* svec;
  const float* rvec;  // EPI_STORE: per-sample scale of (acc + bias) * s before the residual: rvec[row / rvec_rows]
  int rvec_rows;
  float* colsq;       // EPI_GELU_GP: [M / rows_per_sample, N] += column sums of out2^2 per sample
  // implicit-GEMM conv forms: output extent, filter extent, padding, channel chunks per tap (CONVK) / channel tiles per
  // tap (CONVMN), channels of the activation operand
  int cOW, cOH, cOD, cKW, cKH, cpw, cph, cpd, cchunks, ccin;
  int cbx, cby, cbz, cbn;  // voxel box of one K block (CONVMN): the walk along K advances by one box, no divisions
  int csw, csh, csd;       // stride: input coordinate = output coordinate * stride + tap - padding
  // CONVK output scatter (transposed convs as parity-class sub-convolutions): when opx != 0 output voxel (n, z, y, x)
  // of this launch lives at out + n*opn + z*opz + y*opy + x*opx (elements) instead of at row m of a dense matrix
  long long opx, opy, opz, opn;
  // CONVK tap selection: K block tap t reads weight columns [ctap[t]*cin, (ctap[t]+1)*cin) (ntap == 0: t itself)
  int ntap;
  int ctap[27];
  int cpxn, cpyn;  // CONVKP: 16 x 8 voxel patches per output row / column
  int cwbytes;     // CONVKPW: bytes of the resident filter (taps x chunks sub-tiles of cwrows x BKE)
  int cwrows;      // CONVKPW: output channels per sub-tile = MMA N (32 or 64)
};

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// Output tensor maps of the TMA-store epilogues: [M, N] 16-bit, box 32 columns x 32 rows, SWIZZLE_64B -- the layout the
// epilogue warps already write their staging tile in (16-byte group g of row r at r*64 + ((g ^ (r>>1 & 3)) << 4)).
struct OutMaps {
  CUtensorMap o, o2;
};
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
// one elected lane owns the warp's bulk-store groups: wait until at most PENDING of them still read shared memory,
// publish the other lanes' staging writes to the async proxy, issue the store
template <int PENDING>
__device__ __forceinline__ void tma_stage_store(uint32_t tile, int lane, const uint4* vals, const CUtensorMap* m, int col0,
                                                long long row0) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
  __syncwarp();
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + lane * 64 + ((g ^ sw) << 4)), "r"(vals[g].x),
                 "r"(vals[g].y), "r"(vals[g].z), "r"(vals[g].w)
                 : "memory");
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(m, tile, col0, (int)row0);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// flattened output-voxel index -> (x, y, z, n)
__device__ __forceinline__ void voxel_coords(const GemmParams& p, int v, int& x, int& y, int& z, int& n) {
  x = v % p.cOW;
  v /= p.cOW;
  y = v % p.cOH;
  v /= p.cOH;
  z = v % p.cOD;
  n = v / p.cOD;
}

template <bool BF16>
__device__ __forceinline__ void load8(const void* p, float* v) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  float2 a = H16<BF16>::unpack(q.x), b = H16<BF16>::unpack(q.y), c = H16<BF16>::unpack(q.z),
         d = H16<BF16>::unpack(q.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <bool BF16>
__device__ __forceinline__ void store8(void* p, const float* v) {
  uint4 q;
  q.x = H16<BF16>::pack(v[0], v[1]);
  q.y = H16<BF16>::pack(v[2], v[3]);
  q.z = H16<BF16>::pack(v[4], v[5]);
  q.w = H16<BF16>::pack(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = q;
}

template <bool BF16>
__device__ __forceinline__ void unpack8(const uint4& q, float* v) {
  const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = H16<BF16>::unpack(w4[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
template <bool BF16>
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 q;
  q.x = H16<BF16>::pack(v[0], v[1]); q.y = H16<BF16>::pack(v[2], v[3]);
  q.z = H16<BF16>::pack(v[4], v[5]); q.w = H16<BF16>::pack(v[6], v[7]);
  return q;
}

// Auxiliary 16-bit operands of one 32-column chunk of this warp's 32 rows, fetched ahead of the accumulator drain.
// Loads are row-contiguous: lane l fetches, for i = 0..3, the 16-byte group (l & 3) of row (l >> 2) + 8 i, so one
// instruction covers 8 rows x 64 B; `unstage` transposes through the warp's staging tile into the row-per-lane order
// of the TMEM accumulator (lane = row, 4 groups of 8 columns).
struct AuxRegs {
  uint4 a[4];  // residual (EPI_STORE) | u (EPI_DGELU) | g (EPI_DGELU_GRN)
  uint4 b[4];  // gp (EPI_DGELU_GRN)
};

template <int EPI>
__device__ __forceinline__ const uint16_t* aux_a_ptr(const GemmParams& p, long long& lda) {
  const uint16_t* pa = nullptr;
  lda = 0;
  if constexpr (EPI == VB200_EPI_STORE) { pa = reinterpret_cast<const uint16_t*>(p.residual); lda = p.ldr; }
  if constexpr (EPI == VB200_EPI_DGELU) { pa = reinterpret_cast<const uint16_t*>(p.aux); lda = p.ldaux; }
  if constexpr (EPI == VB200_EPI_DGELU_GRN) {
    if (p.tvec != nullptr) { pa = reinterpret_cast<const uint16_t*>(p.aux); lda = p.ldaux; }
  }
  return pa;
}

template <int EPI>
__device__ __forceinline__ void load_aux(const GemmParams& p, long long row0, int col0, int lane, AuxRegs& x) {
  long long lda;
  const uint16_t* pa = aux_a_ptr<EPI>(p, lda);
  const int col = col0 + (lane & 3) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = row0 + (lane >> 2) + 8 * i;
    const bool ok = row < p.M && col < p.N;
    x.a[i] = make_uint4(0, 0, 0, 0);
    if (pa != nullptr && ok) x.a[i] = __ldg(reinterpret_cast<const uint4*>(pa + row * lda + col));
    if constexpr (EPI == VB200_EPI_DGELU_GRN) {
      x.b[i] = make_uint4(0, 0, 0, 0);
      if (ok) x.b[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.aux2) + row * p.ldaux2 + col));
    }
  }
}

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void lds_f8(uint32_t addr, float* f) {  // 8 consecutive fp32 (32-byte aligned)
  const uint4 a = lds128(addr), b = lds128(addr + 16);
  f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
  f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
}

// (8 rows x 4 groups per register) -> (row = lane, group g in register g), through the 32 x 64 B staging tile
// (stg = shared-space address of this warp's tile)
__device__ __forceinline__ void unstage(uint32_t stg, int lane, uint4* v) {
  __syncwarp();
  const int ch = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = (lane >> 2) + 8 * i;
    sts128(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4), v[i]);
  }
  __syncwarp();
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) v[g] = lds128(stg + lane * 64 + ((g ^ sw) << 4));
}

// Store one 32-row x 64-byte tile (this warp's rows, 16 B per lane and group) through a swizzled shared-memory
// staging tile so that global stores are row-contiguous: 4 lanes cover one row's 64 B, one instruction covers 8 rows.
// `vals` = this lane's row: 4 x uint4.  ATOMIC: red.add.v4.f32 instead of a store (fp32 data).
// `rowoff` (optional): byte offsets of this lane's four rows (lane >> 2) + 8 i relative to gbase, for scattered outputs.
template <bool ATOMIC>
__device__ __forceinline__ void stage_store(uint32_t stg, int lane, const uint4* vals, void* gbase, long long ld_bytes,
                                            long long row0, int rows_valid, long long col_byte0, int cols16_valid,
                                            const long long* rowoff = nullptr) {
  if (vals != nullptr) {  // (nullptr: the caller has written the tile already, behind a __syncwarp of its own)
    __syncwarp();
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int g = 0; g < 4; ++g) sts128(stg + lane * 64 + ((g ^ sw) << 4), vals[g]);
  }
  __syncwarp();
  const int ch = lane & 3;
  if (rowoff != nullptr) {
    if (ch < cols16_valid) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = (lane >> 2) + 8 * i;
        if (rr < rows_valid) {
          const uint4 q = lds128(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(gbase) + rowoff[i] + col_byte0 + ch * 16) = q;
        }
      }
    }
    return;
  }
  uint8_t* dst0 = reinterpret_cast<uint8_t*>(gbase) + (row0 + (lane >> 2)) * ld_bytes + col_byte0 + ch * 16;
  if (!ATOMIC && rows_valid == 32 && cols16_valid == 4) {  // interior tile: no per-row predicates
    uint4 q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = (lane >> 2) + 8 * i;
      q[i] = lds128(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dst0 + 8 * i * ld_bytes) = q[i];
    return;
  }
  if (ch < cols16_valid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = (lane >> 2) + 8 * i;
      if (rr < rows_valid) {
        const uint4 q = lds128(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
        uint8_t* dst = dst0 + 8 * i * ld_bytes;
        if constexpr (ATOMIC) {
          atomicAdd(reinterpret_cast<float4*>(dst), make_float4(__uint_as_float(q.x), __uint_as_float(q.y),
                                                                 __uint_as_float(q.z), __uint_as_float(q.w)));
        } else {
          *reinterpret_cast<uint4*>(dst) = q;
        }
      }
    }
  }
}

// Epilogue math for 8 consecutive columns of one row.  cv = shared-space address of the staged [bias | s | t] of the
// tile (fp32).  v: accumulators in, primary result out; w: secondary result (dual-output epilogues).
template <int EPI, int BN, bool BF16, int CVP>
__device__ __forceinline__ void epilogue_math8(const GemmParams& p, float* v, float* w, int cl, uint32_t cv,
                                               const uint4& xa, const uint4& xb, float rs) {
  if constexpr (EPI != VB200_EPI_DGELU_GRN) {  // (a data gradient carries no bias)
    float b8[8];
    lds_f8(cv + cl * 4, b8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += b8[j];
  }
  if constexpr (EPI == VB200_EPI_STORE) {
    if (p.svec != nullptr) {  // per-column fp32 scale (ConvNeXt-V1 layer scale gamma): (acc + bias) * s
      float s8[8];
      lds_f8(cv + (CVP + cl) * 4, s8);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= s8[j];
    }
    if (p.act == VB200_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
    } else if (p.act == VB200_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gelu_f(v[j]);
    }
    if (p.rvec != nullptr) {  // stochastic depth: the residual branch scaled per sample (row)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= rs;
    }
    if (p.residual != nullptr) {
      float q[8];
      unpack8<BF16>(xa, q);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += q[j];
    }
  } else if constexpr (EPI == VB200_EPI_GELU_DUAL) {
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = gelu_f(v[j]);
  } else if constexpr (EPI == VB200_EPI_GELU_GP) {
    // v = gelu'(u), w = gelu(u): the backward never needs u itself
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float2 d, gl;
      gelu_gp2(make_float2(v[j], v[j + 1]), d, gl);
      v[j] = d.x; v[j + 1] = d.y;
      w[j] = gl.x; w[j + 1] = gl.y;
    }
  } else if constexpr (EPI == VB200_EPI_DGELU) {
    float u[8];
    unpack8<BF16>(xa, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= dgelu_f(u[j]);
  } else if constexpr (EPI == VB200_EPI_DGELU_GRN) {
    // dh = (acc * s[n,col] + g * t[n,col]) * gp:  g = aux, gp = aux2 = gelu'(u) saved by the forward epilogue
    float g[8], gp[8], s8[8], t8[8];
    unpack8<BF16>(xa, g);
    unpack8<BF16>(xb, gp);
    lds_f8(cv + (CVP + cl) * 4, s8);
    lds_f8(cv + (2 * CVP + cl) * 4, t8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(g[j], t8[j], v[j] * s8[j]) * gp[j];
  }
}

__device__ __forceinline__ void lds_f2x4(uint32_t addr, float2* f) {  // 8 consecutive fp32 as 4 pairs
  const uint4 a = lds128(addr), b = lds128(addr + 16);
  f[0] = make_float2(__uint_as_float(a.x), __uint_as_float(a.y));
  f[1] = make_float2(__uint_as_float(a.z), __uint_as_float(a.w));
  f[2] = make_float2(__uint_as_float(b.x), __uint_as_float(b.y));
  f[3] = make_float2(__uint_as_float(b.z), __uint_as_float(b.w));
}

// Fused GRN + GELU backward of one row's 32-column chunk: dh = (acc * s + g * t) * gp on the packed fp32x2 pipe (three
// instructions per column pair); the s / t vectors of the next 4 columns are requested before the current 4 are used, so
// their shared-memory latency runs under the math instead of in front of every group; results go straight into the
// warp's staging tile (stg_row = this lane's 64-byte row, sw = its swizzle phase).
// cv_s / cv_t: shared-space addresses of s and t at this chunk's first column.
// Steps [H0, H1) of 4 columns each (8 steps = the chunk); xa / xb = g / gp of the 8-column groups H0/2 ... as uint4.
template <bool BF16, int H0, int H1>
__device__ __forceinline__ void dgelu_grn_steps(const uint32_t* acc, const uint4* xa, const uint4* xb, uint32_t cv_s,
                                                uint32_t cv_t, uint32_t stg_row, int sw) {
  uint4 sv[2], tv[2];  // s / t of 4 columns, fetched one step ahead (16 registers in flight)
  sv[H0 & 1] = lds128(cv_s + H0 * 16);
  tv[H0 & 1] = lds128(cv_t + H0 * 16);
#pragma unroll
  for (int h = H0; h < H1; ++h) {
    if (h + 1 < H1) {
      sv[(h + 1) & 1] = lds128(cv_s + (h + 1) * 16);
      tv[(h + 1) & 1] = lds128(cv_t + (h + 1) * 16);
    }
    const int g = h >> 1, gl = g - (H0 >> 1), w = (h & 1) * 2;
    const uint32_t ga[4] = {xa[gl].x, xa[gl].y, xa[gl].z, xa[gl].w};
    const uint32_t gb[4] = {xb[gl].x, xb[gl].y, xb[gl].z, xb[gl].w};
    const uint4 s4 = sv[h & 1], t4 = tv[h & 1];
    const float2 s2[2] = {make_float2(__uint_as_float(s4.x), __uint_as_float(s4.y)),
                          make_float2(__uint_as_float(s4.z), __uint_as_float(s4.w))};
    const float2 t2[2] = {make_float2(__uint_as_float(t4.x), __uint_as_float(t4.y)),
                          make_float2(__uint_as_float(t4.z), __uint_as_float(t4.w))};
    uint32_t q[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2 a2 = make_float2(__uint_as_float(acc[h * 4 + 2 * j]), __uint_as_float(acc[h * 4 + 2 * j + 1]));
      const float2 v2 = __fmul2_rn(__ffma2_rn(H16<BF16>::unpack(ga[w + j]), t2[j], __fmul2_rn(a2, s2[j])),
                                   H16<BF16>::unpack(gb[w + j]));
      q[j] = H16<BF16>::pack(v2.x, v2.y);
    }
    // this lane's row of the staging tile (swizzled 16-byte groups), 8 bytes per step
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(stg_row + ((g ^ sw) << 4) + (h & 1) * 8), "r"(q[0]), "r"(q[1]) : "memory");
  }
}

// GRN statistic: column sums of g^2 over a warp's 32 rows, read back (as stored, 16-bit rounded) from the staged tile -
// lane = column; row r, 16-byte group q sits at r * 64 + ((q ^ (r >> 1 & 3)) << 4).  One atomicAdd per column.
template <bool BF16>
__device__ __forceinline__ void colsq_from_tile(uint32_t tile, int lane, int rows_valid, float* colsq, int N, long long sample,
                                                int col0, int n_lim) {
  const uint32_t gt = tile + static_cast<uint32_t>((lane & 7) * 2);
  const int q8 = lane >> 3;
  float sq = 0.f;
#pragma unroll 8
  for (int rr = 0; rr < 32; ++rr) {
    uint16_t hv;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(gt + rr * 64 + ((q8 ^ ((rr >> 1) & 3)) << 4)));
    const float f = H16<BF16>::to_f(*reinterpret_cast<const typename H16<BF16>::T*>(&hv));
    if (rr < rows_valid) sq = fmaf(f, f, sq);
  }
  const int col = col0 + lane;
  if (col < n_lim) atomicAdd(colsq + sample * N + col, sq);
}

// BF16: element type of the 16-bit epilogue operands / outputs (compile-time so that the unrolled epilogue of a chunk is
// one basic block; the MMA element type comes from the instruction descriptor, p.bf16)
template <int BN, int MODE, int EPI, int BKE = BK, bool BF16 = true>
__global__ void __launch_bounds__(64 + 32 * EpiWarps<BN, EPI>::value, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ OutMaps tmOut, const GemmParams p) {
  constexpr int NUM_EPI_WARPS = EpiWarps<BN, EPI>::value;
  constexpr bool WRES = MODE == MODE_CONVKPW;
  constexpr bool PATCH = MODE == MODE_CONVKP || MODE == MODE_CONVKP2 || WRES;
  constexpr bool PAIR = MODE == MODE_KMAJOR2 || MODE == MODE_CONVKP2;
  static_assert(MODE != MODE_KMAJOR2 || BN == 256, "CTA-pair GEMM: 256-wide tiles");
  static_assert(MODE != MODE_CONVKP2 || BKE == 64, "CTA-pair patch conv: 64-channel K blocks");
  constexpr bool AUXT = EpiWarps<BN, EPI>::aux_tma(PAIR);
  using C = Cfg<BN, BKE, NUM_EPI_WARPS, EpiWarps<BN, EPI>::tma_store, PATCH, WRES, PAIR, AUXT>;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int COL_GROUPS = NUM_EPI_WARPS / 4;  // warps sharing a TMEM lane quarter split the tile's columns
  constexpr bool MN_MAJOR = MODE == MODE_MNMAJOR || MODE == MODE_CONVMN;
  static_assert(BKE == 64 || (BKE == 32 && (MODE == MODE_CONVK || PATCH)), "BKE = 32 is the 32-channel conv form only");
  static_assert(!WRES || BN == 64, "resident-filter form: one 64-wide output tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* w_bar = tmem_empty + 3;             // CONVKPW: resident filter landed
  uint64_t* aux_bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + 128);  // AUXT: one per epilogue warp
  uint8_t* wres = smem + C::W_OFFSET;           // CONVKPW: [tap][chunk][BN x BKE] sub-tiles

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], PAIR ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);  // pair: both CTAs' epilogues release the leader
    }
    if constexpr (WRES) mbar_init(w_bar, 1);
    if constexpr (AUXT) {
      for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(&aux_bars[i], 1);
      tma_prefetch_desc(&tmOut.o);
      tma_prefetch_desc(&tmOut.o2);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc2(tmem_ptr, C::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_ptr, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  // MN-major forms with M <= 64 (weight gradients of <= 64-channel layers): rows 64..127 of the A tile are all padding.
  // They are zeroed once here and their TMA box is never issued, which saves a third of the boxes per K block.
  const bool a_half = MN_MAJOR && p.M <= 64;
  if (a_half) {
    for (int s = 0; s < C::STAGES; ++s) {
      uint4* z = reinterpret_cast<uint4*>(smem + s * C::STAGE_BYTES + 64 * BK * 2);
      for (int i = threadIdx.x; i < 64 * BK * 2 / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles = p.tiles_m * p.tiles_n;
  const int units = tiles * p.k_splits;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (WRES) {  // the whole filter, once: one box per (tap, channel chunk)
        mbar_expect_tx(w_bar, p.cwbytes);
        const int sub = p.cwrows * BKE * 2;
        for (int i = 0; i * sub < p.cwbytes; ++i) tma_load_2d(wres + i * sub, &tmB, w_bar, i * BKE, 0);
      }
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int split = unit / tiles;
        const int t = unit - split * tiles;
        // pair form: CTAs 2c, 2c+1 (one cluster) walk units 2q, 2q+1 in lockstep = the two 128-row halves of pair tile q
        const int m0 = PAIR ? (((t >> 1) / p.tiles_n) * 2 + (int)rank) * BM : (t / p.tiles_n) * BM;
        const int n0 = PAIR ? ((t >> 1) % p.tiles_n) * BN : (t % p.tiles_n) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        // conv forms: CONVK walks (tap, channel chunk) along K for a fixed 128-voxel box; CONVMN walks 64-voxel boxes
        // along K for a fixed (tap, channel tile)
        // (the producer is ONE thread: integer divisions per K block would bound the whole kernel, so tap and voxel
        //  coordinates are decomposed once per unit and then advanced with carries)
        int cx = 0, cy = 0, cz = 0, cn = 0, chunk = 0, ci0 = 0, kw = 0, kh = 0, kd = 0, tapi = 0;
        if constexpr (PATCH) {  // M tile -> (n, z, patch row, patch column); K blocks walk (kd, kw, chunk)
          int mt = PAIR ? ((t >> 1) / p.tiles_n) * 2 + (int)rank : t / p.tiles_n;  // pair: two x-adjacent patches
          cx = (mt % p.cpxn) * 16 - p.cpw;
          mt /= p.cpxn;
          cy = (mt % p.cpyn) * 8 - p.cph;
          mt /= p.cpyn;
          cz = mt % p.cOD - p.cpd;
          cn = mt / p.cOD;
        }
        if constexpr (MODE == MODE_CONVK) {
          voxel_coords(p, m0, cx, cy, cz, cn);
          const int tap = kb0 / p.cchunks;
          tapi = tap;
          chunk = kb0 - tap * p.cchunks;
          kw = tap % p.cKW;
          kh = (tap / p.cKW) % p.cKH;
          kd = tap / (p.cKW * p.cKH);
          cx = cx * p.csw - p.cpw;
          cy = cy * p.csh - p.cph;
          cz = cz * p.csd - p.cpd;
        }
        if constexpr (MODE == MODE_CONVMN) {
          const int tn = t % p.tiles_n;
          const int tap = tn / p.cchunks;
          ci0 = (tn - tap * p.cchunks) * BN;
          kw = tap % p.cKW;
          kh = (tap / p.cKW) % p.cKH;
          kd = tap / (p.cKW * p.cKH);
          voxel_coords(p, kb0 * BK, cx, cy, cz, cn);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          if constexpr (AUXT) mbar_wait_backoff(&empty_bar[stage], phase ^ 1, 64);
          else mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if constexpr (PAIR && PATCH) {
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            tma_load_5d_pair(sa, &tmA, lead_bar, chunk * BKE, cx + kw, cy, cz + kd, cn);
#pragma unroll
            for (int h = 0; h < 3; ++h)  // this CTA's half of the output channels of each kh sub-tile
              tma_load_2d_pair(sb + h * C::B_TAP_BYTES, &tmB, lead_bar, ((kd * 3 + h) * p.cKW + kw) * p.ccin + chunk * BKE,
                               n0 + (int)rank * (BN / 2));
            if (++chunk == p.cchunks) {
              chunk = 0;
              if (++kw == p.cKW) {
                kw = 0;
                ++kd;
              }
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          } else if constexpr (PAIR) {
            // both CTAs' boxes are counted on the LEADER's full barrier (it issues the MMA that reads both halves)
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_pair(sa, &tmA, lead_bar, kb * BK, m0);
            const int brow = (p.b_batch_rows > 0 ? (m0 / p.b_batch_rows) * p.N + n0 : n0) + (int)rank * (BN / 2);
            tma_load_2d_pair(sb, &tmB, lead_bar, kb * BK, brow);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], a_half ? C::STAGE_BYTES - 64 * BK * 2 : C::STAGE_BYTES);
          if constexpr (PATCH) {
            tma_load_5d(sa, &tmA, &full_bar[stage], chunk * BKE, cx + kw, cy, cz + kd, cn);
            if constexpr (!WRES) {
#pragma unroll
              for (int h = 0; h < 3; ++h)
                tma_load_2d(sb + h * C::B_TAP_BYTES, &tmB, &full_bar[stage],
                            ((kd * 3 + h) * p.cKW + kw) * p.ccin + chunk * BKE, n0);
            }
            if (++chunk == p.cchunks) {
              chunk = 0;
              if (++kw == p.cKW) {
                kw = 0;
                ++kd;
              }
            }
          } else if constexpr (MODE == MODE_CONVK) {
            tma_load_5d(sa, &tmA, &full_bar[stage], chunk * BKE, cx + kw, cy + kh, cz + kd, cn);
            tma_load_2d(sb, &tmB, &full_bar[stage], (p.ntap ? p.ctap[tapi] : tapi) * p.ccin + chunk * BKE, n0);
            if (++chunk == p.cchunks) {  // next tap
              chunk = 0;
              ++tapi;
              if (++kw == p.cKW) {
                kw = 0;
                if (++kh == p.cKH) {
                  kh = 0;
                  ++kd;
                }
              }
            }
          } else if constexpr (MODE == MODE_CONVMN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], m0, kb * BK);
            if (!a_half) tma_load_2d(sa + 64 * BK * 2, &tmA, &full_bar[stage], m0 + 64, kb * BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_5d(sb + j * (64 * BK * 2), &tmB, &full_bar[stage], ci0 + j * 64, cx * p.csw + kw - p.cpw,
                          cy * p.csh + kh - p.cph, cz * p.csd + kd - p.cpd, cn);
            cx += p.cbx;  // next 64-voxel box of the flattened (n, z, y, x) order
            if (cx >= p.cOW) {
              cx = 0;
              cy += p.cby;
              if (cy >= p.cOH) {
                cy = 0;
                cz += p.cbz;
                if (cz >= p.cOD) {
                  cz = 0;
                  cn += p.cbn;
                }
              }
            }
          } else if constexpr (!MN_MAJOR) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
            const int brow = p.b_batch_rows > 0 ? (m0 / p.b_batch_rows) * p.N + n0 : n0;
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, brow);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], m0, kb * BK);
            if (!a_half) tma_load_2d(sa + 64 * BK * 2, &tmA, &full_bar[stage], m0 + 64, kb * BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (64 * BK * 2), &tmB, &full_bar[stage], n0 + j * 64, kb * BK);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loops (warp-uniform control flow keeps the address arithmetic on the uniform datapath);
    // one elected lane issues tcgen05.mma / tcgen05.commit.  Descriptors are built once per stage and advanced by
    // adding byte offsets >> 4 to their address field: with 16-32-cycle MMAs (N = 32 / 64 tiles) the issue loop of this
    // single thread, not the tensor pipe, was the measured limit.
    const uint32_t idesc = make_idesc(PAIR ? 2 * BM : BM, WRES ? p.cwrows : BN, p.bf16 != 0, MN_MAJOR, MN_MAJOR);
    const uint64_t desc_a0 = BKE == 32 ? make_smem_desc_sw64(0, 0, 512)
                             : (MN_MAJOR ? make_smem_desc(0, 64 * BK * 2, 1024) : make_smem_desc(0, 0, 1024));
    const uint64_t desc_b0 = desc_a0;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t wres0 = smem_u32(wres);
    const bool leader = elect_one();  // elected once: tcgen05.commit tracks the MMAs of the thread that issues it
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (WRES) {
      mbar_wait(w_bar, 0);
      tc_fence_after();
    }
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      if (PAIR && rank != 0) break;  // the leader CTA issues for the pair
      const int split = unit / tiles;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      if constexpr (AUXT) mbar_wait_backoff(&tmem_empty[acc], acc_phase ^ 1, 32);
      else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      // CONVKPW: resident sub-tile index of tap (kd, h = 0, kw), chunk: +1 per K block, +2 filter rows when kw wraps
      int widx = 0, winplane = 0;
      const int wplane = p.cKW * p.cchunks;
      const uint32_t wsub = static_cast<uint32_t>(p.cwrows) * BKE * 2;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem0 + stage * C::STAGE_BYTES;
        const uint32_t sb = sa + C::A_BYTES;
        const uint64_t da_s = desc_a0 + (sa >> 4);
        const uint64_t db_s = desc_b0 + (sb >> 4);
        if (leader) {
          if constexpr (PATCH) {
            // three kh taps: the haloed A box viewed 16 voxel rows (one patch line) further down, its own B sub-tile
#pragma unroll
            for (int h = 0; h < 3; ++h) {
              // weights: this stage's sub-tile of tap (kd, h, kw), or its resident copy [tap][chunk]
              const uint64_t db_h = WRES ? desc_b0 + ((wres0 + static_cast<uint32_t>(widx + h * wplane) * wsub) >> 4)
                                         : db_s + ((h * C::B_TAP_BYTES) >> 4);
#pragma unroll
              for (int k = 0; k < BKE / 16; ++k) {
                if constexpr (PAIR)
                  tc_mma_f16_2(d_tmem, da_s + ((h * 16 * (BKE * 2) + k * 32) >> 4), db_h + ((k * 32) >> 4), idesc,
                               (kb > kb0 || h > 0 || k > 0) ? 1u : 0u);
                else
                  tc_mma_f16(d_tmem, da_s + ((h * 16 * (BKE * 2) + k * 32) >> 4), db_h + ((k * 32) >> 4), idesc,
                             (kb > kb0 || h > 0 || k > 0) ? 1u : 0u);
              }
            }
          } else {
            // K-major: step 16 elements (32 B) inside the swizzled row; MN-major: step 16 k rows = 2048 B
            constexpr int KSTEP = MN_MAJOR ? 2048 : 32;
#pragma unroll
            for (int k = 0; k < BKE / 16; ++k) {
              if constexpr (PAIR)
                tc_mma_f16_2(d_tmem, da_s + ((k * KSTEP) >> 4), db_s + ((k * KSTEP) >> 4), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else
                tc_mma_f16(d_tmem, da_s + ((k * KSTEP) >> 4), db_s + ((k * KSTEP) >> 4), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          if constexpr (PAIR) tc_commit2(&empty_bar[stage]);  // frees the slot in both CTAs
          else tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        }
        __syncwarp();
        if constexpr (WRES) {
          ++widx;
          if (++winplane == wplane) {
            winplane = 0;
            widx += 2 * wplane;
          }
        }
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) {  // accumulator complete -> epilogue (of both CTAs in the pair form)
        if constexpr (PAIR) tc_commit2(&tmem_full[acc]);
        else tc_commit(&tmem_full[acc]);
      }
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue warps =====================
    const int e = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    const int half = e >> 2;       // which group of BN / COL_GROUPS columns
    float* colv = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);
    const uint32_t stg = smem_u32(smem + C::STG_OFFSET + e * C::STG_PER_WARP);
    constexpr bool TMA_STORE = EpiWarps<BN, EPI>::tma_store;
    constexpr int NCH = BN / (32 * COL_GROUPS);  // 32-column chunks per warp
    uint32_t tile_it = 0;  // accumulator buffer = bit 0, its barrier phase = bit 1
    // AUXT: this warp's g / gp chunk (32 rows x 64 B each) is fetched by TMA into its second and third tile (SWIZZLE_64B,
    // the layout `unstage` reads) and the chunk after it is requested as soon as this one sits in registers, so the
    // operand latency runs under the math and stores of the previous chunk instead of in front of every chunk.
    // The request cursor (unit, pair-tile row / column, chunk) runs one chunk ahead of the consumer and advances without
    // divisions: a step of gridDim.x units is a fixed (rows, columns) step through the pair-tile grid.
    const uint32_t aux_bar = smem_u32(aux_bars + e);
    uint32_t aux_phase = 0;
    int pf_u = blockIdx.x, pf_c = -1, pf_mi = 0, pf_ni = 0, pf_dq = 0, pf_dr = 0;
    if constexpr (AUXT) {
      const int q0 = blockIdx.x >> 1, qs = gridDim.x >> 1;
      pf_mi = q0 / p.tiles_n;
      pf_ni = q0 - pf_mi * p.tiles_n;
      pf_dq = qs / p.tiles_n;
      pf_dr = qs - pf_dq * p.tiles_n;
    }
    auto aux_prefetch = [&]() {  // warp-uniform: advance to the next chunk this warp owns and request it
      if constexpr (AUXT) {
        ++pf_c;
        while (pf_u < units) {
          const int ac0 = pf_ni * BN + half * (BN / COL_GROUPS) + pf_c * 32;
          if (pf_c < NCH && ac0 < p.N) {
            if (lane == 0) {
              const int ar0 = (pf_mi * 2 + (int)rank) * BM + quarter * 32;
              const bool with_g = p.tvec != nullptr;
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(aux_bar), "r"(with_g ? 4096u : 2048u)
                           : "memory");
              if (with_g)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                                 "r"(stg + 2048u), "l"(reinterpret_cast<uint64_t>(&tmOut.o)), "r"(aux_bar), "r"(ac0), "r"(ar0)
                             : "memory");
              asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                               "r"(stg + 4096u), "l"(reinterpret_cast<uint64_t>(&tmOut.o2)), "r"(aux_bar), "r"(ac0), "r"(ar0)
                           : "memory");
            }
            return;
          }
          pf_u += gridDim.x;
          pf_c = 0;
          pf_mi += pf_dq;
          pf_ni += pf_dr;
          if (pf_ni >= p.tiles_n) {
            pf_ni -= p.tiles_n;
            ++pf_mi;
          }
        }
      }
    };
    aux_prefetch();
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const int acc = tile_it & 1;
      const uint32_t acc_phase = (tile_it >> 1) & 1;
      const int split = unit / tiles;
      const int t = unit - split * tiles;
      const int m0 = PAIR ? (((t >> 1) / p.tiles_n) * 2 + (int)rank) * BM : (t / p.tiles_n) * BM;
      int n0 = PAIR ? ((t >> 1) % p.tiles_n) * BN : (t % p.tiles_n) * BN;
      int n_lim = p.N;
      if constexpr (MODE == MODE_CONVMN) {  // N tile = (tap, channel tile): columns [tap*cin + ci0, (tap+1)*cin)
        const int tn = t % p.tiles_n;
        const int tap = tn / p.cchunks;
        n0 = tap * p.ccin + (tn - tap * p.cchunks) * BN;
        n_lim = (tap + 1) * p.ccin;
      }
      // this warp's per-column vectors (bias and, for the fused GRN backward, s and t of the tile's sample) for its own
      // BN / COL_GROUPS columns, in its private slice of shared memory: a __syncwarp instead of a barrier across all epilogue
      // warps per tile (the barrier cost 12 % of the epilogue warps' time in the source-level profile)
      constexpr int CW = BN / COL_GROUPS;
      float* cvp = colv + e * (PATCH ? 1 : 3) * CW;
      const uint32_t cv = smem_u32(cvp) - static_cast<uint32_t>(half * CW) * 4u;  // indexed by the column within the tile
      __syncwarp();  // the previous tile's reads of this slice are done
#pragma unroll
      for (int i = lane; i < CW; i += 32) {
        const int col = n0 + half * CW + i;
        const bool ok = col < p.N;
        cvp[i] = (p.bias != nullptr && split == 0 && ok) ? __ldg(p.bias + col) : 0.0f;
        if constexpr ((EPI == VB200_EPI_DGELU_GRN || EPI == VB200_EPI_STORE) && !PATCH) {
          const long long ns = p.rows_per_sample > 0 ? m0 / p.rows_per_sample : 0;
          cvp[CW + i] = (p.svec != nullptr && ok) ? __ldg(p.svec + ns * p.N + col) : 1.0f;
          cvp[2 * CW + i] = (p.tvec != nullptr && ok) ? __ldg(p.tvec + ns * p.N + col) : 0.0f;
        }
      }
      __syncwarp();
      const long long row0 = m0 + quarter * 32;
      long long rowoff[4];
      bool scatter = false;
      if constexpr (PATCH) {  // rows of the tile = voxels of a 16 x 8 patch: each goes to its own output row
        scatter = true;
        int mt = PAIR ? ((t >> 1) / p.tiles_n) * 2 + (int)rank : t / p.tiles_n;
        const int px = mt % p.cpxn;
        mt /= p.cpxn;
        const int py = mt % p.cpyn;
        mt /= p.cpyn;  // = n * OD + z
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = quarter * 32 + (lane >> 2) + 8 * i;
          const long long row = ((long long)mt * p.cOH + py * 8 + (r >> 4)) * p.cOW + px * 16 + (r & 15);
          rowoff[i] = row * p.ldo * 2;
        }
      }
      if constexpr (MODE == MODE_CONVK && EPI == VB200_EPI_STORE) {
        scatter = p.opx != 0;
        if (scatter) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int vx, vy, vz, vn;
            voxel_coords(p, (int)min(row0 + (lane >> 2) + 8 * i, (long long)p.M - 1), vx, vy, vz, vn);
            rowoff[i] = (vn * p.opn + vz * p.opz + vy * p.opy + vx * p.opx) * 2;
          }
        }
      }
      long long lda_unused;
      const bool has_a = aux_a_ptr<EPI>(p, lda_unused) != nullptr;  // warp-uniform
      constexpr bool APRE = NUM_EPI_WARPS == 8;  // aux operands fetched one chunk ahead (second register buffer)
      AuxRegs aux[APRE ? 2 : 1];
      const int cc0 = half * (BN / COL_GROUPS);
      if constexpr (!AUXT) load_aux<EPI>(p, row0, n0 + cc0, lane, aux[0]);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
      // accumulator chunks are fetched one ahead of the math (two register buffers); the TMEM buffer is handed back
      // to the MMA warp as soon as this warp's last chunk sits in registers
      // (not for EPI_DGELU_GRN, whose two double-buffered aux operands already fill the register file)
      constexpr bool TPRE = EPI != VB200_EPI_DGELU_GRN && NUM_EPI_WARPS == 8;
      uint32_t r[TPRE ? 2 : 1][32];
      const int rows_valid = (int)min(32LL, (long long)p.M - row0);
      float rscale = 1.0f;  // this lane's row: TMEM lane = tile row
      if constexpr (EPI == VB200_EPI_STORE) {
        if (p.rvec != nullptr) rscale = __ldg(p.rvec + min(row0 + lane, (long long)p.M - 1) / p.rvec_rows);
      }
      if (TPRE && n0 + cc0 < n_lim) tmem_ld32(t_addr + cc0, r[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int cc = cc0 + c * 32;
        const int col0 = n0 + cc;
        if (col0 < n_lim) {  // warp-uniform
          const bool more = c + 1 < NCH && col0 + 32 < n_lim;
          if constexpr (!TPRE) tmem_ld32(t_addr + cc, r[0]);
          const int AB = APRE ? (c & 1) : 0;
          if constexpr (AUXT) {
            asm volatile(
                "{\n\t.reg .pred P1;\n\t"
                "AUX_WAIT:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                "@P1 bra AUX_DONE;\n\t"
                "bra AUX_WAIT;\n\t"
                "AUX_DONE:\n\t}\n" ::"r"(aux_bar),
                "r"(aux_phase)
                : "memory");
            aux_phase ^= 1;
          } else {
            if constexpr (APRE) {
              if (more) load_aux<EPI>(p, row0, col0 + 32, lane, aux[(c + 1) & 1]);
            } else {
              if (c > 0) load_aux<EPI>(p, row0, col0, lane, aux[0]);
            }
            if (has_a) unstage(stg, lane, aux[AB].a);
            if constexpr (EPI == VB200_EPI_DGELU_GRN) unstage(stg, lane, aux[AB].b);
          }
          tmem_ld_wait();
          if (TPRE && more) {
            tmem_ld32(t_addr + cc + 32, r[(c + 1) & 1]);
          } else if (!more) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              // relaxed: the barrier guards TMEM only (tcgen05.fence above); a release would wait for every store in flight
              if constexpr (PAIR) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
              else mbar_arrive_relaxed(&tmem_empty[acc]);
            }
          }
          // this warp's 32 rows x 32 columns: math per row, then row-contiguous stores through the staging tile
          const int cols8_valid = min(4, (n_lim - col0) >> 3);
          if (rows_valid > 0) {
            uint4 o1[4], o2[4];
            float f32buf[32];
            if constexpr (EPI == VB200_EPI_DGELU_GRN) {
              const uint32_t cv_s = cv + static_cast<uint32_t>(CW + cc) * 4u, cv_t = cv + static_cast<uint32_t>(2 * CW + cc) * 4u;
              const int sw = (lane >> 1) & 3;
              if constexpr (AUXT) {
                // g / gp leave their tiles 16 columns at a time (16 registers instead of 32); once the second half is in
                // registers the tiles are free and the next chunk is requested, under the second half's math and the stores
                uint4 xa[2], xb[2];
                __syncwarp();  // the previous chunk's row-contiguous reads of the staging tile are done
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                  xa[g] = has_a ? lds128(stg + 2048 + lane * 64 + ((g ^ sw) << 4)) : make_uint4(0, 0, 0, 0);
                  xb[g] = lds128(stg + 4096 + lane * 64 + ((g ^ sw) << 4));
                }
                dgelu_grn_steps<BF16, 0, 4>(r[0], xa, xb, cv_s, cv_t, stg + lane * 64, sw);
#pragma unroll
                for (int g = 2; g < 4; ++g) {
                  xa[g - 2] = has_a ? lds128(stg + 2048 + lane * 64 + ((g ^ sw) << 4)) : make_uint4(0, 0, 0, 0);
                  xb[g - 2] = lds128(stg + 4096 + lane * 64 + ((g ^ sw) << 4));
                }
                __syncwarp();
                if (lane == 0) fence_proxy_async_smem();  // the warp's generic-proxy reads precede the async-proxy refill
                aux_prefetch();
                dgelu_grn_steps<BF16, 4, 8>(r[0], xa, xb, cv_s, cv_t, stg + lane * 64, sw);
              } else {
                dgelu_grn_steps<BF16, 0, 8>(r[0], aux[AB].a, aux[AB].b, cv_s, cv_t, stg + lane * 64, sw);
              }
            }
#pragma unroll
            for (int g = 0; g < (EPI == VB200_EPI_DGELU_GRN ? 0 : 4); ++g) {
              float v[8], w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[TPRE ? (c & 1) : 0][g * 8 + j]);
              epilogue_math8<EPI, BN, BF16, CW>(p, v, w, cc + g * 8, cv, aux[AB].a[g], aux[AB].b[g], rscale);
              if constexpr (EPI == VB200_EPI_F32) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f32buf[g * 8 + j] = v[j];
              } else {
                o1[g] = pack8<BF16>(v);
                if constexpr (EPI == VB200_EPI_GELU_DUAL || EPI == VB200_EPI_GELU_GP) o2[g] = pack8<BF16>(w);
              }
            }
            if constexpr (EPI == VB200_EPI_F32) {
              float* ob = reinterpret_cast<float*>(p.out) + (long long)split * p.split_out_stride;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {  // 16 fp32 columns (64 B per row) at a time
                uint4 q[4];
#pragma unroll
                for (int g = 0; g < 4; ++g)
                  q[g] = make_uint4(__float_as_uint(f32buf[hh * 16 + g * 4]), __float_as_uint(f32buf[hh * 16 + g * 4 + 1]),
                                    __float_as_uint(f32buf[hh * 16 + g * 4 + 2]), __float_as_uint(f32buf[hh * 16 + g * 4 + 3]));
                const int colh = col0 + hh * 16;
                float* dst = ob;
                long long ldb = p.ldo * 4, cb = (long long)colh * 4;
                int lim = n_lim;
                if (p.n_split > 0) {  // second destination for the trailing columns (bias-gradient column of [l | 1])
                  if (colh >= p.n_split) {
                    dst = reinterpret_cast<float*>(p.out2);
                    ldb = p.ldo2 * 4;
                    cb = (long long)(colh - p.n_split) * 4;
                  } else {
                    lim = min(n_lim, p.n_split);
                  }
                }
                const int c4v = min(4, max(0, (lim - colh) >> 2));
                if (p.atomic_out)
                  stage_store<true>(stg, lane, q, dst, ldb, row0, rows_valid, cb, c4v);
                else
                  stage_store<false>(stg, lane, q, dst, ldb, row0, rows_valid, cb, c4v);
              }
            } else if constexpr (TMA_STORE) {
              if constexpr (EPI == VB200_EPI_GELU_DUAL || EPI == VB200_EPI_GELU_GP) {
                // two tiles, two stores in flight: each waits only for the store that last read its own tile
                tma_stage_store<1>(stg, lane, o1, &tmOut.o, col0, row0);
                tma_stage_store<1>(stg + 2048, lane, o2, &tmOut.o2, col0, row0);
                if constexpr (EPI == VB200_EPI_GELU_GP) {
                  if (p.colsq != nullptr) colsq_from_tile<BF16>(stg + 2048, lane, rows_valid, p.colsq, p.N, row0 / p.rows_per_sample, col0, n_lim);
                }
              }
            } else {
              if constexpr (EPI == VB200_EPI_DGELU_GRN)
                stage_store<false>(stg, lane, nullptr, p.out, p.ldo * 2, row0, rows_valid, (long long)col0 * 2, cols8_valid);
              else
                stage_store<false>(stg, lane, o1, p.out, p.ldo * 2, row0, rows_valid, (long long)col0 * 2, cols8_valid,
                                   scatter ? rowoff : nullptr);
              if constexpr (EPI == VB200_EPI_GELU_DUAL || EPI == VB200_EPI_GELU_GP)
                stage_store<false>(stg, lane, o2, p.out2, p.ldo2 * 2, row0, rows_valid, (long long)col0 * 2, cols8_valid);
              if constexpr (EPI == VB200_EPI_GELU_GP) {  // the staging tile still holds g
                if (p.colsq != nullptr) colsq_from_tile<BF16>(stg, lane, rows_valid, p.colsq, p.N, row0 / p.rows_per_sample, col0, n_lim);
              }
            }
          } else if constexpr (AUXT) {  // rows past M: the (zero-filled) chunk still has to be retired
            __syncwarp();
            aux_prefetch();
          }
        }
      }
      if (n0 + cc0 >= n_lim) {  // this warp's column range lies past the tile's valid columns: nothing drained above
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
          else mbar_arrive_relaxed(&tmem_empty[acc]);
        }
      }
      ++tile_it;
    }
    if constexpr (TMA_STORE) {  // the staging tiles must outlive the bulk stores that read them
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) ==
            cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// 2-D row-major 16-bit tensor [rows, cols] with leading dimension ld (elements); box = {box_c, box_r}
int make_tmap_2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld,
                 int box_c, int box_r, bool bf16, bool sw64 = false) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(VB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 8) != 0)
    return fail(VB200_ERR_UNSUPPORTED, "TMA operand needs 16-byte aligned base and ld %% 8 == 0 (ld=%lld)", ld);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_r)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return VB200_OK;
}

// channels-last activation [N, D, H, W, C] as a 5-D map (C, X, Y, Z, N); box = box_c channels x (bx, by, bz, bn) voxels.
// In shared memory the box is rows of box_c channels (128 B, or 64 B with SWIZZLE_64B) in (n, z, y, x) order: exactly
// the K-major / MN-major swizzled operand tile of the 2-D forms.
static int make_tmap_conv(CUtensorMap* m, const void* base, int N, int D, int H, int W, int Cc, int box_c,
                          const int* vbox, const int* stride /* w, h, d */, bool bf16, bool sw64) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(VB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || Cc % 8 != 0)
    return fail(VB200_ERR_UNSUPPORTED, "conv operand needs a 16-byte aligned base and C %% 8 == 0 (C=%d)", Cc);
  cuuint64_t dims[5] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2,
                           (cuuint64_t)D * H * W * Cc * 2};
  // strided convs: the box spans s*b voxels and the element stride s makes TMA deliver every s-th one (b voxels)
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)(vbox[0] * stride[0]), (cuuint32_t)(vbox[1] * stride[1]),
                       (cuuint32_t)(vbox[2] * stride[2]), (cuuint32_t)vbox[3]};
  cuuint32_t estr[5] = {1, (cuuint32_t)stride[0], (cuuint32_t)stride[1], (cuuint32_t)stride[2], 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VB200_ERR_CUDA, "cuTensorMapEncodeTiled (5-D conv operand) failed (%d)", (int)r);
  return VB200_OK;
}

// `rows` consecutive voxels of the flattened (n, z, y, x) output order form one box iff each extent divides / is divided
// by what is left of `rows`; a tile larger than the whole tensor spills into zero-filled samples.
static bool conv_box(int rows, int OW, int OH, int OD, int NB, int* box) {
  const int dims[4] = {OW, OH, OD, NB};
  int rem = rows;
  for (int i = 0; i < 4; ++i) {
    if (rem >= dims[i]) {
      if (rem % dims[i]) return false;
      box[i] = dims[i];
      rem /= dims[i];
    } else {
      if (dims[i] % rem) return false;
      box[i] = rem;
      rem = 1;
    }
  }
  if (rem > 1) box[3] *= rem;
  return box[0] <= 256 && box[1] <= 256 && box[2] <= 256 && box[3] <= 256;
}

int sm_count() {  // of the current device (one process may drive several)
  static std::atomic<int> cache[64];
  const int dev = PerDeviceOnce::device();
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

template <int BN, int MODE, int EPI, int BKE, bool BF16>
static int launch_dt(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid,
                     cudaStream_t st) {
  static PerDeviceOnce once;  // per instantiation and device
  const int dev = PerDeviceOnce::device();
  auto kern = gemm_kernel<BN, MODE, EPI, BKE, BF16>;
  constexpr bool LPAIR = MODE == MODE_KMAJOR2 || MODE == MODE_CONVKP2;
  using LC = Cfg<BN, BKE, EpiWarps<BN, EPI>::value, EpiWarps<BN, EPI>::tma_store,
                 MODE == MODE_CONVKP || MODE == MODE_CONVKP2 || MODE == MODE_CONVKPW, MODE == MODE_CONVKPW, LPAIR,
                 EpiWarps<BN, EPI>::aux_tma(LPAIR)>;
  const int smem_bytes = MODE == MODE_CONVKPW ? LC::W_OFFSET + 1024 + p.cwbytes : LC::SMEM_BYTES;
  if (once.need(dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MODE == MODE_CONVKPW ? 232448 : LC::SMEM_BYTES);
    if (e != cudaSuccess) return fail(VB200_ERR_CUDA, "smem attribute: %s", cudaGetErrorString(e));
    once.done(dev);
  }
  OutMaps om{};
  if constexpr (EpiWarps<BN, EPI>::tma_store) {
    if (int rc = make_tmap_2d(&om.o, p.out, p.M, p.N, p.ldo, 32, 32, BF16, true)) return rc;
    if (p.out2 != nullptr)
      if (int rc = make_tmap_2d(&om.o2, p.out2, p.M, p.N, p.ldo2, 32, 32, BF16, true)) return rc;
  }
  if constexpr (EpiWarps<BN, EPI>::aux_tma(LPAIR)) {  // the maps carry the epilogue's input operands instead
    if (p.tvec != nullptr)
      if (int rc = make_tmap_2d(&om.o, p.aux, p.M, p.N, p.ldaux, 32, 32, BF16, true)) return rc;
    if (int rc = make_tmap_2d(&om.o2, p.aux2, p.M, p.N, p.ldaux2, 32, 32, BF16, true)) return rc;
  }
  if (smem_bytes > 232448) return fail(VB200_ERR_UNSUPPORTED, "resident filter does not fit shared memory (%d B)", smem_bytes);
  if constexpr (MODE == MODE_KMAJOR2 || MODE == MODE_CONVKP2) {  // clusters of two CTAs: the pair shares one cta_group::2 MMA
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(64 + 32 * EpiWarps<BN, EPI>::value);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, om, p);
    if (e != cudaSuccess) return fail(VB200_ERR_CUDA, "vb200_gemm (CTA pair): %s", cudaGetErrorString(e));
    return check_launch("vb200_gemm");
  }
  kern<<<grid, 64 + 32 * EpiWarps<BN, EPI>::value, smem_bytes, st>>>(ta, tb, om, p);
  return check_launch("vb200_gemm");
}

template <int BN, int MODE, int EPI, int BKE = BK>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid,
                  cudaStream_t st) {
  if constexpr (EPI == VB200_EPI_F32) {  // fp32 output: no 16-bit epilogue operand, one instantiation serves both
    return launch_dt<BN, MODE, EPI, BKE, true>(ta, tb, p, grid, st);
  } else {
    return p.bf16 ? launch_dt<BN, MODE, EPI, BKE, true>(ta, tb, p, grid, st)
                  : launch_dt<BN, MODE, EPI, BKE, false>(ta, tb, p, grid, st);
  }
}

// K-major GEMM on CTA pairs (256-wide tiles): the epilogues of the decoder-stage GEMMs
static int launch_epi_pair(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid,
                           cudaStream_t st) {
  switch (epi) {
    case VB200_EPI_STORE: return launch<256, MODE_KMAJOR2, VB200_EPI_STORE>(ta, tb, p, grid, st);
    case VB200_EPI_DGELU_GRN: return launch<256, MODE_KMAJOR2, VB200_EPI_DGELU_GRN>(ta, tb, p, grid, st);
    case VB200_EPI_GELU_GP: return launch<256, MODE_KMAJOR2, VB200_EPI_GELU_GP>(ta, tb, p, grid, st);
  }
  return fail(VB200_ERR_INVALID, "no CTA-pair form of epilogue %d", epi);
}

template <int BN, bool MN_MAJOR>
static int launch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                      int grid, cudaStream_t st) {
  if constexpr (MN_MAJOR) {
    if (epi == VB200_EPI_F32) return launch<BN, MODE_MNMAJOR, VB200_EPI_F32>(ta, tb, p, grid, st);
    return fail(VB200_ERR_UNSUPPORTED, "mn_major GEMM supports EPI_F32 only");
  } else {
    switch (epi) {
      case VB200_EPI_STORE: return launch<BN, MODE_KMAJOR, VB200_EPI_STORE>(ta, tb, p, grid, st);
      case VB200_EPI_GELU_DUAL: return launch<BN, MODE_KMAJOR, VB200_EPI_GELU_DUAL>(ta, tb, p, grid, st);
      case VB200_EPI_DGELU: return launch<BN, MODE_KMAJOR, VB200_EPI_DGELU>(ta, tb, p, grid, st);
      case VB200_EPI_DGELU_GRN: return launch<BN, MODE_KMAJOR, VB200_EPI_DGELU_GRN>(ta, tb, p, grid, st);
      case VB200_EPI_GELU_GP: return launch<BN, MODE_KMAJOR, VB200_EPI_GELU_GP>(ta, tb, p, grid, st);
      case VB200_EPI_F32: return launch<BN, MODE_KMAJOR, VB200_EPI_F32>(ta, tb, p, grid, st);
    }
    return fail(VB200_ERR_INVALID, "unknown epilogue %d", epi);
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_gemm(const vb200_gemm_desc* d, vb200_stream_t stream) {
  VB_REQUIRE(d != nullptr, "null descriptor");
  VB_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "bad GEMM shape %d x %d x %d", d->M, d->N, d->K);
  VB_REQUIRE(d->A && d->B && d->out, "null operand pointer");
  VB_SUPPORTED(d->dtype == VB200_BF16 || d->dtype == VB200_FP16, "dtype %d", d->dtype);
  VB_SUPPORTED(d->N % 8 == 0, "N (%d) must be a multiple of 8", d->N);
  VB_SUPPORTED(d->ldo % 8 == 0, "ldo (%lld) must be a multiple of 8", (long long)d->ldo);
  const bool bf16 = d->dtype == VB200_BF16;
  const int epi = d->epilogue;
  if (epi == VB200_EPI_GELU_DUAL || epi == VB200_EPI_GELU_GP)
    VB_REQUIRE(d->out2 != nullptr && d->ldo2 % 8 == 0, "GELU_DUAL needs out2 with ldo2 %% 8 == 0");
  if (epi == VB200_EPI_DGELU)
    VB_REQUIRE(d->aux != nullptr && d->ldaux % 8 == 0, "DGELU needs aux with ldaux %% 8 == 0");
  if (d->residual) VB_REQUIRE(d->ldr % 8 == 0, "ldr %% 8");
  if (epi == VB200_EPI_DGELU_GRN && (d->svec || d->tvec))
    VB_REQUIRE(d->rows_per_sample % BM == 0, "rows_per_sample (%d) must be a multiple of %d", d->rows_per_sample, BM);
  if (epi == VB200_EPI_DGELU_GRN) VB_REQUIRE(d->bias == nullptr && d->k_splits <= 1, "DGELU_GRN: no bias, no K split");
  if (epi == VB200_EPI_DGELU_GRN)
    VB_REQUIRE(d->aux2 && d->ldaux2 % 8 == 0 && (!d->tvec || (d->aux && d->ldaux % 8 == 0)) &&
                   ((!d->tvec && !d->svec) || d->rows_per_sample > 0),
               "DGELU_GRN needs aux2 (gp); tvec needs aux (g); svec / tvec need rows_per_sample");
  const int nb = d->b_batch_rows > 0 ? (d->M + d->b_batch_rows - 1) / d->b_batch_rows : 1;
  if (d->b_batch_rows > 0)
    VB_REQUIRE(!d->mn_major && d->b_batch_rows % BM == 0, "b_batch_rows (%d) must be a multiple of %d (K-major only)",
               d->b_batch_rows, BM);
  int splits = d->k_splits > 0 ? d->k_splits : 1;
  VB_REQUIRE(splits == 1 || epi == VB200_EPI_F32, "k_splits > 1 needs EPI_F32");
  VB_REQUIRE(splits == 1 || d->atomic_out || d->split_out_stride > 0,
             "k_splits > 1 needs atomic_out or per-split slabs");

  // tile width: widest tile that still gives every SM work, else narrower
  const int sms = sm_count();
  const int tiles_m = (d->M + BM - 1) / BM;
  int bn = 256;
  if (d->N <= 64) bn = 64;
  else if (d->N <= 128) bn = 128;
  else {
    auto units = [&](int b) { return (long long)tiles_m * ((d->N + b - 1) / b) * splits; };
    if (units(256) < sms) bn = (units(128) < sms && d->N > 64) ? 64 : 128;
    // avoid > 1/3 of a 256-wide last tile being padding when the problem is tiny
  }
  const int tiles_n = (d->N + bn - 1) / bn;
  const int kb_total = (d->K + BK - 1) / BK;
  if (splits > kb_total) splits = kb_total;
  int kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per - 1) / kb_per;  // no empty split

  // CTA-pair form: K-major, 256-wide tiles, an even number of 128-row tiles, enough pair tiles for every SM pair
  static const int pair_mode = [] { const char* e = getenv("VB200_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  const bool pair = pair_mode && !d->mn_major && bn == 256 && splits == 1 && tiles_m % 2 == 0 && sms % 2 == 0 &&
                    (epi == VB200_EPI_STORE || epi == VB200_EPI_DGELU_GRN || epi == VB200_EPI_GELU_GP) &&
                    (long long)tiles_m * tiles_n >= sms && (d->b_batch_rows == 0 || d->b_batch_rows % (2 * BM) == 0) &&
                    (d->rows_per_sample == 0 || d->rows_per_sample % (2 * BM) == 0) && (d->rvec_rows == 0 || d->rvec_rows % (2 * BM) == 0);
  CUtensorMap ta, tb;
  int rc;
  if (!d->mn_major) {
    if ((rc = make_tmap_2d(&ta, d->A, d->M, d->K, d->lda, BK, BM, bf16))) return rc;
    if ((rc = make_tmap_2d(&tb, d->B, (long long)d->N * nb, d->K, d->ldb, BK, pair ? bn / 2 : bn, bf16))) return rc;
  } else {
    if ((rc = make_tmap_2d(&ta, d->A, d->K, d->M, d->lda, 64, BK, bf16))) return rc;
    if ((rc = make_tmap_2d(&tb, d->B, d->K, d->N, d->ldb, 64, BK, bf16))) return rc;
  }
  GemmParams p{};  // conv-only fields stay zero
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.tiles_m = tiles_m; p.tiles_n = tiles_n; p.k_splits = splits;
  p.kb_total = kb_total; p.kb_per_split = kb_per;
  p.bf16 = bf16 ? 1 : 0;
  p.act = d->act;
  p.atomic_out = d->atomic_out;
  p.ldo = d->ldo; p.ldo2 = d->ldo2; p.ldr = d->ldr; p.ldaux = d->ldaux; p.ldaux2 = d->ldaux2;
  p.b_batch_rows = d->b_batch_rows; p.rows_per_sample = d->rows_per_sample;
  if (d->rvec != nullptr) {
    VB_REQUIRE(epi == VB200_EPI_STORE && !d->mn_major && d->rvec_rows > 0, "rvec: K-major EPI_STORE with rvec_rows > 0");
    p.rvec = d->rvec;
    p.rvec_rows = d->rvec_rows;
  }
  if (d->colsq != nullptr) {
    VB_REQUIRE(epi == VB200_EPI_GELU_GP && !d->mn_major && d->rows_per_sample > 0 && d->rows_per_sample % 32 == 0,
               "colsq: EPI_GELU_GP with rows_per_sample %% 32 == 0 (an epilogue warp's 32 rows belong to one sample)");
    p.colsq = d->colsq;
  }
  if (d->n_split > 0) {
    VB_REQUIRE(epi == VB200_EPI_F32 && d->out2 != nullptr && d->n_split % 16 == 0 && d->n_split < d->N &&
                   d->ldo2 >= d->N - d->n_split && d->ldo2 % 4 == 0 && d->split_out_stride == 0,
               "n_split: EPI_F32 with out2, n_split %% 16 == 0, ldo2 >= N - n_split");
    p.n_split = d->n_split;
  }
  p.aux2 = d->aux2; p.tvec = d->tvec; p.svec = d->svec;
  p.split_out_stride = d->atomic_out ? 0 : d->split_out_stride;
  p.out = d->out; p.out2 = d->out2; p.bias = d->bias; p.residual = d->residual; p.aux = d->aux;
  const long long units = (long long)tiles_m * tiles_n * splits;
  const int grid = (int)(units < sms ? units : sms);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pair) return launch_epi_pair(epi, ta, tb, p, grid & ~1, st);
  if (!d->mn_major) {
    if (bn == 256) return launch_epi<256, false>(epi, ta, tb, p, grid, st);
    if (bn == 128) return launch_epi<128, false>(epi, ta, tb, p, grid, st);
    return launch_epi<64, false>(epi, ta, tb, p, grid, st);
  } else {
    if (bn == 256) return launch_epi<256, true>(epi, ta, tb, p, grid, st);
    if (bn == 128) return launch_epi<128, true>(epi, ta, tb, p, grid, st);
    return launch_epi<64, true>(epi, ta, tb, p, grid, st);
  }
}

// ------------------------------------------------------------------------------------- implicit-GEMM conv3d
struct ConvShape {
  int OD, OH, OW, taps;
  int stride[3];  // w, h, d
  long long pixels;
};
static int conv_shape(const vb200_conv3d_desc* d, ConvShape* s) {
  VB_REQUIRE(d != nullptr, "null descriptor");
  VB_REQUIRE(d->N > 0 && d->D > 0 && d->H > 0 && d->W > 0 && d->cin > 0 && d->cout > 0, "bad conv3d extent");
  VB_REQUIRE(d->kd > 0 && d->kh > 0 && d->kw > 0 && d->pd >= 0 && d->ph >= 0 && d->pw >= 0, "bad conv3d filter");
  VB_SUPPORTED(d->dtype == VB200_BF16 || d->dtype == VB200_FP16, "dtype %d", d->dtype);
  s->stride[0] = d->sw > 0 ? d->sw : 1;
  s->stride[1] = d->sh > 0 ? d->sh : 1;
  s->stride[2] = d->sd > 0 ? d->sd : 1;
  VB_SUPPORTED(s->stride[0] <= 8 && s->stride[1] <= 8 && s->stride[2] <= 8, "conv3d: stride > 8");
  VB_REQUIRE(d->D + 2 * d->pd >= d->kd && d->H + 2 * d->ph >= d->kh && d->W + 2 * d->pw >= d->kw, "conv3d: empty output");
  VB_REQUIRE(d->xd >= 0 && d->xh >= 0 && d->xw >= 0, "conv3d: negative extra extent");
  s->OD = (d->D + 2 * d->pd - d->kd) / s->stride[2] + 1 + d->xd;
  s->OH = (d->H + 2 * d->ph - d->kh) / s->stride[1] + 1 + d->xh;
  s->OW = (d->W + 2 * d->pw - d->kw) / s->stride[0] + 1 + d->xw;
  s->taps = d->kd * d->kh * d->kw;
  s->pixels = (long long)d->N * s->OD * s->OH * s->OW;
  VB_SUPPORTED(s->pixels < (1LL << 31) - 256, "conv3d: too many output voxels");
  VB_SUPPORTED(d->cout % 8 == 0, "conv3d: cout (%d) must be a multiple of 8", d->cout);
  return VB200_OK;
}

static void conv_geom(GemmParams* p, const vb200_conv3d_desc* d, const ConvShape& s) {
  p->cOW = s.OW; p->cOH = s.OH; p->cOD = s.OD;
  p->cKW = d->kw; p->cKH = d->kh;
  p->cpw = d->pw; p->cph = d->ph; p->cpd = d->pd;
  p->csw = s.stride[0]; p->csh = s.stride[1]; p->csd = s.stride[2];
  p->ccin = d->cin;
}

// box of `rows` consecutive output voxels, whose (strided) input box must also respect TMA's 256-element box limit
static bool conv_box_strided(int rows, const ConvShape& s, int NB, int* box) {
  if (!conv_box(rows, s.OW, s.OH, s.OD, NB, box)) return false;
  return box[0] * s.stride[0] <= 256 && box[1] * s.stride[1] <= 256 && box[2] * s.stride[2] <= 256;
}

extern "C" int vb200_conv3d_igemm_supported(const vb200_conv3d_desc* d, int wgrad) {
  ConvShape s;
  if (conv_shape(d, &s)) return 0;
  int box[4];
  if (wgrad) return d->cin % 8 == 0 && conv_box_strided(BK, s, d->N, box) ? 1 : 0;
  return d->cin % 32 == 0 && conv_box_strided(BM, s, d->N, box) ? 1 : 0;
}

extern "C" int vb200_conv3d_igemm(const vb200_conv3d_desc* d, vb200_stream_t stream) {
  ConvShape s;
  if (int rc = conv_shape(d, &s)) return rc;
  VB_REQUIRE(d->x && d->w && d->out, "null operand pointer");
  VB_SUPPORTED(d->cin % 32 == 0, "conv3d_igemm: cin (%d) must be a multiple of 32", d->cin);
  int box[4];
  VB_SUPPORTED(conv_box_strided(BM, s, d->N, box),
               "conv3d_igemm: 128 consecutive output voxels of %dx%dx%d are not a box", s.OD, s.OH, s.OW);
  const bool bf16 = d->dtype == VB200_BF16;
  const int bke = d->cin % 64 == 0 ? 64 : 32;
  const int sms = sm_count();
  const int M = (int)s.pixels, N = d->cout, K = s.taps * d->cin;
  const int wtaps = d->w_taps > 0 ? d->w_taps : s.taps;  // taps held by the weight matrix (tap selection: > s.taps)
  VB_REQUIRE(d->tapmap == nullptr || (s.taps <= 27 && d->w_taps > 0), "conv3d_igemm: tapmap needs w_taps and <= 27 taps");
  const bool scatter = d->out_pitch[0] != 0;
  VB_REQUIRE(!scatter || d->residual == nullptr, "conv3d_igemm: scattered output excludes the residual operand");
  const int tiles_m = (M + BM - 1) / BM;
  int bn = 256;
  if (N <= 64) bn = 64;
  else if (N <= 128) bn = 128;
  else if ((long long)tiles_m * ((N + 255) / 256) < sms) bn = (long long)tiles_m * ((N + 127) / 128) < sms ? 64 : 128;
  CUtensorMap ta, tb;
  if (int rc = make_tmap_conv(&ta, d->x, d->N, d->D, d->H, d->W, d->cin, bke, box, s.stride, bf16, bke == 32)) return rc;
  if (int rc = make_tmap_2d(&tb, d->w, N, (long long)wtaps * d->cin, (long long)wtaps * d->cin, bke, bn, bf16, bke == 32))
    return rc;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  if (d->tapmap != nullptr) {
    p.ntap = s.taps;
    for (int i = 0; i < s.taps; ++i) {
      VB_REQUIRE(d->tapmap[i] >= 0 && d->tapmap[i] < wtaps, "conv3d_igemm: tapmap[%d] out of range", i);
      p.ctap[i] = d->tapmap[i];
    }
  }
  if (scatter) {
    p.opx = d->out_pitch[0]; p.opy = d->out_pitch[1]; p.opz = d->out_pitch[2]; p.opn = d->out_pitch[3];
    VB_REQUIRE(p.opx % 8 == 0 && p.opy % 8 == 0 && p.opz % 8 == 0 && p.opn % 8 == 0, "conv3d_igemm: out_pitch %% 8");
  }
  p.tiles_m = tiles_m; p.tiles_n = (N + bn - 1) / bn; p.k_splits = 1;
  p.kb_total = K / bke; p.kb_per_split = p.kb_total;
  p.bf16 = bf16 ? 1 : 0;
  p.act = d->act;
  p.ldo = d->ldo > 0 ? d->ldo : N;
  p.ldr = d->ldr > 0 ? d->ldr : N;
  VB_REQUIRE(p.ldo % 8 == 0 && p.ldr % 8 == 0, "conv3d_igemm: ldo / ldr must be multiples of 8");
  p.out = d->out; p.bias = d->bias; p.residual = d->residual;
  conv_geom(&p, d, s);
  p.cchunks = d->cin / bke;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // patch form (see MODE_CONVKP): 3-row filters at stride 1 on extents in whole 16 x 8 patches, up to 128 output channels
  const bool patch = d->kh == 3 && s.stride[0] == 1 && s.stride[1] == 1 && s.stride[2] == 1 && bn <= 128 &&
                     s.OW % 16 == 0 && s.OH % 8 == 0 && d->xh == 0 && d->xw == 0 && d->xd == 0 && !scatter &&
                     d->tapmap == nullptr && d->residual == nullptr;
  if (patch) {
    const int pbox[4] = {16, 10, 1, 1};
    const int one[3] = {1, 1, 1};
    if (int rc = make_tmap_conv(&ta, d->x, d->N, d->D, d->H, d->W, d->cin, bke, pbox, one, bf16, bke == 32)) return rc;
    p.cpxn = s.OW / 16;
    p.cpyn = s.OH / 8;
    p.tiles_m = d->N * s.OD * p.cpyn * p.cpxn;
    p.kb_total = d->kd * d->kw * p.cchunks;
    p.kb_per_split = p.kb_total;
    const long long punits = (long long)p.tiles_m * p.tiles_n;
    const int pgrid = (int)(punits < sms ? punits : sms);
    // the whole filter fits next to the A ring: keep it resident (one fetch per CTA instead of three sub-tiles per K block)
    const int wrows = N <= 32 ? 32 : 64;
    const long long wbytes = (long long)s.taps * p.cchunks * wrows * bke * 2;
    if (bn == 64 && wbytes <= 112 * 1024 && p.tiles_m >= 2 * pgrid) {
      p.cwbytes = (int)wbytes;
      p.cwrows = wrows;
      if (int rc = make_tmap_2d(&tb, d->w, N, K, K, bke, wrows, bf16, bke == 32)) return rc;
      return bke == 64 ? launch<64, MODE_CONVKPW, VB200_EPI_STORE, 64>(ta, tb, p, pgrid, st)
                       : launch<64, MODE_CONVKPW, VB200_EPI_STORE, 32>(ta, tb, p, pgrid, st);
    }
    if (bke == 64) {
      // CTA pairs (two x-adjacent patches per MMA, each CTA stages half of the filter rows): shared-memory operand reads per
      // MMA drop from A + B to A + B / 2 - these <= 128-channel tiles are bound by exactly that
      static const int conv_pair = [] { const char* e = getenv("VB200_CONV_PAIR"); return e ? atoi(e) : 1; }();
      if (conv_pair && p.tiles_n == 1 && p.cpxn % 2 == 0 && p.tiles_m % 2 == 0 && sms % 2 == 0 && punits >= sms && N % 16 == 0 &&
          N == bn) {
        if (int rc = make_tmap_2d(&tb, d->w, N, (long long)wtaps * d->cin, (long long)wtaps * d->cin, bke, bn / 2, bf16, false))
          return rc;
        if (bn == 128) return launch<128, MODE_CONVKP2, VB200_EPI_STORE, 64>(ta, tb, p, pgrid & ~1, st);
        return launch<64, MODE_CONVKP2, VB200_EPI_STORE, 64>(ta, tb, p, pgrid & ~1, st);
      }
      if (bn == 128) return launch<128, MODE_CONVKP, VB200_EPI_STORE, 64>(ta, tb, p, pgrid, st);
      return launch<64, MODE_CONVKP, VB200_EPI_STORE, 64>(ta, tb, p, pgrid, st);
    }
    if (bn == 128) return launch<128, MODE_CONVKP, VB200_EPI_STORE, 32>(ta, tb, p, pgrid, st);
    return launch<64, MODE_CONVKP, VB200_EPI_STORE, 32>(ta, tb, p, pgrid, st);
  }
  const long long units = (long long)p.tiles_m * p.tiles_n;
  const int grid = (int)(units < sms ? units : sms);
  if (bke == 64) {
    if (bn == 256) return launch<256, MODE_CONVK, VB200_EPI_STORE, 64>(ta, tb, p, grid, st);
    if (bn == 128) return launch<128, MODE_CONVK, VB200_EPI_STORE, 64>(ta, tb, p, grid, st);
    return launch<64, MODE_CONVK, VB200_EPI_STORE, 64>(ta, tb, p, grid, st);
  }
  if (bn == 256) return launch<256, MODE_CONVK, VB200_EPI_STORE, 32>(ta, tb, p, grid, st);
  if (bn == 128) return launch<128, MODE_CONVK, VB200_EPI_STORE, 32>(ta, tb, p, grid, st);
  return launch<64, MODE_CONVK, VB200_EPI_STORE, 32>(ta, tb, p, grid, st);
}

extern "C" int vb200_conv3d_igemm_wgrad(const vb200_conv3d_desc* d, vb200_stream_t stream) {
  ConvShape s;
  if (int rc = conv_shape(d, &s)) return rc;
  VB_REQUIRE(d->x && d->dout && d->dw, "null operand pointer");
  VB_SUPPORTED(d->cin % 8 == 0, "conv3d_igemm_wgrad: cin (%d) must be a multiple of 8", d->cin);
  int box[4];
  VB_SUPPORTED(conv_box_strided(BK, s, d->N, box),
               "conv3d_igemm_wgrad: 64 consecutive output voxels of %dx%dx%d are not a box", s.OD, s.OH, s.OW);
  const bool bf16 = d->dtype == VB200_BF16;
  const int sms = sm_count();
  const int bn = d->cin >= 256 ? 256 : (d->cin >= 128 ? 128 : 64);
  const int ctiles = (d->cin + bn - 1) / bn;
  CUtensorMap ta, tb;
  if (int rc = make_tmap_2d(&ta, d->dout, s.pixels, d->cout, d->cout, 64, BK, bf16)) return rc;
  if (int rc = make_tmap_conv(&tb, d->x, d->N, d->D, d->H, d->W, d->cin, 64, box, s.stride, bf16, false)) return rc;
  GemmParams p{};
  p.M = d->cout; p.N = s.taps * d->cin; p.K = (int)s.pixels;
  p.tiles_m = (d->cout + BM - 1) / BM; p.tiles_n = s.taps * ctiles;
  p.kb_total = (p.K + BK - 1) / BK;
  const long long tiles = (long long)p.tiles_m * p.tiles_n;
  // K-split so that the units fill (at most) two whole waves of CTAs: a third partial wave would cost 50 %
  int splits = d->k_splits > 0 ? d->k_splits : (int)((2LL * sms) / tiles);
  if (splits > p.kb_total / 8) splits = p.kb_total / 8;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.k_splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.bf16 = bf16 ? 1 : 0;
  p.atomic_out = 1;
  p.ldo = p.N;
  p.out = d->dw;
  conv_geom(&p, d, s);
  p.cchunks = ctiles;
  p.cbx = box[0]; p.cby = box[1]; p.cbz = box[2]; p.cbn = box[3];
  const long long units = tiles * p.k_splits;
  const int grid = (int)(units < sms ? units : sms);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bn == 256) return launch<256, MODE_CONVMN, VB200_EPI_F32>(ta, tb, p, grid, st);
  if (bn == 128) return launch<128, MODE_CONVMN, VB200_EPI_F32>(ta, tb, p, grid, st);
  return launch<64, MODE_CONVMN, VB200_EPI_F32>(ta, tb, p, grid, st);
}

extern "C" int vb200_last_error(char* buf, size_t n) {
  if (buf == nullptr || n == 0) return VB200_ERR_INVALID;
  snprintf(buf, n, "%s", g_err);
  return VB200_OK;
}
extern "C" int vb200_abi_version(void) { return 1; }
extern "C" int64_t vb200_launch_count(void) { return g_launches.load(); }
