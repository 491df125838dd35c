/* viscy_b200 C ABI -- the drop-in boundary for the VisCy convolutional hot path on B200 (sm_100a).
 *
 * The reference (mehta-lab/VisCy) has NO native layer: its hot path is torch.nn modules dispatching
 * to cuDNN/cuBLAS/ATen.  Every entry point below therefore replaces a *library
 * call site* of the reference, cited per function as packages/viscy-models/src/viscy_models/<file>:<line>
 * (abbreviated VM/...) or as the timm/monai class the reference composes.
 *
 * Conventions
 *  - plain pointers + sizes only; all device buffers are owned by the caller (torch tensors);
 *    the library never allocates or frees device memory and keeps no pointer past return.
 *  - activations are channels-last: [B,H,W,C] / [B,D,H,W,C], 16-bit (bf16 = 0, fp16 = 1);
 *    parameters and statistics are fp32 unless stated.
 *  - every launch goes to the cudaStream_t passed in; no host synchronisation; CUDA-graph capturable.
 *  - return 0 on success; non-zero = error, message via vb200_last_error().  Unsupported
 *    shape/dtype is an error (VB200_ERR_UNSUPPORTED), never a silent fallback.
 */
#ifndef VISCY_B200_H
#define VISCY_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vb200_stream_t; /* cudaStream_t */

enum { VB200_OK = 0, VB200_ERR_INVALID = 1, VB200_ERR_UNSUPPORTED = 2, VB200_ERR_CUDA = 3 };
enum { VB200_BF16 = 0, VB200_FP16 = 1 };

/* GEMM epilogues (fused into the tcgen05 kernel's TMEM->register drain) */
enum {
  VB200_EPI_STORE = 0,     /* out = act(acc + bias[col]) (+ residual[row,col]) -> 16-bit          */
  VB200_EPI_GELU_DUAL = 1, /* u = acc + bias; out = u; out2 = gelu(u); optional sum(g^2) partials */
  VB200_EPI_DGELU = 2,     /* out = (acc + bias) * gelu'(u), u = aux[row,col]  (dgrad through GELU)      */
  VB200_EPI_F32 = 3        /* out(fp32) = acc (+ bias[col]); optional atomic accumulate / K-split slabs */
};
enum { VB200_ACT_NONE = 0, VB200_ACT_RELU = 1, VB200_ACT_GELU = 2 };

/* D[M,N] = A[M,K] * B[N,K]^T  (mn_major = 0; both operands K-major: forward / dgrad form), or
 * D[M,N] = sum_k At[k,M] * Bt[k,N] (mn_major = 1; both operands MN-major: wgrad form, k = pixels).
 * Replaces: nn.Linear / 1x1 nn.Conv2d of timm ConvNeXtBlock.mlp (VM/components/blocks.py:60-74,
 * VM/unet/unext2.py:40-49), the k==stride patchify convs (VM/components/stems.py:26-50,
 * timm ConvNeXtStage.downsample) and their autograd dgrad/wgrad (cuBLAS / cuDNN today). */
typedef struct vb200_gemm_desc {
  int32_t M, N, K;
  int32_t dtype;         /* VB200_BF16 | VB200_FP16 : type of A, B and 16-bit outputs */
  int32_t mn_major;      /* 0: A[M,K], B[N,K] row-major; 1: A[K,M], B[K,N] row-major */
  int32_t epilogue;      /* VB200_EPI_* */
  int32_t act;           /* VB200_ACT_* (EPI_STORE only) */
  int32_t k_splits;      /* >=1; mn_major wgrad form: split the K (pixel) range */
  int32_t atomic_out;    /* EPI_F32: 1 = red.add into out, 0 = plain store */
  int64_t lda, ldb;      /* leading dimensions in elements */
  int64_t ldo, ldo2, ldr, ldaux; /* leading dims (elements) of out, out2, residual, aux */
  int64_t split_out_stride;      /* EPI_F32: elements between per-split output slabs (0 = same slab) */
  const void* A;
  const void* B;
  void* out;
  void* out2;
  const float* bias;     /* [N] or NULL */
  const void* residual;  /* 16-bit [M,ldr] or NULL */
  const void* aux;       /* EPI_DGELU: u (pre-activation), 16-bit [M,ldaux] */
} vb200_gemm_desc;

int vb200_gemm(const vb200_gemm_desc* d, vb200_stream_t stream);

/* copies the last error message of the calling thread into buf (NUL terminated) */
int vb200_last_error(char* buf, size_t n);
/* library / kernel ABI version, bumped on any struct change */
int vb200_abi_version(void);
/* number of kernel launches issued by this library in the calling process since load */
int64_t vb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VISCY_B200_H */
