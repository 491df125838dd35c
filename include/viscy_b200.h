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
enum { VB200_BF16 = 0, VB200_FP16 = 1, VB200_FP32 = 2 /* only where an entry point says so (loss inputs) */ };

/* GEMM epilogues (fused into the tcgen05 kernel's TMEM->register drain) */
enum {
  VB200_EPI_STORE = 0,     /* out = act((acc + bias[col]) * s[col]) (+ residual[row,col]) -> 16-bit; s = svec or 1 */
  VB200_EPI_GELU_DUAL = 1, /* u = acc + bias; out = u; out2 = gelu(u); optional sum(g^2) partials */
  VB200_EPI_DGELU = 2,     /* out = (acc + bias) * gelu'(u), u = aux[row,col]  (dgrad through GELU)      */
  VB200_EPI_F32 = 3,       /* out(fp32) = acc (+ bias[col]); optional atomic accumulate / K-split slabs */
  VB200_EPI_DGELU_GRN = 4, /* out = (acc*s[n,col] + g*t[n,col]) * gp: g = aux, gp = aux2 (= gelu'(u) from EPI_GELU_GP),
                              n = row / rows_per_sample; s, t optional (NULL: 1 and 0) */
  VB200_EPI_GELU_GP = 5    /* u = acc + bias; out = gelu'(u); out2 = gelu(u) */
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
  int32_t b_batch_rows;  /* > 0 (K-major only, multiple of 128): B is [nb][N][K], rows [i*b_batch_rows, ..) of D use B[i]
                            (per-sample GRN-scaled fc2 weights) */
  int32_t rows_per_sample; /* EPI_DGELU_GRN: rows per sample */
  int64_t lda, ldb;      /* leading dimensions in elements */
  int64_t ldo, ldo2, ldr, ldaux, ldaux2; /* leading dims (elements) of out, out2, residual, aux, aux2 */
  int64_t split_out_stride;      /* EPI_F32: elements between per-split output slabs (0 = same slab) */
  const void* A;
  const void* B;
  void* out;
  void* out2;
  const float* bias;     /* [N] or NULL */
  const void* residual;  /* 16-bit [M,ldr] or NULL */
  const void* aux;       /* EPI_DGELU: u (pre-activation), 16-bit [M,ldaux]; EPI_DGELU_GRN: g = gelu(u) */
  const void* aux2;      /* EPI_DGELU_GRN: gp = gelu'(u), 16-bit [M,ldaux2] */
  const float* tvec;     /* EPI_DGELU_GRN: t [nsamples, N] fp32 or NULL */
  const float* svec;     /* EPI_DGELU_GRN: s [nsamples, N] fp32 or NULL; EPI_STORE: per-column scale [N] or NULL */
  int32_t n_split;       /* EPI_F32, optional (multiple of 16, < N): columns >= n_split are written to out2 (fp32, row pitch
                            ldo2) at column - n_split.  With B = [l | 1] the weight-gradient GEMM writes dW to out and the
                            bias gradient (the ones column) to out2 */
  int32_t rvec_rows;     /* rows per sample for rvec */
  const float* rvec;     /* EPI_STORE, optional: per-sample fp32 scale applied to (acc + bias) * s before the residual is
                            added, row r uses rvec[r / rvec_rows] (timm DropPath / stochastic depth on the residual branch) */
  float* colsq;          /* EPI_GELU_GP, optional: fp32 [M / rows_per_sample, N] (pre-zeroed), accumulates sum over each sample's
                            rows of out2^2 (the 16-bit-rounded GELU output) = the GRN statistic, so that no separate pass over the
                            hidden tensor is needed.  Needs rows_per_sample % 32 == 0 (an epilogue warp's rows lie in one sample) */
} vb200_gemm_desc;

int vb200_gemm(const vb200_gemm_desc* d, vb200_stream_t stream);

/* ---- implicit-GEMM 3-D convolution on the same tcgen05 pipeline (any filter extent, stride <= 8, zero padding).
 * im2col is folded into TMA: the channels-last activation is one 5-D tensor map (C, X, Y, Z, N) and each K block of
 * the contraction -- (filter tap, 64- or 32-channel chunk) -- is one box fetched at the tap-shifted voxel coordinate,
 * out-of-range voxels zero-filled by the TMA unit (= the padding).  No patch matrix ever exists in HBM.
 *   forward : out[v, co] = act(bias[co] + sum_{tap,ci} x[v*s + tap - pad, ci] * w[co, tap, ci]) (+ residual[v, co])
 *   dgrad   : (stride 1) the same entry point on dout with the flipped / transposed filter and padding k-1-p
 *   wgrad   : dw[co, tap, ci] += sum_v dout[v, co] * x[v*s + tap - pad, ci]   (fp32, red.add; K-split over voxels)
 * Strided convolutions use the tensor map's element strides: a box of s*b voxels per dimension delivers b voxels.
 * The data gradient of nn.ConvTranspose3d is the strided forward of its adjoint conv, its weight gradient the strided
 * wgrad with the roles of x and dout exchanged.  nn.ConvTranspose3d forward (and the data gradient of a strided conv)
 * is one forward launch per output parity class r = o mod s: a stride-1 conv over the taps {k : (k - r - p) mod s == 0}
 * (tapmap into the full filter), written to the strided sub-lattice of the output (out_pitch) -- no column matrix.
 * Replaces nn.Conv3d(k=3, padding=1) of Block.proj (VM/unet/blocks.py:88-113), ResnetBlock / ConvBottleneck3D
 * (VM/unet/blocks.py:116-188,233-292), UNet3DBase inconv / outconv (VM/unet/unet3d_base.py:90-138), the padded Conv3d
 * of ConvBlock3D (VM/components/conv_block_3d.py:261-274) and their autograd dgrad / wgrad (cuDNN today). */
typedef struct vb200_conv3d_desc {
  int32_t N, D, H, W;   /* input extent in voxels */
  int32_t cin;          /* channels of x (= its row pitch); forward: multiple of 32, wgrad: multiple of 8 */
  int32_t cout;         /* output channels (= row pitch of dout and, unless ldo is set, of out); multiple of 8 */
  int32_t kd, kh, kw;   /* filter extent */
  int32_t pd, ph, pw;   /* zero padding; output extent = (in + 2p - k) / s + 1 */
  int32_t sd, sh, sw;   /* stride (0 is read as 1) */
  int32_t dtype;        /* VB200_BF16 | VB200_FP16 */
  int32_t act;          /* VB200_ACT_* applied to the forward output */
  int32_t k_splits;     /* wgrad: split of the voxel range (0 = pick) */
  int32_t w_taps;       /* forward: taps held by w when a tapmap selects among them (0 = kd*kh*kw) */
  int32_t xd, xh, xw;   /* extra output extent per dimension (asymmetric high-side padding); 0 */
  int32_t reserved;
  int64_t ldo, ldr;     /* row pitch (elements) of out / residual; 0 = cout */
  int64_t out_pitch[4]; /* forward, optional: output voxel (n,z,y,x) is written at out + n*p[3] + z*p[2] + y*p[1] + x*p[0]
                           (elements): a transposed conv runs as one launch per output parity class.  0 = dense rows */
  const int32_t* tapmap; /* forward, optional (host memory): tap t of this launch uses weight columns of tap tapmap[t] */
  const void* x;        /* [N,D,H,W,cin] 16-bit */
  const void* w;        /* forward: [cout, kd*kh*kw*cin] 16-bit, K order (kd, kh, kw, ci) */
  const float* bias;    /* [cout] or NULL */
  const void* residual; /* [N,OD,OH,OW,ldr] 16-bit or NULL */
  void* out;            /* forward: [N,OD,OH,OW,ldo] 16-bit */
  const void* dout;     /* wgrad: [N,OD,OH,OW,cout] 16-bit */
  float* dw;            /* wgrad: fp32 [cout, kd*kh*kw*cin], accumulated (pre-zeroed by the caller) */
} vb200_conv3d_desc;
int vb200_conv3d_igemm(const vb200_conv3d_desc* d, vb200_stream_t stream);
int vb200_conv3d_igemm_wgrad(const vb200_conv3d_desc* d, vb200_stream_t stream);
/* host-only geometry query (no launch): 1 when the forward (wgrad = 0) / weight-gradient (wgrad = 1) form applies */
int vb200_conv3d_igemm_supported(const vb200_conv3d_desc* d, int wgrad);
/* weight gradient for few-channel stride-1 convs with kh == 3 (same descriptor, same dw layout): a K block is an 8 x 8
 * voxel patch, the x patch carries a y halo and the three kh taps are row-shifted views of that one box, accumulated in
 * three TMEM accumulators (conv3d_wgrad_sm100.cu).  Needs OH % 8 == 0 and OW % 8 == 0. */
int vb200_conv3d_wgrad_kh3(const vb200_conv3d_desc* d, vb200_stream_t stream);
int vb200_conv3d_wgrad_kh3_supported(const vb200_conv3d_desc* d);

/* ---- ConvNeXt / ConvNeXt-V2 block pieces (timm ConvNeXtBlock as composed by VM/unet/unext2.py:40-49,
 * VM/components/blocks.py:54-74, VM/contrastive/encoder.py:93-99).  All activations channels-last 16-bit. ---- */

/* depthwise 7x7, pad 3: y = dwconv(x; wt) (+ bias) (+ add).  wt is tap-major fp32 [49][C].
 * forward: wt = conv_dw.weight^T, bias = conv_dw.bias.  dgrad: wt = flipped taps, add = residual gradient. */
int vb200_dwconv7(const void* x, const float* wt, const float* bias, const void* add, void* y,
                  int B, int H, int W, int C, int dtype, vb200_stream_t stream);
/* dwt[49][C] += sum_pixels dy * shifted x;  db[C] += sum dy  (outputs pre-zeroed by the caller) */
int vb200_dwconv7_wgrad(const void* x, const void* dy, float* dwt, float* db, int B, int H, int W, int C,
                        int dtype, vb200_stream_t stream);

/* LayerNorm over the channel dim of [M, C] rows (timm LayerNorm / LayerNorm2d, eps 1e-6) */
int vb200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                        float* rstd, int64_t M, int C, float eps, int dtype, vb200_stream_t stream);
/* same with a row pitch ldy (elements) for y.  ones = n > 0: y[:, C:C+8n] = {1,0,..,0} (ldy >= C + 8 n); ones2 != NULL:
 * the same columns at ones2[row * ld2 + col2 ..].  A weight-gradient GEMM against [l | 1] then yields the bias gradient as an
 * extra column, which replaces the column-sum passes over dh / dout (C % 8 == 0, C <= 2048) */
int vb200_layernorm_fwd_ld(const void* x, const float* gamma, const float* beta, void* y, int64_t ldy, int ones,
                           void* ones2, int64_t ld2, int col2, float* mean, float* rstd, int64_t M, int C, float eps,
                           int dtype, vb200_stream_t stream);
/* dgamma, dbeta are accumulated into (pre-zeroed by the caller) */
int vb200_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                        const float* gamma, void* dx, float* dgamma, float* dbeta, int64_t M, int C,
                        int dtype, vb200_stream_t stream);

/* GELU + GlobalResponseNorm (timm GlobalResponseNormMlp) on h [B, R, C]:
 *   sumsq[n,c] += sum_r gelu(h)^2 ; s[n,c] = 1 + w[c] * Gx / (mean_c Gx + eps), Gx = sqrt(sumsq) ;
 *   y = gelu(h) * s[n,c] + b[c]   (== x + addcmul(bias, weight, x * Nx)) */
int vb200_grn_sumsq(const void* h, float* sumsq, int B, int R, int C, int dtype, vb200_stream_t stream);
int vb200_grn_coef_fwd(const float* sumsq, const float* w, float* s, int B, int C, float eps, vb200_stream_t stream);
int vb200_grn_apply_fwd(const void* h, const float* s, const float* b, void* y, int B, int R, int C, int dtype,
                        vb200_stream_t stream);
/* backward: S1[n,c] += sum_r dy*gelu(h), sdy[c] += sum dy (= d grn.bias);  t from S1 (and d grn.weight += ...);
 *   dh = (dy * s + gelu(h) * t) * gelu'(h), dbias[c] += sum_r dh (fc1 bias gradient; may be NULL).
 *   With t == NULL and s == ones the apply kernel is a plain GELU backward. */
int vb200_grn_bwd_reduce(const void* h, const void* dy, float* S1, float* sdy, int B, int R, int C, int dtype,
                         vb200_stream_t stream);
int vb200_grn_coef_bwd(const float* sumsq, const float* S1, const float* w, float* t, float* dw, int B, int C,
                       float eps, vb200_stream_t stream);
int vb200_grn_apply_bwd(const void* h, const void* dy, const float* s, const float* t, void* dh, float* dbias,
                        int B, int R, int C, int dtype, vb200_stream_t stream);

/* ---- fused GRN path (pixels-per-sample % 128 == 0): the GRN scale is folded into per-sample fc2 weights so the
 * hidden tensor y = GRN(g) is never materialised, and the GRN + GELU backward rides in the fc2-dgrad epilogue ---- */
/* mode 0: out[n,c] += sum_r x[n,r,c]; mode 1: sum of squares; mode 2: both in one pass (out [2,B,C]: sums, then sums of
 * squares -- BatchNorm statistics).  x [B,R,C] 16-bit, C % 8 == 0, out pre-zeroed */
int vb200_colreduce(const void* x, float* out, int B, int64_t R, int C, int mode, int dtype, vb200_stream_t stream);
/* same on rows with pitch ld >= C (elements, % 8); pivot (fp32 [C] or NULL) is subtracted from every element before
 * summing (shifted one-pass BatchNorm statistics: mean = p + S1/M, var = S2/M - (S1/M)^2) */
int vb200_colreduce_ld(const void* x, float* out, int B, int64_t R, int C, int64_t ld, int mode, const float* pivot,
                       int dtype, vb200_stream_t stream);
/* out[n][j][k] = W2[j][k] * s[n][k] (16-bit) */
int vb200_grn_pack_w2(const float* W2, const float* s, void* out, int nb, int C, int C4, int dtype,
                      vb200_stream_t stream);
/* per-sample scaled weights (as vb200_grn_pack_w2) and the effective bias (as vb200_grn_bias_eff) in one launch */
int vb200_grn_prepare(const float* s, const float* gb, const float* W2, const float* b2, void* w2s, float* b2e, int nb,
                      int C, int C4, int dtype, vb200_stream_t stream);
/* out[j] = b2[j] + sum_k W2[j][k] * bgrn[k] */
int vb200_grn_bias_eff(const float* W2, const float* bgrn, const float* b2, float* out, int C, int C4, vb200_stream_t stream);
/* from per-sample wgrad partials P[n][j][k]: dW2 (overwritten), S1 [nb,C4] and dbgrn [C4] (accumulated, pre-zeroed) */
int vb200_grn_wgrad_finish(const float* P, const float* W2, const float* s, const float* bgrn, const float* db2,
                           float* dW2, float* S1, float* dbgrn, int nb, int C, int C4, vb200_stream_t stream);
/* same with a row pitch ldp for P; db2_in == NULL: the fc2 bias gradient is column C4 of P summed over samples (the
 * ones column of the GELU output buffer) and is written to db2_out [C] */
int vb200_grn_wgrad_finish_ld(const float* P, int64_t ldp, const float* W2, const float* s, const float* bgrn,
                              const float* db2_in, float* dW2, float* S1, float* dbgrn, float* db2_out, int nb, int C,
                              int C4, vb200_stream_t stream);
/* ConvNeXt-V1 layer scale, backward side (timm ConvNeXtBlock `x = x * gamma`, VM/contrastive/encoder.py:93-124 via
 * convnext_tiny) in one launch: G [C, ldg] fp32 = dout^T y, W2 [C, C4], gamma / b2 / db_raw [C] ->
 * dW2 = G * gamma[:, None]; db2 = db_raw * gamma; dgamma (pre-zeroed) += sum_j G * W2 + db_raw * b2;
 * w2t [C4, C] 16-bit = (W2 * gamma / max|gamma|)^T (operand of the fc2 data gradient); sv [C4] = max|gamma|.
 * db_raw / db2 may be NULL (no bias). */
int vb200_layerscale_bwd(const float* G, int64_t ldg, const float* W2, const float* gamma, const float* b2,
                         const float* db_raw, void* w2t, float* dgamma, float* dW2, float* db2, float* sv, int C, int C4,
                         int dtype, vb200_stream_t stream);
/* vb200_grn_coef_fwd + vb200_grn_prepare in one launch: s (fp32 [nb,C4], written), w2s, b2e from sumsq */
int vb200_grn_prepare2(const float* sumsq, const float* gw, const float* gb, const float* W2, const float* b2,
                       float* s_out, void* w2s, float* b2e, int nb, int C, int C4, float eps, int dtype,
                       vb200_stream_t stream);

/* ---- GroupNorm (+ timestep scale / shift) + activation on rows [N, R, C] (VM/unet/blocks.py:88-113 with the UNet3DBase
 * defaults norm="group", activation="silu", VM/unet/unet3d_base.py:58-72).  Statistics come from vb200_colreduce (mode 2,
 * per sample); the per-(sample, channel) coefficients are formed by the caller:
 *   y = act(a[n,c] * x + b[n,c]),  act: 0 none, 1 ReLU, 2 SiLU, 3 LeakyReLU(0.01), 4 ELU(1), 5 SELU (the activations of
 *   ConvBlock3D, VM/components/conv_block_3d.py:213-229; with a = Dropout3d scale, b = 0 this is its dropout + activation)
 *   backward: dv = dy * act'(a x + b);  s1[n,c] += sum_r dv, s2[n,c] += sum_r dv * x  (pre-zeroed);
 *             dx = c1 * dv + c2 * x + c3 with coef = five fp32 [N, C] planes (a, b, c1, c2, c3) ---- */
int vb200_affine_nc_act(const void* x, const float* a, const float* b, void* y, int64_t N, int64_t R, int C, int act,
                        int dtype, vb200_stream_t stream);
int vb200_gn_bwd_reduce(const void* dy, const void* x, const float* a, const float* b, float* s1, float* s2, int64_t N,
                        int64_t R, int C, int act, int dtype, vb200_stream_t stream);
int vb200_gn_bwd_apply(const void* dy, const void* x, const float* coef, void* dx, int64_t N, int64_t R, int C, int act,
                       int dtype, vb200_stream_t stream);

/* out[c] += sum_rows x[r][c]  (bias gradients; out pre-zeroed) */
int vb200_colsum(const void* x, float* out, int64_t M, int C, int dtype, vb200_stream_t stream);

/* ---- layout kernels ---- */
/* monai SubpixelUpsample(pre_conv=None, scale 2) + torch.cat([up, skip], 1) (VM/components/blocks.py:137-172):
 * out[n,2h+i,2w+j,c'] = prev[n,h,w,c'*4+i*2+j] (c' < Cp/4) else skip[n,2h+i,2w+j,c'-Cp/4] */
int vb200_pixshuf_cat_fwd(const void* prev, const void* skip, void* out, int B, int h, int w, int Cp, int Cs,
                          vb200_stream_t stream);
int vb200_pixshuf_cat_bwd(const void* dout, void* dprev, void* dskip, int B, int h, int w, int Cp, int Cs,
                          vb200_stream_t stream);
/* rows for the k2s2 downsample conv (timm ConvNeXtStage.downsample[1]): dst[(n,oh,ow),(kh,kw,c)] = src[n,2oh+kh,2ow+kw,c];
 * inverse != 0 scatters rows back (dgrad) */
int vb200_patchify2(const void* src, void* dst, int B, int H, int W, int C, int inverse, vb200_stream_t stream);
/* rows for UNeXt2Stem / StemDepthtoChannels (VM/components/stems.py:26-50,117-134): NCDHW input (x_dtype 0 bf16,
 * 1 fp16, 2 fp32) -> A[(n,oh,ow), ((c*D+z)*kH+kh)*kW+kw], row pitch Kpad */
int vb200_stem_patchify(const void* x, int x_dtype, void* A, int B, int Cin, int D, int H, int W, int kH, int kW,
                        int Kpad, int dtype, vb200_stream_t stream);
/* generic NDHWC (C % 8 == 0) im2col / col2im, geom = {N,D,H,W,C, kd,kh,kw, sd,sh,sw, pd,ph,pw, OD,OH,OW};
 * K order is (kd,kh,kw,c).  Lowers nn.Conv3d / ConvTranspose3d of VM/components/heads.py:607-628,
 * VM/unet/blocks.py:88-113, VM/unet/unet3d_base.py:90-138, VM/components/conv_block_3d.py:261-274 onto vb200_gemm. */
int vb200_im2col3d(const void* u, void* col, const int32_t* geom, vb200_stream_t stream);
int vb200_col2im3d(const void* dcol, void* du, const int32_t* geom, int dtype, vb200_stream_t stream);
/* conv_dw.weight [C,1,7,7] -> tap-major wt [49][C] and flipped wtf [49][C] (fp32) */
int vb200_dw_pack(const float* w, float* wt, float* wtf, int C, vb200_stream_t stream);
/* all weight packs of a step in one launch.  table (device, int64[n_items][7]) = {src, dst, dst2, R, Cc, kind, first_block}:
 * kind 0 cast [R,Cc]; 1 cast + transpose; 2 depthwise taps (src [C=R][49] -> dst, dst2 fp32 [49][C], dst2 flipped);
 * each block converts 1024 elements */
int vb200_pack_multi(const void* table, int n_items, int64_t total_blocks, int dtype, vb200_stream_t stream);
/* fp32 [R,Cc] -> 16-bit (weight packing); transpose != 0 writes [Cc,R] */
int vb200_cast_pack(const float* src, void* dst, int64_t R, int64_t Cc, int transpose, int dtype, vb200_stream_t stream);
/* nn.Conv3d weight [Co, Ci, KD, KH, KW] fp32 -> 16-bit GEMM rows (VM/unet/unet3d.py, unet25d.py, conv_block_3d.py convs):
 * flipped == 0: out [cout_pad][(kd, kh, kw, cin_pad)] (forward / weight-gradient order);
 * flipped != 0: out [cin_pad][(kd, kh, kw, cout_pad)] of the spatially flipped filter (the data gradient's filter).
 * Padding rows / channels are zero-filled.  Up to 27 taps. */
int vb200_conv_weight_rows(const float* w, void* out, int Co, int Ci, int KD, int KH, int KW, int cin_pad, int cout_pad,
                           int flipped, int dtype, vb200_stream_t stream);

/* ---- 3-D U-Net family: UNet3DBase / Unet3d (VM/unet/unet3d_base.py:145-198, VM/unet/blocks.py:88-113) ---- */
/* model boundary: x (N,C,S) [x_dtype 0 bf16 | 1 fp16 | 2 fp32] -> y (N,S,Cpad) 16-bit channels-last, zero channel padding */
int vb200_to_channels_last(const void* x, int x_dtype, void* y, int64_t N, int C, int Cpad, int64_t S, int dtype,
                           vb200_stream_t stream);
int vb200_from_channels_last(const void* y, void* x, int64_t N, int C, int Cpad, int64_t S, vb200_stream_t stream);
/* y = act(x * scale[c] + shift[c]) on rows [M,C]: nn.BatchNorm3d (folded statistics) + nn.ReLU of Block.forward */
int vb200_affine_act(const void* x, const float* scale, const float* shift, void* y, int64_t M, int C, int relu, int dtype,
                     vb200_stream_t stream);
/* BatchNorm backward: s1[c] += sum dy', s2[c] += sum dy' * xhat (dy' = dy masked by y > 0 when relu); s1, s2 pre-zeroed */
int vb200_bn_bwd_reduce(const void* dy, const void* x, const void* y, const float* mean, const float* rstd, float* s1,
                        float* s2, int64_t M, int C, int relu, int dtype, vb200_stream_t stream);
/* dx = g[c] * (dy' - m1[c] - xhat * m2[c]) */
int vb200_bn_bwd_apply(const void* dy, const void* x, const void* y, const float* mean, const float* rstd, const float* g,
                       const float* m1, const float* m2, void* dx, int64_t M, int C, int relu, int dtype,
                       vb200_stream_t stream);
/* the same with the raw operands: g = gamma * rstd, m1 = s1 * inv_m, m2 = s2 * inv_m formed in the kernel (s1 / s2 = the
 * column sums of vb200_bn_bwd_reduce, inv_m = 1 / rows in training mode, 0 in eval mode) */
int vb200_bn_bwd_apply_raw(const void* dy, const void* x, const void* y, const float* mean, const float* rstd, const float* gamma,
                           const float* s1, const float* s2, float inv_m, void* dx, int64_t M, int C, int relu, int dtype,
                           vb200_stream_t stream);
/* BatchNorm3d statistics -> apply-pass operands (nn.BatchNorm3d of VM/unet/unet3d.py / conv_block_3d.py, training and eval):
 * sums [2][Cc] = column sums of (x - pivot) and (x - pivot)^2 over M rows (null: eval mode, running statistics);
 * scale = weight * rstd, shift = bias - mean * scale, mean, rstd (fp32 [Cc]; channels >= Cn are padding: scale = shift = 0);
 * momentum >= 0 updates run_mean / run_var in place (unbiased variance), as torch does in training mode */
int vb200_bn_finalize(const float* sums, const float* pivot, const float* weight, const float* bias, float* run_mean,
                      float* run_var, int Cn, int Cc, double M, float eps, float momentum, float* scale, float* shift,
                      float* mean, float* rstd, vb200_stream_t stream);
/* torch.cat([a, b], 1) on channels-last rows (inverse != 0: split out back into a and b) */
int vb200_cat2(void* a, void* b, void* out, int64_t M, int Ca, int Cb, int inverse, vb200_stream_t stream);
/* y = x (+ other) (+ bias[c]) on rows [M,C] */
int vb200_add_rows(const void* x, const void* other, const float* bias, void* y, int64_t M, int C, int dtype,
                   vb200_stream_t stream);

/* ---- 2.5-D U-Net: Unet25d / ConvBlock3D (VM/unet/unet25d.py:206-251, VM/components/conv_block_3d.py:261-298) on
 * channels-last rows [N,D,H,W,C], C % 8 == 0.  The block convolutions are vb200_conv3d_igemm* / vb200_gemm. ---- */
/* nn.Dropout3d (scale [N,C] fp32 = keep mask / (1-p), NULL = none) fused with nn.ReLU (conv_block_3d.py:268-272):
 * forward (gate == NULL): y = relu?(x * scale[n,c]);  backward (gate = forward output): y = x * scale[n,c] * (gate > 0) */
int vb200_scale_relu(const void* x, const void* gate, const float* scale, void* y, int64_t N, int64_t rows_per_sample,
                     int C, int relu, int dtype, vb200_stream_t stream);
/* nn.AvgPool3d((1,2,2), stride (1,2,2)) (unet25d.py:104-112) on P = N*D planes [P,H,W,C] -> [P,H/2,W/2,C];
 * backward != 0: src = dy [P,H/2,W/2,C], dst = dx [P,H,W,C] */
int vb200_avgpool_hw2(const void* src, void* dst, int64_t P, int H, int W, int C, int backward, int dtype,
                      vb200_stream_t stream);
/* nn.Upsample(scale_factor=(1,2,2), mode="trilinear", align_corners=False) (unet25d.py:114-116): [P,H,W,C] -> [P,2H,2W,C];
 * backward != 0: src = dy [P,2H,2W,C], dst = dx [P,H,W,C] (the adjoint) */
int vb200_upsample2x_hw(const void* src, void* dst, int64_t P, int H, int W, int C, int backward, int dtype,
                        vb200_stream_t stream);

/* ---- ContrastiveEncoder pooled head + projection MLP (VM/contrastive/encoder.py:114-124,138-154) ---- */
/* nn.BatchNorm1d over rows of x [B,C] 16-bit (+ optional fused ReLU); training: batch statistics, var_unbiased (for the
 * running-var update) written when non-NULL; eval: run_mean / run_var are used */
int vb200_bn_rows_fwd(const void* x, const float* gamma, const float* beta, const float* run_mean, const float* run_var,
                      void* y, float* mean, float* rstd, float* var_unbiased, int B, int C, float eps, int training,
                      int relu, int dtype, vb200_stream_t stream);
int vb200_bn_rows_bwd(const void* dy, const void* x, const void* y, const float* gamma, const float* mean,
                      const float* rstd, void* dx, float* dgamma, float* dbeta, int B, int C, int training, int relu,
                      int dtype, vb200_stream_t stream);
/* out[b,r,c] = src[b,c] * scale: gradient of timm's global average pool (SelectAdaptivePool2d('avg')) */
int vb200_bcast_rows(const void* src, void* out, int B, int R, int C, float scale, int dtype, vb200_stream_t stream);

/* ---- implicit-GEMM Conv3d k=3, stride 1, small channel counts (tcgen05, im2col folded into 5-D TMA boxes) ----
 * Replaces nn.Conv3d(k=3) of monai Convolution in PixelToVoxelHead (VM/components/heads.py:607-628) and its cuDNN
 * dgrad / wgrad.  geom = {N, D, H, W, pd, ph, pw}; u [N,D,H,W,cin] 16-bit channels-last (cin in 8|16|32);
 * out [N,OD,OH,OW,co_store], OD = D + 2*pd - 2 etc.
 * wpack: 16-bit [KCH][cout_pad][8], K order (kd, kh, channel chunk, kw), KCH = 27*cin/8 padded to even with a zero
 * chunk; cout_pad in {16,32} is the MMA N; co_store <= cout_pad channels are written per voxel.
 * The data gradient is the same call with flipped / transposed weights and padding 2 - p. */
int vb200_conv3d_k3(const void* u, const void* wpack, const float* bias, void* out, const int32_t* geom, int cin,
                    int cout_pad, int co_store, int dtype, vb200_stream_t stream);
/* dw[cout][9 (kd,kh)][3 (kw)][8 (ci)] fp32 (pre-zeroed, accumulated with atomics); cin == 8; dz [N,OD,OH,OW,cout] */
int vb200_conv3d_k3_wgrad(const void* u, const void* dz, float* dw, const int32_t* geom, int cin, int cout, int dtype,
                          vb200_stream_t stream);

/* ---- FCMAE (VM/unet/fcmae.py) ---- */
/* Sparse (masked) path of MaskedConvNeXtV2Block on channels-last rows (masked_patchify / masked_unpatchify / `x *= unmasked`,
 * fcmae.py:91-141, 215-227): dst[r,:] = (map[r] >= 0 ? src[map[r],:] : 0) + (base ? base[r,:] : 0), r < n_dst, C % 8 == 0.
 * gather: map = row indices of the unmasked pixels; scatter: map = inverse index (-1 at masked pixels), base = shortcut. */
int vb200_rows_select(const void* src, const int32_t* map, const void* base, void* dst, int64_t n_dst, int C, int dtype,
                      vb200_stream_t stream);
/* PixelToVoxelShuffleHead (VM/components/heads.py:657-695) = monai SubpixelUpsample(scale r, pre_conv=None, pad_pool):
 * forward (backward == 0): src = decoder rows [B,h,w,Cq*r*r] -> dst [B,Cq,h*r,w*r] (NCHW; == NCDHW after the head's reshape)
 *   dst = AvgPool2d(r, stride 1)(ConstantPad2d((r-1,0,r-1,0))(pixel_shuffle(src, r)))   (pool == 0: the shuffle alone)
 * backward (!= 0): src = d dst [B,Cq,h*r,w*r], dst = d decoder rows [B,h,w,Cq*r*r] */
int vb200_shuffle_pool(const void* src, void* dst, int B, int h, int w, int Cq, int r, int pool, int backward, int dtype,
                       vb200_stream_t stream);

/* ---- MixedLoss / ms_ssim_25d (VU/losses/mixed_loss.py:42-69, VU/evaluation/metrics.py:174-349) ----
 * One pyramid level of ms_ssim_25d in one pass over preds x and target y, both [B,C,D,H,W] contiguous (NCDHW), each of
 * dtype VB200_BF16 | VB200_FP16 | VB200_FP32.  Rounding points follow _compute_ssim_and_cs_bf16 (metrics.py:174-262): the
 * five window inputs (x, y, x*x, y*y, x*y; products formed in fp32) and the five window means are rounded to bf16, the window
 * weight is bf16(1 / (D*kh*kw)), everything else is fp32.  flags: 1 = SSIM (window (D,kh,kw), valid), 2 = L1 / L2 sums,
 * 4 = avg_pool3d((1,2,2)) of both volumes into x_pool / y_pool [B,C,D,H/2,W/2] (their own dtypes).
 *   acc [B][4] += per-sample SUMS: ssim map, contrast-sensitivity map, |x-y|, (x-y)^2   (caller zeroes, divides by counts)
 *   mu  [5][B*C][H-kh+1][W-kw+1] fp32 = the window means (for the backward pass), or NULL
 *   data_range = device scalar target.max() of this level (metrics.py:296); data_range_next (pre-set to -inf, or NULL)
 *   receives max(y_pool) */
int vb200_ssim25d_level_fwd(const void* x, const void* y, int x_dtype, int y_dtype, int B, int C, int D, int H, int W, int kh,
                            int kw, const float* data_range, float* acc, float* mu, void* x_pool, void* y_pool,
                            float* data_range_next, int flags, vb200_stream_t stream);
/* d loss / d x of one level: g_ssim / g_cs [B] = upstream gradients of the per-sample MEANS of the two maps, g_l1 / g_l2 =
 * device scalars, upstream of mean|x-y| and mean (x-y)^2, g_pool = gradient of x_pool (x's dtype); any of them may be NULL.
 * dx [B,C,D,H,W] in x's dtype is overwritten. */
int vb200_ssim25d_level_bwd(const void* x, const void* y, int x_dtype, int y_dtype, int B, int C, int D, int H, int W, int kh,
                            int kw, const float* data_range, const float* mu, const float* g_ssim, const float* g_cs,
                            const float* g_l1, const float* g_l2, const void* g_pool, void* dx, vb200_stream_t stream);
/* max over n elements (target.max(), metrics.py:296) into *out, which the caller pre-sets to -inf */
int vb200_max_f(const void* x, int dtype, int64_t n, float* out, vb200_stream_t stream);

/* ---- Prediction path (CY/engine.py:61-71 _center_crop_to_shape, :711-805 predict_sliding_windows;
 * VU/callbacks/prediction_writer.py:74-111 _blend_in) ----
 * dst [B*C, Z, H, W] <- window prediction src [B*C, d, Hs, Ws] centre-cropped at (oy, ox), blended into Z range [z0, z0+d):
 * z0 == 0: dst = src; else samples = min(z0+1, d), f_i = min(d-i, samples), dst = dst*(f-1)/f + src/f.  dtypes: VB200_BF16 |
 * VB200_FP16 | VB200_FP32, independent for dst and src; arithmetic in fp32. */
int vb200_blend_window(void* dst, const void* src, int dst_dtype, int src_dtype, int64_t BC, int Z, int H, int W, int d,
                       int Hs, int Ws, int z0, int oy, int ox, vb200_stream_t stream);

/* ---- Optimizer step of the training path (VU/optimizers.py:10-61 configure_adamw_scheduler; CY/engine.py:547-554 and
 * the contrastive engine's configure_optimizers: torch.optim.AdamW over model.parameters()) ----
 * One AdamW step (decoupled weight decay, bias-corrected moments, fp32) over n_tensors parameter tensors:
 *   p -= lr*wd*p;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps),  t = *step + 1
 * table: DEVICE [n_tensors][4] pointers {p, m, v, step}: step = fp32 count of this tensor's completed steps, advanced by the
 * call; grads: HOST array of n_tensors device pointers (fp32, dense);
 * chunk_start: HOST [n_tensors+1] prefix sums of chunks per tensor; chunks: DEVICE [chunk_start[n_tensors]][4] int32
 * {tensor index within its window of 448 tensors, element offset, count <= 2048, 0}.  done: DEVICE uint32 scratch (zero).
 * lr_ptr / grad_scale / found_inf: optional DEVICE fp32
 * scalars (scheduler-driven lr; GradScaler: gradients are divided by *grad_scale and written back, *found_inf == 1 skips the
 * step). */
int vb200_adamw_step(const void* table, void* const* grads, const int32_t* chunk_start, int n_tensors, const void* chunks,
                     uint32_t* done, float lr, const float* lr_ptr, float beta1, float beta2, float eps,
                     float weight_decay, int maximize, const float* grad_scale, const float* found_inf, vb200_stream_t stream);

/* ---- The step's default loss (CY/engine.py:197 nn.MSELoss on the autocast prediction and the fp32 target) ----
 * dtypes: VB200_BF16 | VB200_FP16 | VB200_FP32, independent for pred and target; arithmetic in fp32; n elements, bases 32-byte
 * aligned.  vb200_mse_sum: *sum += scale * sum_i (pred_i - target_i)^2 (sum pre-zeroed; scale = 1/n for the mean).
 * vb200_mse_bwd: dpred_i = (pred_i - target_i) * (*gout) * scale in pred's dtype (scale = 2/n for the mean). */
int vb200_mse_sum(const void* pred, const void* target, int pred_dtype, int target_dtype, int64_t n, float scale, float* sum,
                  vb200_stream_t stream);
int vb200_mse_bwd(const void* pred, const void* target, int pred_dtype, int target_dtype, int64_t n, const float* gout,
                  float scale, void* dpred, vb200_stream_t stream);

/* ---- PixelToVoxelHead (VM/components/heads.py:594-641) ---- */
/* forward (backward == 0): dec [B,h,w,4*Cm] -> u [B,Dz,2h,2w,Cu] = unfold(pool(pixelshuffle2(dec)));
 * backward (!= 0): src = du, dst = ddec */
int vb200_head_shuffle_pool(const void* src, void* dst, int B, int h, int w, int Cm, int Dz, int Cu, int pool,
                            int backward, int dtype, vb200_stream_t stream);
/* InstanceNorm3d statistics of z [B,R,C]: mean, rstd [B,C] (sum, sumsq are scratch [B,C]) */
int vb200_instnorm_stats(const void* z, float* sum, float* sumsq, float* mean, float* rstd, int B, int64_t R, int C,
                         float eps, int dtype, vb200_stream_t stream);
/* out (B,Co,Dz,2H,2W) NCDHW 16-bit = PixelShuffle2(Conv3d_k1(PReLU(InstanceNorm(z)))) */
int vb200_head_tail_fwd(const void* z, const float* mean, const float* rstd, const float* alpha, int alpha_n,
                        const float* W1, const float* b1, void* out, int B, int Dz, int H, int W, int Cmid, int Co4,
                        int dtype, vb200_stream_t stream);
/* phase 0: reductions sdp, sdpx [B,Cmid], db1 [Co4], dalpha (pre-zeroed) + materialised act [B,R,Cmid] and
 * dt [B,R,ldt] (ldt = Co4 rounded up to 8) for the dW1 = dt^T act GEMM;  phase 1: dz [B,R,Cmid], dbz [Cmid] */
int vb200_head_tail_bwd(int phase, const void* z, const float* mean, const float* rstd, const float* alpha,
                        int alpha_n, const float* W1, const void* dout, float* sdp, float* sdpx, float* db1,
                        float* dalpha, void* act_out, void* dt_out, void* dz, float* dbz, int B, int Dz, int H, int W,
                        int Cmid, int Co4, int dtype, vb200_stream_t stream);
/* Streaming form of vb200_head_tail_bwd for the BASELINE head geometry (Cmid == 32, Co4 == 8, 256 % W == 0,
 * Dz*H*W % 256 == 0; anything else returns VB200_ERR_UNSUPPORTED): z tiles and the pixel-shuffled dout lines stream through
 * a bulk-copy / mbarrier ring, nothing is materialised.  phase 0 accumulates sdp, sdpx [B,32], db1 [8], dalpha, dW1 [8,32]
 * (all pre-zeroed, fp32); phase 1 writes dz [B,R,32] and accumulates dbz [32] (pre-zeroed) from the finished sums. */
int vb200_head_tail_bwd_stream(int phase, const void* z, const float* mean, const float* rstd, const float* alpha,
                               int alpha_n, const float* W1, const void* dout, float* sdp, float* sdpx, float* db1,
                               float* dalpha, float* dW1, void* dz, float* dbz, int B, int Dz, int H, int W, int Cmid, int Co4,
                               int dtype, vb200_stream_t stream);

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
