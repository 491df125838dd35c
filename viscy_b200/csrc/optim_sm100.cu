// AdamW step of the training path (VU/optimizers.py:10 configure_adamw_scheduler, CY/engine.py:547-554 and the contrastive
// engine's AdamW(self.parameters(), lr): torch.optim.AdamW semantics, decoupled weight decay) as ONE launch over every
// parameter tensor of a group.  HBM-bound: 16 B read + 12 B written per element.  A chunk table maps grid-stride work
// items to (tensor, offset, count), so the hundreds of small tensors of a ConvNeXt (biases, LayerNorm / GRN vectors)
// ride in the same wave as the large weight matrices instead of taking a launch slot each.
#include "common.cuh"

namespace vb {

struct AdamTensor {  // device-resident, static: parameter, its two moments and its own step count (torch keeps one per
  float* p;          // parameter: a tensor that joins later - no gradient in its first steps - starts its bias correction
  float* m;          // from step 1)
  float* v;
  float* step;
};
constexpr int ADAM_MAX_TENSORS = 448;  // gradient pointers travel by value (they move from step to step in eager mode)
struct AdamGrads {
  float* g[ADAM_MAX_TENSORS];
};
struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay;
  int maximize;
  const float* lr_ptr;      // optional device scalar (capturable schedulers)
  const float* grad_scale;  // optional: GradScaler's scale, gradients are divided by it (and written back unscaled)
  const float* found_inf;   // optional: 1.0 -> the whole step is skipped
};

constexpr int ADAM_CHUNK = 2048;  // elements per work item: 256 threads x 2 x float4

__device__ __forceinline__ void adam_math(float& p, float& g, float& m, float& v, const AdamHyper& h, float lr, float inv_scale,
                                          float step_size, float bc2_sqrt) {
  g *= inv_scale;
  const float gs = g;  // stored back when a scale is given
  if (h.maximize) g = -g;
  if (h.weight_decay != 0.f) p -= lr * h.weight_decay * p;
  m = fmaf(h.beta1, m, fmaf(-h.beta1, g, g));
  v = fmaf(h.beta2, v, fmaf(-h.beta2, g * g, g * g));
  const float denom = sqrtf(v) / bc2_sqrt + h.eps;
  p -= step_size * m / denom;
  g = gs;
}

// chunks[i] = {tensor (relative to this launch's first), offset, count, 0}.  Each tensor's step holds its number of completed
// steps; the last block of the launch advances them (every block has read what it needs by then).
__global__ void __launch_bounds__(256)
adamw_kernel(const AdamTensor* __restrict__ tab, int n_tensors, const __grid_constant__ AdamGrads grads,
             const int4* __restrict__ chunks, int n_chunks, const AdamHyper h, unsigned int* done) {
  const bool skip = h.found_inf != nullptr && *h.found_inf == 1.f;
  if (!skip) {
    const float lr = h.lr_ptr != nullptr ? *h.lr_ptr : h.lr;
    const float inv_scale = h.grad_scale != nullptr ? 1.f / *h.grad_scale : 1.f;
    const bool store_g = h.grad_scale != nullptr;
    for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
      const int4 c = chunks[ci];
      const AdamTensor t = tab[c.x];
      const float s = *t.step + 1.f;
      const float step_size = lr / (1.f - powf(h.beta1, s));
      const float bc2_sqrt = sqrtf(1.f - powf(h.beta2, s));
      float* __restrict__ pp = t.p + c.y;
      float* __restrict__ mp = t.m + c.y;
      float* __restrict__ vp = t.v + c.y;
      float* __restrict__ gp = grads.g[c.x] + c.y;
      const bool vec = ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(mp) | reinterpret_cast<uintptr_t>(vp) |
                         reinterpret_cast<uintptr_t>(gp)) & 15) == 0;
      if (vec) {
        const int n4 = c.z >> 2;
#pragma unroll
        for (int u = 0; u < ADAM_CHUNK / 1024; ++u) {
          const int i = u * 256 + threadIdx.x;
          if (i < n4) {
            float4 p4 = reinterpret_cast<float4*>(pp)[i], g4 = reinterpret_cast<float4*>(gp)[i];
            float4 m4 = reinterpret_cast<float4*>(mp)[i], v4 = reinterpret_cast<float4*>(vp)[i];
            adam_math(p4.x, g4.x, m4.x, v4.x, h, lr, inv_scale, step_size, bc2_sqrt);
            adam_math(p4.y, g4.y, m4.y, v4.y, h, lr, inv_scale, step_size, bc2_sqrt);
            adam_math(p4.z, g4.z, m4.z, v4.z, h, lr, inv_scale, step_size, bc2_sqrt);
            adam_math(p4.w, g4.w, m4.w, v4.w, h, lr, inv_scale, step_size, bc2_sqrt);
            reinterpret_cast<float4*>(pp)[i] = p4;
            reinterpret_cast<float4*>(mp)[i] = m4;
            reinterpret_cast<float4*>(vp)[i] = v4;
            if (store_g) reinterpret_cast<float4*>(gp)[i] = g4;
          }
        }
        for (int i = (n4 << 2) + threadIdx.x; i < c.z; i += 256) {  // tail of the tensor's last chunk
          float p1 = pp[i], g1 = gp[i], m1 = mp[i], v1 = vp[i];
          adam_math(p1, g1, m1, v1, h, lr, inv_scale, step_size, bc2_sqrt);
          pp[i] = p1; mp[i] = m1; vp[i] = v1;
          if (store_g) gp[i] = g1;
        }
      } else {  // a gradient view at an odd offset
        for (int i = threadIdx.x; i < c.z; i += 256) {
          float p1 = pp[i], g1 = gp[i], m1 = mp[i], v1 = vp[i];
          adam_math(p1, g1, m1, v1, h, lr, inv_scale, step_size, bc2_sqrt);
          pp[i] = p1; mp[i] = m1; vp[i] = v1;
          if (store_g) gp[i] = g1;
        }
      }
    }
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    if (!skip)
      for (int t = threadIdx.x; t < n_tensors; t += 256) *tab[t].step += 1.f;
    if (threadIdx.x == 0) *done = 0u;
  }
}

}  // namespace vb

using namespace vb;

// table: device [n] {p, m, v, step}; grads: HOST array of n device pointers; chunk_start: HOST [n + 1] prefix of chunks per
// tensor; chunks: device [chunk_start[n]] int4 {tensor % ADAM_MAX_TENSORS-window index, offset, count, 0} - see optim.py.
extern "C" int vb200_adamw_step(const void* table, void* const* grads, const int32_t* chunk_start, int n_tensors,
                                const void* chunks, uint32_t* done, float lr, const float* lr_ptr, float beta1,
                                float beta2, float eps, float weight_decay, int maximize, const float* grad_scale,
                                const float* found_inf, vb200_stream_t stream) {
  VB_REQUIRE(table && grads && chunk_start && chunks && done, "null pointer");
  VB_REQUIRE(n_tensors > 0, "no tensors");
  VB_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "betas (%g, %g) / eps %g", beta1, beta2, eps);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AdamHyper h{lr, beta1, beta2, eps, weight_decay, maximize, lr_ptr, grad_scale, found_inf};
  for (int t0 = 0; t0 < n_tensors; t0 += ADAM_MAX_TENSORS) {
    const int t1 = t0 + ADAM_MAX_TENSORS < n_tensors ? t0 + ADAM_MAX_TENSORS : n_tensors;
    AdamGrads g{};
    for (int t = t0; t < t1; ++t) {
      VB_REQUIRE(grads[t] != nullptr, "gradient %d is null", t);
      g.g[t - t0] = reinterpret_cast<float*>(grads[t]);
    }
    const int c0 = chunk_start[t0], c1 = chunk_start[t1];
    const int nc = c1 - c0;
    const int grid = nc < 148 * 8 ? (nc > 0 ? nc : 1) : 148 * 8;
    adamw_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const AdamTensor*>(table) + t0, t1 - t0, g,
                                       reinterpret_cast<const int4*>(chunks) + c0, nc, h, done);
    if (int rc = check_launch("vb200_adamw_step")) return rc;
  }
  return VB200_OK;
}
