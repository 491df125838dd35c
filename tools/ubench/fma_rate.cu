// FP32 FMA issue-rate microbenchmark (sm_100a): scalar FFMA vs packed FFMA2, register operands, with and without a
// shared (reusable) multiplicand.  Prints lane-MACs per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, int iters, float a0, float b0) {
  float2 acc[16];
  float2 x[8], w[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f, i);
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = make_float2(a0 + i, a0 - i); w[i] = make_float2(b0 + i, b0 * i); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) {  // scalar FFMA, distinct operands
          acc[i].x = fmaf(x[(i + r) & 7].x, w[r].x, acc[i].x);
          acc[i].y = fmaf(x[(i + r) & 7].y, w[r].y, acc[i].y);
        } else if (MODE == 1) {  // FFMA2, one operand shared by 16 consecutive instructions (reuse)
          acc[i] = __ffma2_rn(x[(i + r) & 7], w[r], acc[i]);
        } else {  // FFMA2, both operands vary
          acc[i] = __ffma2_rn(x[(i + r) & 7], w[(i * 3 + r) & 7], acc[i]);
        }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int blocks_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * blocks_per_sm * 128 * 4);
  const int iters = 4000;
  k<MODE><<<sms * blocks_per_sm, 128>>>(out, 10, 1.f, 1.f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms * blocks_per_sm, 128>>>(out, iters, 1.0001f, 0.9999f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double macs = (double)sms * blocks_per_sm * 128 * iters * 8 * 16 * 2;
  printf("%-28s warps/SM %2d: %7.3f ms  %6.1f TMAC/s  = %5.1f lane-MAC/clk/SM at %d MHz nominal\n", name, blocks_per_sm * 4, ms,
         macs / ms / 1e9, macs / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
  cudaFree(out);
}

int main() {
  for (int b : {1, 2, 4}) {
    run<0>("FFMA scalar", b);
    run<1>("FFMA2 shared multiplicand", b);
    run<2>("FFMA2 distinct operands", b);
  }
  return 0;
}
