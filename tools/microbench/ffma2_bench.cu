// Microbenchmark: packed FP32 (fma.rn.f32x2, sm_100) against scalar FFMA, alone and mixed with ALU work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  const float2 m = make_float2(0.999999f, 0.999998f), c = make_float2(1e-7f, 2e-7f);
  unsigned u0 = threadIdx.x, u1 = threadIdx.x * 3u, u2 = 7u, u3 = 11u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }   // 2 FFMA
        if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }                                       // 1 FFMA2
        if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, c); u0 = (u0 ^ u1) + u2; u1 = (u1 & u3) + u0; }          // FFMA2 + 4 ALU
        if (MODE == 3) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); u0 = (u0 ^ u1) + u2; u1 = (u1 & u3) + u0; }
        if (MODE == 4) { a[i] = __fmul2_rn(a[i], m); a[i] = __fadd2_rn(a[i], c); }
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u0 + u1);
}

template <int MODE>
void run(const char* name, float* buf, double flop_per_inner) {
  const int blocks = 148 * 8, threads = 256, iters = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(buf, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double inner = (double)blocks * threads * iters * 4 * 8;
  printf("%-28s %8.3f ms  %7.2f TFLOP/s  (%.3f inner-iter/clk/SM at 1965 MHz)\n", name, best, inner * flop_per_inner / (best * 1e-3) * 1e-12,
         inner / (best * 1e-3) / 148 / 1.965e9);
}

int main() {
  float* buf; cudaMalloc(&buf, 148 * 8 * 256 * sizeof(float));
  run<0>("2x FFMA", buf, 4);
  run<1>("1x FFMA2", buf, 4);
  run<2>("FFMA2 + 4 ALU", buf, 4);
  run<3>("2x FFMA + 4 ALU", buf, 4);
  run<4>("FMUL2 + FADD2", buf, 4);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
