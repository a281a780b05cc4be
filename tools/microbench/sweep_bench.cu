// Microbenchmark: the Rayleigh secular-function sweep (surfdisp_core.cuh, same code as the product kernel) alone,
// as a function of the warps resident per SM.  Tells how far the root-search kernel (4 warps per scheduler, the
// state machine's registers around the sweep) is from what the sweep itself can reach.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o sweep_bench sweep_bench.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../pysurfinv_b200/csrc/surfdisp_core.cuh"
using namespace sd;

// one group of 4 lanes = one layer stack (like the product kernel), 8 stacks per warp
template <int MINBLK>
__global__ void __launch_bounds__(128, MINBLK) sweep_kernel(const float4* __restrict__ stacks, int mm, int reps, float T, float* out) {
  extern __shared__ float4 rec[];
  const int groups = blockDim.x / 4;
  for (int i = threadIdx.x; i < groups * mm; i += blockDim.x) rec[i] = stacks[(size_t)((blockIdx.x * groups) % 4096) * mm + i];
  __syncthreads();
  const float4* myrec = rec + (threadIdx.x / 4) * mm;
  float acc = 0.f;
  float c0 = 3.0f + 0.01f * (threadIdx.x & 3);
  for (int r = 0; r < reps; ++r) {
    V2 e2, e3;
    const V2 d = rayleigh_adjoint2(v2(c0, c0 + 0.005f), T, mm, myrec, false, e2, e3);
    acc += vx(d) * 1e-30f + vy(e2) * 1e-30f;
    c0 += 1e-4f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MINBLK>
double run(const float4* stacks, int mm, float T, float* out, int ctas_per_sm, size_t pad_smem) {
  const int reps = 64, threads = 128;
  const size_t smem = (size_t)(threads / 4) * mm * sizeof(float4) + pad_smem;
  cudaFuncSetAttribute(sweep_kernel<MINBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_kernel<MINBLK>, threads, smem);
  const int blocks = 148 * occ * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    sweep_kernel<MINBLK><<<blocks, threads, smem>>>(stacks, mm, reps, T, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double steps = (double)blocks * threads * 2.0 * reps * (mm - 1);   // layer steps of single velocities
  printf("mm %3d T %5.1f  CTAs/SM %2d (asked %2d, %2d warps/SM)  %8.3f ms  %6.2f ps per layer-step  %6.1f TFLOP-eq/s\n", mm, T, occ, ctas_per_sm,
         occ * 4, best, best * 1e-3 / steps * 1e12, steps * 190.0 / (best * 1e-3) * 1e-12);
  return best;
}

int main() {
  const int MM = 64, NST = 4096;
  std::vector<float4> h((size_t)NST * MM);
  for (int s = 0; s < NST; ++s)
    for (int i = 0; i < MM; ++i) {
      const float z = (float)i / MM;
      const float b = 3.2f + 1.4f * z + 0.05f * ((s * 7 + i * 13) % 11) / 11.f;
      h[(size_t)s * MM + i] = make_float4(1.75f * b, b, 2.6f + 0.6f * z, (i == MM - 1) ? 0.f : 2.5f + 0.5f * z);
    }
  float4* stacks; cudaMalloc(&stacks, h.size() * sizeof(float4));
  cudaMemcpy(stacks, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, (size_t)148 * 64 * 128 * sizeof(float));
  // long period: every layer step in the thin tier; short period: thick tier mixed in
  for (float T : {40.f, 8.f}) {
    // 1..4 CTAs per SM (smem padding limits residency), then the register-limited maximum
    run<1>(stacks, MM, T, out, 1, 200 * 1024 - 32 * MM * 16);
    run<2>(stacks, MM, T, out, 2, 100 * 1024 - 32 * MM * 16);
    run<3>(stacks, MM, T, out, 3, 70 * 1024 - 32 * MM * 16);
    run<4>(stacks, MM, T, out, 4, 50 * 1024 - 32 * MM * 16);
    run<5>(stacks, MM, T, out, 5, 0);
    run<6>(stacks, MM, T, out, 6, 0);
    run<8>(stacks, MM, T, out, 8, 0);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
