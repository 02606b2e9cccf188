// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition (1, 2 or 4 warps per scheduler).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mufu_kernel(float* out, long long* cyc, int iters) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 8);
  for (int warps : {4, 8, 16}) {
    const int iters = 2048;
    for (int rep = 0; rep < 2; ++rep) { mufu_kernel<<<148, warps * 32>>>(o, c, iters); cudaDeviceSynchronize(); }
    long long cyc; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
    const double per_smsp_inst = double(iters) * 16 * (warps / 4.0);
    printf("%2d warps/SM: %.2f cycles per warp-wide EX2 per scheduler (%.1f lanes/clk/SM)\n", warps,
           cyc / per_smsp_inst, 32.0 * 4 / (cyc / per_smsp_inst));
  }
  return 0;
}
