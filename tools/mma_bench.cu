// Micro-benchmark: sustained cycles per tcgen05.mma for the shapes the attention kernel uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../tts-with-diffusion-model_b200/csrc/common.cuh"
using namespace vb200;
namespace vb200 { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -2; } }

// mode 0: SS (A smem K-major, B smem K-major); mode 1: TS (A from TMEM, B smem MN-major, like P V)
// mode 2: TS N=64, consecutive MMAs alternate between two accumulators (is the 56-cycle floor a
//         dependent-accumulate latency?)
// mode 3: one attention block's worth per iteration, 4 SS N=128 then 8 TS N=64 (as the kernel issues)
// mode 4: the same 12 MMAs interleaved SS, TS, TS, SS, TS, TS, ...
// tmem_cols = 256 lets two CTAs share an SM (two issuing warps on one tensor pipe)
__global__ void __launch_bounds__(128, 2) mma_bench_kernel(int M, int N, int mode, int iters, long long* out,
                                                           int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, tmem_cols); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {                 // warp-uniform issue loop: one elected lane executes the MMAs
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(M, N, false, mode == 1);
    const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem));
    const uint64_t db = mode == 1 ? umma_desc_mnmajor_sw128(smem_u32(smem + 16384), 1024)
                                  : umma_desc_kmajor_sw128(smem_u32(smem + 16384));
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    const uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
    const uint64_t dbv = umma_desc_mnmajor_sw128(smem_u32(smem + 32768), 1024);
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode <= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (leader) {
            if (mode == 1) umma_ts(tb + 128, tb + k * 8, db + k * 128, idesc, k != 0);
            else umma_ss(tb, da + 2 * k, db + 2 * k, idesc, k != 0);
          }
        }
      } else if (mode == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (leader) umma_ts(tb + 128 + (k & 1) * 64, tb + k * 8, dbv + k * 128, idesc_o, k > 1);
      } else if (mode == 3) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (leader) umma_ss(tb, da + 2 * k, db + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (leader) umma_ts(tb + 192, tb + 128 + k * 8, dbv + k * 128, idesc_o, k != 0);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (leader) {
            umma_ss(tb, da + 2 * k, db + 2 * k, idesc_s, k != 0);
            umma_ts(tb + 192, tb + 128 + 2 * k * 8, dbv + 2 * k * 128, idesc_o, k != 0);
            umma_ts(tb + 192, tb + 128 + (2 * k + 1) * 8, dbv + (2 * k + 1) * 128, idesc_o, true);
          }
        }
      }
      if ((it & 15) == 15) { if (leader) umma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; }
    }
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, phase);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && leader) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, tmem_cols); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct { int M, N, mode, ctas; const char* name; } cfg[] = {
    {128, 256, 0, 1, "SS M128 N256"}, {128, 128, 0, 1, "SS M128 N128"}, {128, 64, 0, 1, "SS M128 N64"},
    {128, 64, 1, 1, "TS M128 N64 (B MN-major)"}, {128, 128, 1, 1, "TS M128 N128 (B MN-major)"},
    {128, 16, 0, 1, "SS M128 N16"},
    {128, 64, 2, 1, "TS N64, two accumulators"},
    {128, 128, 0, 2, "SS M128 N128, 2 CTAs/SM"}, {128, 64, 1, 2, "TS M128 N64, 2 CTAs/SM"},
    {128, 64, 3, 1, "block: 4 SS + 8 TS /12"}, {128, 64, 4, 1, "block interleaved /12"},
    {128, 64, 3, 2, "block, 2 CTAs/SM /12"}, {128, 64, 4, 2, "block interleaved, 2 CTAs/SM /12"}};
  for (auto& c : cfg) {
    const int iters = 4096;
    for (int rep = 0; rep < 2; ++rep) {
      mma_bench_kernel<<<148 * c.ctas, 128, 64 * 1024>>>(c.M, c.N, c.mode, iters, d, c.ctas == 2 ? 256 : 512);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double per = c.mode >= 3 ? 12.0 : 4.0;
    printf("%-34s %8.1f cycles / MMA per issuing warp (nominal %d)\n", c.name, double(cyc) / (iters * per), c.M * c.N / 256);
  }
  return 0;
}
