// Plain CUDA-core versions of the GEMM and attention entry points.  They exist so that tests
// (and bring-up) can check the tcgen05 kernels against an independent GPU implementation of the
// same contract; the model path never calls them.
#include "common.cuh"

namespace vb200 {

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

template <typename OutT>
__device__ __forceinline__ void store_out(OutT* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ void store_out<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// 32x32 output tile per block of 32x8 threads, K walked in chunks of 32 through shared memory.
template <typename OutT>
__global__ void __launch_bounds__(256) gemm_simt_kernel(
    OutT* __restrict__ out, const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W,
    const float* __restrict__ bias, const float* residual, int M, int N, int K, int epi, int a_f16) {
  __shared__ float sa[32][33], sw[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const int k = k0 + tx;
      sa[i][tx] = !(m0 + i < M && k < K) ? 0.f
                  : a_f16 ? __half2float(reinterpret_cast<const __half*>(A)[static_cast<size_t>(m0 + i) * K + k])
                          : __bfloat162float(A[static_cast<size_t>(m0 + i) * K + k]);
      sw[i][tx] = !(n0 + i < N && k < K) ? 0.f
                  : a_f16 ? __half2float(reinterpret_cast<const __half*>(W)[static_cast<size_t>(n0 + i) * K + k])
                          : __bfloat162float(W[static_cast<size_t>(n0 + i) * K + k]);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[r] += sa[ty + 8 * r][k] * sw[tx][k];
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n >= N) return;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + ty + 8 * r;
    if (m >= M) continue;
    float v = acc[r];
    if (epi != VB200_EPI_NONE) v += bias[n];
    if (epi == VB200_EPI_BIAS_GELU) v = gelu_erf(v);
    if (epi == VB200_EPI_BIAS_RESIDUAL) v += residual[static_cast<size_t>(m) * N + n];
    store_out<OutT>(out + static_cast<size_t>(m) * N + n, v);
  }
}

// One warp per (query row, head): online softmax over the utterance's keys, head_dim 64.
__global__ void __launch_bounds__(256) attn_simt_kernel(
    __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ qkv,
    const int32_t* __restrict__ cu_rows, int B, int M, int n_heads, float scale) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int d = n_heads * 64;
  const long total = static_cast<long>(M) * n_heads;
  for (long w = static_cast<long>(blockIdx.x) * wpb + (threadIdx.x >> 5); w < total;
       w += static_cast<long>(gridDim.x) * wpb) {
    const int row = static_cast<int>(w / n_heads), h = static_cast<int>(w % n_heads);
    int b = 0;
    while (b + 1 < B && cu_rows[b + 1] <= row) ++b;   // B is small; linear scan
    const int k0 = cu_rows[b], k1 = cu_rows[b + 1];
    const __nv_bfloat16* q = qkv + static_cast<size_t>(row) * 3 * d + h * 64;
    const float q0 = __bfloat162float(q[lane]), q1 = __bfloat162float(q[lane + 32]);
    float mx = -INFINITY, sum = 0.f, o0 = 0.f, o1 = 0.f;
    for (int j = k0; j < k1; ++j) {
      const __nv_bfloat16* kk = qkv + static_cast<size_t>(j) * 3 * d + d + h * 64;
      const __nv_bfloat16* vv = kk + d;
      float s = q0 * __bfloat162float(kk[lane]) + q1 * __bfloat162float(kk[lane + 32]);
      s = warp_sum(s) * scale;
      const float nm = fmaxf(mx, s);
      const float corr = __expf(mx - nm), p = __expf(s - nm);
      sum = sum * corr + p;
      o0 = o0 * corr + p * __bfloat162float(vv[lane]);
      o1 = o1 * corr + p * __bfloat162float(vv[lane + 32]);
      mx = nm;
    }
    __nv_bfloat16* o = out + static_cast<size_t>(row) * d + h * 64;
    o[lane] = __float2bfloat16_rn(o0 / sum);
    o[lane + 32] = __float2bfloat16_rn(o1 / sum);
  }
}

}  // namespace vb200

using namespace vb200;

extern "C" int vb200_gemm_bf16_simt(void* out, vb200_dtype out_dtype, const void* A, vb200_dtype a_dtype, const void* W,
                                    const float* bias, const float* residual, int32_t M, int32_t N,
                                    int32_t K, vb200_epilogue epi, vb200_stream_t stream) {
  if (M <= 0 || N <= 0) return VB200_OK;
  VB_REQUIRE(out && A && W, "gemm_simt: null pointer");
  VB_REQUIRE(epi == VB200_EPI_NONE || bias, "gemm_simt: epilogue %d needs bias", static_cast<int>(epi));
  VB_REQUIRE(epi != VB200_EPI_BIAS_RESIDUAL || (residual && out_dtype == VB200_F32),
             "gemm_simt: BIAS_RESIDUAL needs residual and fp32 output");
  if (M <= 0 || N <= 0) return VB200_OK;
  dim3 grid((N + 31) / 32, (M + 31) / 32), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(A);
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(W);
  if (out_dtype == VB200_F32)
    gemm_simt_kernel<float><<<grid, block, 0, st>>>(static_cast<float*>(out), a, w, bias, residual, M, N, K, epi, a_dtype == VB200_F16);
  else if (out_dtype == VB200_BF16)
    gemm_simt_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<__nv_bfloat16*>(out), a, w, bias, residual, M, N, K, epi, a_dtype == VB200_F16);
  else
    gemm_simt_kernel<__half><<<grid, block, 0, st>>>(static_cast<__half*>(out), a, w, bias, residual, M, N, K, epi, a_dtype == VB200_F16);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}

extern "C" int vb200_attn_varlen_simt(void* out_bf16, const void* qkv_bf16, const int32_t* cu_rows,
                                      int32_t B, int32_t max_T, int32_t M, int32_t n_heads,
                                      float scale, vb200_stream_t stream) {
  (void)max_T;
  if (M <= 0) return VB200_OK;
  VB_REQUIRE(out_bf16 && qkv_bf16 && cu_rows, "attn_simt: null pointer");
  VB_REQUIRE(B >= 1 && n_heads >= 1, "attn_simt: bad sizes");
  if (M <= 0) return VB200_OK;
  long total = static_cast<long>(M) * n_heads;
  long grid = (total + 7) / 8;
  const long cap = static_cast<long>(num_sms()) * 32;
  if (grid > cap) grid = cap;
  attn_simt_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(out_bf16), static_cast<const __nv_bfloat16*>(qkv_bf16), cu_rows,
      B, M, n_heads, scale);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}
