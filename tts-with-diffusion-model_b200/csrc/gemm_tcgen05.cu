// out = epilogue(A[M,K] · W[N,K]^T): the linear layers of the denoiser (SURVEY.md §8a rows A1,
// F1, H1; reference base.py:110,129,209-214,355).
//
// Persistent, warp-specialised sm_100a kernel:
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, 4-stage ring)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma cta_group::1 kind::f16, 128x256x16, fp32 in TMEM)
//   warps 2..9    : epilogue       (tcgen05.ld -> bias / GELU / residual -> global)
// TMEM holds two 128x256 fp32 accumulators (512 columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Both operands are K-major (nn.Linear stores W as (N, K)), so no transposes.
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace vb200 {

namespace gemm {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES = BN * BK * 2;            // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 48 KB
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (2 + EPI_WARPS) * 32;   // 320
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 512;
}  // namespace gemm

__device__ __forceinline__ float gelu_erf_fast(float x) {
  // erf via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7), gelu(x) = 0.5 x (1 + erf(x / sqrt 2))
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);   // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

template <int EPI, typename OutT>
__global__ void __launch_bounds__(gemm::THREADS, 1) gemm_tcgen05_kernel(
    const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
    OutT* __restrict__ out, const float* __restrict__ bias, const float* residual, int M, int N,
    int K) {
  using namespace gemm;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;                      // 1024-byte aligned (SWIZZLE_128B atoms)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;   // [2]     MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;       // [2]     epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          tma_load_2d(sa, &tm_a, &full[stage], kb * BK, m_blk * BM);
          tma_load_2d(sa + A_BYTES, &tm_b, &full[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, false, false);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)   // +32 bytes (>>4 = 2) per UMMA_K inside the swizzle atom
            umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);          // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);            // accumulator complete -> epilogue
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue warps
    const int e = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = e >> 2;                   // which 128 of the 256 accumulator columns
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const int row = m_blk * BM + quad * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = half * 4; c < half * 4 + 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        const int n0 = n_blk * BN + c * 32;
        if (row < M && n0 < N) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          const bool full_chunk = (n0 + 32 <= N);
          if (EPI != VB200_EPI_NONE) {
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + i));
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
            } else {
              for (int i = 0; i < 32; ++i) if (n0 + i < N) v[i] += __ldg(bias + n0 + i);
            }
          }
          if (EPI == VB200_EPI_BIAS_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf_fast(v[i]);
          }
          const size_t off = static_cast<size_t>(row) * N + n0;
          if (EPI == VB200_EPI_BIAS_RESIDUAL) {
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(residual + off + i);
                v[i] += r4.x; v[i + 1] += r4.y; v[i + 2] += r4.z; v[i + 3] += r4.w;
              }
            } else {
              for (int i = 0; i < 32; ++i) if (n0 + i < N) v[i] += residual[off + i];
            }
          }
          if (sizeof(OutT) == 4) {
            float* o = reinterpret_cast<float*>(out) + off;
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
              for (int i = 0; i < 32; ++i) if (n0 + i < N) o[i] = v[i];
            }
          } else {
            uint16_t* o = reinterpret_cast<uint16_t*>(out) + off;
            constexpr bool is_bf16 = std::is_same<OutT, __nv_bfloat16>::value;
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 p;
                if (is_bf16) {
                  p.x = pack_bf16x2(v[i], v[i + 1]); p.y = pack_bf16x2(v[i + 2], v[i + 3]);
                  p.z = pack_bf16x2(v[i + 4], v[i + 5]); p.w = pack_bf16x2(v[i + 6], v[i + 7]);
                } else {
                  p.x = pack_f16x2(v[i], v[i + 1]); p.y = pack_f16x2(v[i + 2], v[i + 3]);
                  p.z = pack_f16x2(v[i + 4], v[i + 5]); p.w = pack_f16x2(v[i + 6], v[i + 7]);
                }
                *reinterpret_cast<uint4*>(o + i) = p;
              }
            } else {
              for (int i = 0; i < 32; ++i) {
                if (n0 + i < N) {
                  if (is_bf16) { __nv_bfloat16 h = __float2bfloat16_rn(v[i]); o[i] = *reinterpret_cast<uint16_t*>(&h); }
                  else { __half h = __float2half_rn(v[i]); o[i] = *reinterpret_cast<uint16_t*>(&h); }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
struct TmapKey {
  const void* ptr; uint64_t inner, outer, stride; uint32_t box_outer;
  bool operator<(const TmapKey& o) const {
    return std::tie(ptr, inner, outer, stride, box_outer) <
           std::tie(o.ptr, o.inner, o.outer, o.stride, o.box_outer);
  }
};

// Tensor maps are pure functions of (pointer, shape, box); cache them so steady-state launches
// (and CUDA-graph re-captures) do not re-encode.
int cached_tmap(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer,
                uint64_t stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::map<TmapKey, CUtensorMap> cache;
  const TmapKey key{ptr, inner, outer, stride_bytes, box_outer};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return VB200_OK; }
  }
  const int rc = make_tmap_2d_bf16(out, ptr, inner, outer, stride_bytes, box_inner, box_outer);
  if (rc != VB200_OK) return rc;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return VB200_OK;
}

template <int EPI, typename OutT>
static int launch_gemm(void* out, const CUtensorMap& ta, const CUtensorMap& tb, const float* bias,
                       const float* residual, int M, int N, int K, cudaStream_t st) {
  using namespace gemm;
  auto kern = gemm_tcgen05_kernel<EPI, OutT>;
  static bool configured = false;   // per template instantiation
  if (!configured) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, THREADS, SMEM_BYTES, st>>>(ta, tb, static_cast<OutT*>(out), bias, residual, M, N, K);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}

template <int EPI>
static int launch_gemm_dtype(void* out, vb200_dtype dt, const CUtensorMap& ta, const CUtensorMap& tb,
                             const float* bias, const float* residual, int M, int N, int K,
                             cudaStream_t st) {
  switch (dt) {
    case VB200_F32: return launch_gemm<EPI, float>(out, ta, tb, bias, residual, M, N, K, st);
    case VB200_BF16: return launch_gemm<EPI, __nv_bfloat16>(out, ta, tb, bias, residual, M, N, K, st);
    case VB200_F16: return launch_gemm<EPI, __half>(out, ta, tb, bias, residual, M, N, K, st);
  }
  set_error("gemm: unknown out dtype %d", static_cast<int>(dt));
  return VB200_ERR_INVALID;
}

}  // namespace vb200

using namespace vb200;

extern "C" int vb200_gemm_bf16(void* out, vb200_dtype out_dtype, const void* A, const void* W,
                               const float* bias, const float* residual, int32_t M, int32_t N,
                               int32_t K, vb200_epilogue epi, vb200_stream_t stream) {
  VB_REQUIRE(out && A && W, "gemm: null pointer");
  VB_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad sizes M=%d N=%d K=%d", M, N, K);
  VB_REQUIRE(K % 8 == 0, "gemm: K=%d must be a multiple of 8 (TMA row stride is 16-byte granular)", K);
  VB_REQUIRE(N % 8 == 0, "gemm: N=%d must be a multiple of 8 (16-byte vector stores)", N);
  VB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "gemm: out must be 16-byte aligned");
  VB_REQUIRE(epi == VB200_EPI_NONE || bias, "gemm: epilogue %d needs bias", static_cast<int>(epi));
  VB_REQUIRE(epi != VB200_EPI_BIAS_RESIDUAL || (residual && out_dtype == VB200_F32),
             "gemm: BIAS_RESIDUAL needs residual and fp32 output");
  if (M == 0) return VB200_OK;
  CUtensorMap ta, tb;
  int rc = cached_tmap(&ta, A, K, M, static_cast<uint64_t>(K) * 2, gemm::BK, gemm::BM);
  if (rc != VB200_OK) return rc;
  rc = cached_tmap(&tb, W, K, N, static_cast<uint64_t>(K) * 2, gemm::BK, gemm::BN);
  if (rc != VB200_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (epi) {
    case VB200_EPI_NONE: return launch_gemm_dtype<VB200_EPI_NONE>(out, out_dtype, ta, tb, bias, residual, M, N, K, st);
    case VB200_EPI_BIAS: return launch_gemm_dtype<VB200_EPI_BIAS>(out, out_dtype, ta, tb, bias, residual, M, N, K, st);
    case VB200_EPI_BIAS_GELU: return launch_gemm_dtype<VB200_EPI_BIAS_GELU>(out, out_dtype, ta, tb, bias, residual, M, N, K, st);
    case VB200_EPI_BIAS_RESIDUAL: return launch_gemm<VB200_EPI_BIAS_RESIDUAL, float>(out, ta, tb, bias, residual, M, N, K, st);
  }
  set_error("gemm: unknown epilogue %d", static_cast<int>(epi));
  return VB200_ERR_INVALID;
}
