// Non-causal variable-length flash attention, head_dim 64 (reference
// base.py:112-127 materialises (b, i, j, h) scores + mask + softmax in HBM).
//
// One CTA = (utterance, head, 128-query tile); it walks the utterance's keys in blocks of 128.
//   warp 0 lane 0 : TMA producer  — Q tile once, then K/V blocks through a 2-stage ring
//   warp 1 lane 0 : MMA issuer    — S = Q K^T      (tcgen05.mma 128x128x16, SS, both K-major)
//                                   O_blk = P V    (128x64x16, A = P from TMEM (or smem),
//                                                   B = V straight from the TMA tile, MN-major)
//   warps 2..5    : softmax       — thread = one query row: tcgen05.ld S, online max/sum in
//                                   registers, P (bf16) -> TMEM, O_blk added into fp32 registers
// TMEM: 256 columns per CTA (S 128 | P 64 | O 64), so two CTAs share an SM and one CTA's MMAs
// overlap the other's exponentials.  Keys past the utterance end are masked to -inf, which is
// the reference's key-padding mask (base.py:119-124) in the packed-row layout.
#include "common.cuh"

namespace vb200 {

int cached_tmap(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer,
                uint64_t stride_bytes, uint32_t box_inner, uint32_t box_outer);

namespace attn {
constexpr int BQ = 128, BKV = 128, HD = 64, KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB: 128 rows x 128 B
constexpr int THREADS = 6 * 32;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;
template <bool P_TMEM>
constexpr int smem_bytes() {
  return TILE_BYTES * (1 + 2 * KV_STAGES + (P_TMEM ? 0 : 2)) + 1024 + 128;
}
}  // namespace attn

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
      "r"(r[14]), "r"(r[15])
      : "memory");
}

template <bool P_TMEM>
__global__ void __launch_bounds__(attn::THREADS, P_TMEM ? 2 : 1) flash_attn_kernel(
    const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
    const int32_t* __restrict__ cu_rows, int n_heads, float scale_log2) {
  using namespace attn;
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  const int row0 = cu_rows[b];
  const int T = cu_rows[b + 1] - row0;
  if (qt * BQ >= T) return;                       // uniform early exit, before any allocation
  const int nblk = (T + BKV - 1) / BKV;
  const int d = n_heads * HD;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* s_q = smem;
  uint8_t* s_kv = smem + TILE_BYTES;              // stage s: K at s*2*TILE, V right after
  uint8_t* s_p = s_kv + 2 * KV_STAGES * TILE_BYTES;   // only when !P_TMEM: two 16 KB K-major tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + (P_TMEM ? 0 : 2 * TILE_BYTES));
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                   // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;       // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
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
      // ---------------------------------------------------------- TMA producer
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(s_q, &tm_qkv, q_full, h * HD, row0 + qt * BQ);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sk = s_kv + s * 2 * TILE_BYTES;
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tma_load_2d(sk, &tm_qkv, &kv_full[s], d + h * HD, row0 + j * BKV);
        tma_load_2d(sk + TILE_BYTES, &tm_qkv, &kv_full[s], 2 * d + h * HD, row0 + j * BKV);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, false, true);   // B = V, MN-major
      const uint32_t t_s = tmem_base + COL_S, t_p = tmem_base + COL_P, t_o = tmem_base + COL_O;
      const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(s_q));
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(s_kv));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
      for (int j = 0; j < nblk; ++j) {
        const int s = j % KV_STAGES;
        mbar_wait(p_full, j & 1);                       // P_j written, S free
        if (j > 0) mbar_wait(o_empty, (j - 1) & 1);     // O_blk_{j-1} consumed
        tc_fence_after();
        const uint32_t sv = smem_u32(s_kv + s * 2 * TILE_BYTES + TILE_BYTES);
        if (P_TMEM) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            const uint64_t dv = umma_desc_mnmajor_sw128(sv + k * 16 * 128, 1024);
            umma_ts(t_o, t_p + k * 8, dv, idesc_o, k != 0);
          }
        } else {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            const uint64_t dp = umma_desc_kmajor_sw128(smem_u32(s_p) + (k / 4) * TILE_BYTES) + 2 * (k % 4);
            const uint64_t dv = umma_desc_mnmajor_sw128(sv + k * 16 * 128, 1024);
            umma_ss(t_o, dp, dv, idesc_o, k != 0);
          }
        }
        umma_commit(&kv_empty[s]);                      // K_j / V_j slot free once PV_j retires
        umma_commit(o_full);
        if (j + 1 < nblk) {
          const int s1 = (j + 1) % KV_STAGES;
          mbar_wait(&kv_full[s1], ((j + 1) / KV_STAGES) & 1);
          tc_fence_after();
          const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(s_kv + s1 * 2 * TILE_BYTES));
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_ss(t_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / output warps
    const int quad = warp & 3;
    const int r_tile = quad * 32 + lane;               // query row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + COL_S;
    const uint32_t t_p = tmem_base + lane_off + COL_P;
    const uint32_t t_o = tmem_base + lane_off + COL_O;
    float o_acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
    float mx = -INFINITY, l = 0.f, alpha_pending = 0.f;

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int n_valid = T - j * BKV;                 // keys of this block inside the utterance
      const bool tail = n_valid < BKV;
      // pass 1: block max
      float bm = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_s + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(r[i]);
          if (tail && c * 32 + i >= n_valid) s = -INFINITY;
          bm = fmaxf(bm, s);
        }
      }
      const float mx_new = fmaxf(mx, bm);
      const float alpha = exp2f((mx - mx_new) * scale_log2);   // 0 on the first block
      const float mneg = -mx_new * scale_log2;
      float psum = 0.f;
      // pass 2: P = exp2(s*c - m*c) -> bf16 -> TMEM (or swizzled smem)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_s + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = exp2f(fmaf(__uint_as_float(r[i]), scale_log2, mneg));
          float p1 = exp2f(fmaf(__uint_as_float(r[i + 1]), scale_log2, mneg));
          if (tail) {
            if (c * 32 + i >= n_valid) p0 = 0.f;
            if (c * 32 + i + 1 >= n_valid) p1 = 0.f;
          }
          psum += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        if (P_TMEM) {
          tmem_st_32x16(t_p + c * 16, pk);
        } else {
          // K-major SWIZZLE_128B tile: row r_tile, 16-byte chunk index XOR (row % 8)
          uint8_t* tile = s_p + (c >> 1) * TILE_BYTES + r_tile * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (c & 1) * 4 + q;
            *reinterpret_cast<uint4*>(tile + ((chunk ^ (r_tile & 7)) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
      }
      if (P_TMEM) tmem_st_wait(); else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      l = l * alpha + psum;
      mx = mx_new;
      // fold in the previous block's P V (deferred so this block's exponentials start early)
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_o + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha_pending, __uint_as_float(r[i]));
        }
        tc_fence_before();
        mbar_arrive(o_empty);
      }
      alpha_pending = alpha;
    }
    // last block's P V
    mbar_wait(o_full, (nblk - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(t_o + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha_pending, __uint_as_float(r[i]));
    }
    const int q_row = qt * BQ + r_tile;
    if (q_row < T) {
      const float inv = 1.0f / l;
      __nv_bfloat16* o = out + static_cast<size_t>(row0 + q_row) * d + h * HD;
#pragma unroll
      for (int i = 0; i < HD; i += 8) {
        uint4 p;
        p.x = pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv);
        p.y = pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
        p.z = pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv);
        p.w = pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
        *reinterpret_cast<uint4*>(o + i) = p;
      }
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

template <bool P_TMEM>
static int launch_attn(void* out, const void* qkv, const int32_t* cu_rows, int B, int max_T, int M,
                       int n_heads, float scale, cudaStream_t st) {
  using namespace attn;
  const int d = n_heads * HD;
  CUtensorMap tm;
  int rc = cached_tmap(&tm, qkv, static_cast<uint64_t>(3) * d, M, static_cast<uint64_t>(3) * d * 2, HD, 128);
  if (rc != VB200_OK) return rc;
  auto kern = flash_attn_kernel<P_TMEM>;
  static bool configured = false;
  if (!configured) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<P_TMEM>()));
    configured = true;
  }
  dim3 grid((max_T + BQ - 1) / BQ, n_heads, B);
  kern<<<grid, THREADS, smem_bytes<P_TMEM>(), st>>>(tm, static_cast<__nv_bfloat16*>(out), cu_rows,
                                                    n_heads, scale * 1.4426950408889634f);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}

static int attn_check(void* out, const void* qkv, const int32_t* cu_rows, int B, int max_T, int M,
                      int n_heads) {
  VB_REQUIRE(out && qkv && cu_rows, "flash_attn: null pointer");
  VB_REQUIRE(B >= 1 && B <= 65535 && n_heads >= 1 && n_heads <= 65535 && max_T >= 1 && M >= 0,
             "flash_attn: bad sizes B=%d heads=%d max_T=%d M=%d", B, n_heads, max_T, M);
  return VB200_OK;
}

}  // namespace vb200

using namespace vb200;

extern "C" int vb200_flash_attn_varlen(void* out_bf16, const void* qkv_bf16, const int32_t* cu_rows,
                                       int32_t B, int32_t max_T, int32_t M, int32_t n_heads,
                                       float scale, vb200_stream_t stream) {
  int rc = attn_check(out_bf16, qkv_bf16, cu_rows, B, max_T, M, n_heads);
  if (rc != VB200_OK) return rc;
  if (M == 0) return VB200_OK;
  return launch_attn<true>(out_bf16, qkv_bf16, cu_rows, B, max_T, M, n_heads, scale,
                           static_cast<cudaStream_t>(stream));
}

// Bring-up variant: P goes through shared memory (K-major SWIZZLE_128B) instead of TMEM.
extern "C" int vb200_flash_attn_varlen_psmem(void* out_bf16, const void* qkv_bf16,
                                             const int32_t* cu_rows, int32_t B, int32_t max_T,
                                             int32_t M, int32_t n_heads, float scale,
                                             vb200_stream_t stream) {
  int rc = attn_check(out_bf16, qkv_bf16, cu_rows, B, max_T, M, n_heads);
  if (rc != VB200_OK) return rc;
  if (M == 0) return VB200_OK;
  return launch_attn<false>(out_bf16, qkv_bf16, cu_rows, B, max_T, M, n_heads, scale,
                            static_cast<cudaStream_t>(stream));
}
