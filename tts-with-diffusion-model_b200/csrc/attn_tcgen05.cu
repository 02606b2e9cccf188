// Non-causal variable-length flash attention, head_dim 64 (reference
// base.py:112-127 materialises (b, i, j, h) scores + mask + softmax in HBM).
//
// One CTA = (utterance, head, PAIR of 128-query tiles A/B); it walks the utterance's keys in
// blocks of 128 and keeps the tensor pipe busy by ping-ponging the two tiles.  12 warps:
//   warp 0 lane 0 : TMA producer  — Q_A, Q_B once, then K/V blocks through a 3-stage ring
//   warp 1 lane 0 : MMA issuer    — S_X = Q_X K^T   (tcgen05.mma 128xNx16, SS, both K-major)
//                                   O_X += P_X V    (128x64x16, A = P_X from TMEM, B = V straight
//                                                    from the TMA tile as an MN-major operand)
//   warps 4..7    : softmax of tile A, warps 8..11: softmax of tile B — thread = one query row:
//                   tcgen05.ld of the 128 scores (max pass, then exp2 pass against a lazily updated
//                   reference max); P (bf16) goes to its own TMEM columns and S_X is released
//                   (s_free) as soon as its last 64 columns sit in registers, so the MMA warp issues
//                   S_X(j+1) half an exp pass before P_X(j) V_j.
// O accumulates in TMEM across key blocks (fp32); it is rescaled only when the row max grows by
// more than 2^8 (exact: the common factor cancels in O / l), so the steady state has no TMEM round
// trip for O.  TMEM (512 columns): S_A 128 | S_B 128 | P_A 64 | P_B 64 | O_A 64 | O_B 64.
// Keys past the utterance end are masked to -inf — the reference's key-padding mask
// (base.py:119-124) in the packed-row layout — and the last key block only issues the MMAs
// (N resp. K rounded up to 16) its valid keys need.
// Measured (profiles/): the kernel is bound by the softmax -> MMA -> softmax dependency chain
// (mbarrier hops + MMA latency, ~900 cycles per block), not by MUFU or the tensor pipe: removing
// the exponentials does not change its time.  DESIGN.md §4 lists what was tried against that.
#include "common.cuh"

namespace vb200 {

namespace attn {
constexpr int BQ = 128, BKV = 128, HD = 64, KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB: 128 rows x 128 B
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0, COL_P = 256, COL_O = 384;   // S_X at 128*X, P_X at 256+64*X, O_X at 384+64*X
constexpr int THREADS = 12 * 32;          // control warpgroup + two softmax warpgroups
constexpr int SMEM_BYTES = TILE_BYTES * (2 + 2 * KV_STAGES) + 1024 + 256;
constexpr float RESCALE_LOG2 = 8.0f;
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
__device__ __forceinline__ void tmem_ld_32x32p(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32p(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
      "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// 32 scores -> 32 probabilities (bf16, 16 TMEM columns); row sum tracked on 4 chains.
__device__ __forceinline__ void exp_store_chunk(const uint32_t (&s)[32], uint32_t t_p_chunk, float scale_log2,
                                                float mneg, float (&ps)[4], uint64_t* wait_bar,
                                                uint32_t wait_parity) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), scale_log2, mneg));
    const float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), scale_log2, mneg));
    const float p2 = ex2_approx(fmaf(__uint_as_float(s[i + 2]), scale_log2, mneg));
    const float p3 = ex2_approx(fmaf(__uint_as_float(s[i + 3]), scale_log2, mneg));
    ps[0] += p0; ps[1] += p1; ps[2] += p2; ps[3] += p3;
    pk[i >> 1] = pack_bf16x2(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
  }
  if (wait_bar) {             // P_x(j-1) is still being read by PV_x(j-1) until this barrier flips
    mbar_wait(wait_bar, wait_parity);
    tc_fence_after();
  }
  tmem_st_32x16(t_p_chunk, pk);
}

// max of 32 scores on 4 chains; TAIL masks keys beyond the utterance (and writes the mask back)
template <bool TAIL>
__device__ __forceinline__ void max_chunk(uint32_t (&s)[32], int k_base, int last_valid, float (&bm)[4]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float v0 = __uint_as_float(s[i]), v1 = __uint_as_float(s[i + 1]);
    float v2 = __uint_as_float(s[i + 2]), v3 = __uint_as_float(s[i + 3]);
    if (TAIL) {
      if (k_base + i >= last_valid) v0 = -INFINITY;
      if (k_base + i + 1 >= last_valid) v1 = -INFINITY;
      if (k_base + i + 2 >= last_valid) v2 = -INFINITY;
      if (k_base + i + 3 >= last_valid) v3 = -INFINITY;
      s[i] = __float_as_uint(v0); s[i + 1] = __float_as_uint(v1);
      s[i + 2] = __float_as_uint(v2); s[i + 3] = __float_as_uint(v3);
    }
    bm[0] = fmaxf(bm[0], v0); bm[1] = fmaxf(bm[1], v1); bm[2] = fmaxf(bm[2], v2); bm[3] = fmaxf(bm[3], v3);
  }
}

// Lazy reference max: rescale O_x / l only when the block max exceeds m_ref by more than 2^8.
__device__ __forceinline__ void update_reference(float bm, int j, float scale_log2, uint32_t t_o,
                                                 uint64_t* pv_done_x, float& m_ref, float& l) {
  using namespace attn;
  if (j == 0) { m_ref = bm; return; }
  const bool need = (bm - m_ref) * scale_log2 > RESCALE_LOG2;
  if (__any_sync(0xffffffffu, need)) {
    mbar_wait(pv_done_x, (j - 1) & 1);                    // O_x quiescent: PV_x(j-1) retired
    tc_fence_after();
    const float alpha = need ? ex2_approx((m_ref - bm) * scale_log2) : 1.0f;
    if (need) { m_ref = bm; l *= alpha; }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {                         // 16 columns at a time: this path is rare and
      uint32_t o[16];                                     // must not add to the register peak
      tmem_ld_32x16(t_o + c * 16, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
      tmem_st_32x16(t_o + c * 16, o);
    }
  }
}

// One key block of one query row: two passes over S_x in TMEM (max, then exp), 32 columns at a time.
template <bool TAIL>
__device__ __forceinline__ void softmax_block_classic(uint32_t t_s, uint32_t t_p, uint32_t t_o, uint64_t* pv_done_x,
                                                      uint64_t* s_free_x, int lane, int j, int n_chunks,
                                                      int last_valid, float scale_log2, float& m_ref, float& l) {
  float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    if (TAIL && c >= n_chunks) break;
    uint32_t s[32];
    tmem_ld_32x32p(t_s + c * 32, s);
    tmem_ld_wait();
    max_chunk<TAIL>(s, c * 32, last_valid, bm);
  }
  update_reference(fmaxf(fmaxf(bm[0], bm[1]), fmaxf(bm[2], bm[3])), j, scale_log2, t_o, pv_done_x, m_ref, l);
  float mneg = -m_ref * scale_log2;
  float ps[4] = {0.f, 0.f, 0.f, 0.f};
  float unused[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t prev = (j - 1) & 1;
  if (!TAIL) {
    // chunks 0 and 1 straight from TMEM; chunks 2 and 3 are pulled into registers together, which is
    // the last read of S_x(j): s_free lets the MMA warp issue S_x(j+1) half an exp pass early
    {
      uint32_t s[32];
      tmem_ld_32x32p(t_s, s);
      tmem_ld_wait();
      exp_store_chunk(s, t_p, scale_log2, mneg, ps, j > 0 ? pv_done_x : nullptr, prev);
    }
    {
      uint32_t s[32];
      tmem_ld_32x32p(t_s + 32, s);
      tmem_ld_wait();
      exp_store_chunk(s, t_p + 16, scale_log2, mneg, ps, nullptr, 0);
    }
    uint32_t s2[32], s3[32];
    tmem_ld_32x32p(t_s + 64, s2);
    tmem_ld_32x32p(t_s + 96, s3);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(s_free_x);
    exp_store_chunk(s2, t_p + 32, scale_log2, mneg, ps, nullptr, 0);
    asm volatile("" : "+f"(mneg));                        // keep the two chunks in order (register peak)
    exp_store_chunk(s3, t_p + 48, scale_log2, mneg, ps, nullptr, 0);
  } else {
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      if (c >= n_chunks) break;
      uint32_t s[32];
      tmem_ld_32x32p(t_s + c * 32, s);
      tmem_ld_wait();
      max_chunk<true>(s, c * 32, last_valid, unused);     // re-apply the key mask
      exp_store_chunk(s, t_p + c * 16, scale_log2, mneg, ps, (c == 0 && j > 0) ? pv_done_x : nullptr, prev);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(s_free_x);
  }
  l += (ps[0] + ps[1]) + (ps[2] + ps[3]);
}

__global__ void __launch_bounds__(attn::THREADS, 1) flash_attn_kernel(
    const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
    const int32_t* __restrict__ cu_rows, int n_heads, float scale_log2) {
  using namespace attn;
  const int b = blockIdx.z, h = blockIdx.y, qp = blockIdx.x;
  const int row0 = cu_rows[b];
  const int T = cu_rows[b + 1] - row0;
  if (qp * 2 * BQ >= T) return;                   // uniform early exit, before any allocation
  const bool has_b = qp * 2 * BQ + BQ < T;        // second tile of the pair holds valid rows
  const int nblk = (T + BKV - 1) / BKV;
  const int last_valid = T - (nblk - 1) * BKV;    // keys inside the utterance in the last block
  const int last_n = (last_valid + 15) & ~15;     // MMA extent of the last block (multiple of 16)
  const int d = n_heads * HD;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* s_q = smem;                            // Q_A, Q_B
  uint8_t* s_kv = smem + 2 * TILE_BYTES;          // stage s: K at s*2*TILE, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_kv + 2 * KV_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                   // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;       // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;        // [2]  S_X(j) complete
  uint64_t* p_full = s_full + 2;                  // [2]  P_X(j) written (one arrival per softmax warp)
  uint64_t* pv_done = p_full + 2;                 // [2]  O_X += P_X(j) V_j retired
  uint64_t* s_free = pv_done + 2;                 // [2]  last read of S_X(j) done (one arrival per softmax warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x], 1); mbar_init(&p_full[x], 4); mbar_init(&pv_done[x], 1); mbar_init(&s_free[x], 4);
    }
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

  if (warp < 4) {
    // ============================================================== control warpgroup
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------- TMA producer
      mbar_arrive_expect_tx(q_full, (has_b ? 2 : 1) * TILE_BYTES);
      tma_load_2d(s_q, &tm_qkv, q_full, h * HD, row0 + qp * 2 * BQ);
      if (has_b) tma_load_2d(s_q + TILE_BYTES, &tm_qkv, q_full, h * HD, row0 + qp * 2 * BQ + BQ);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sk = s_kv + s * 2 * TILE_BYTES;
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tma_load_2d(sk, &tm_qkv, &kv_full[s], d + h * HD, row0 + j * BKV);
        tma_load_2d(sk + TILE_BYTES, &tm_qkv, &kv_full[s], 2 * d + h * HD, row0 + j * BKV);
      }
    } else if (warp == 1 && lane == 0) {
      // ---------------------------------------------------------- MMA issuer
      // One thread feeds the tensor pipe; everything it needs per MMA is a 32-bit add on a
      // precomputed descriptor (smem addresses are < 2^18, so the 14-bit start-address field of
      // the low word never carries).
      const uint32_t idesc_o = umma_idesc_bf16(BQ, HD, false, true);        // B = V, MN-major
      const uint32_t idesc_s_full = umma_idesc_bf16(BQ, BKV, false, false);
      const uint32_t idesc_s_last = umma_idesc_bf16(BQ, last_n, false, false);
      const int n_tiles = has_b ? 2 : 1;
      const uint64_t dq_base = umma_desc_kmajor_sw128(smem_u32(s_q));
      const uint64_t dk_base = umma_desc_kmajor_sw128(smem_u32(s_kv));
      const uint64_t dv_base = umma_desc_mnmajor_sw128(smem_u32(s_kv + TILE_BYTES), 1024);
      constexpr uint32_t kTileStep = TILE_BYTES >> 4;          // descriptor units (16 B)
      auto issue_s = [&](int x, int j) {      // S_x = Q_x K_j^T
        const uint64_t dq = dq_base + static_cast<uint32_t>(x) * kTileStep;
        const uint64_t dk = dk_base + static_cast<uint32_t>(j % KV_STAGES) * (2 * kTileStep);
        const uint32_t idesc_s = (j == nblk - 1) ? idesc_s_last : idesc_s_full;
        const uint32_t t_dst = tmem_base + COL_S + x * 128;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_dst, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[x]);
      };
      auto issue_pv = [&](int x, int j) {     // O_x (+)= P_x V_j
        const uint64_t dv = dv_base + static_cast<uint32_t>(j % KV_STAGES) * (2 * kTileStep);
        const uint32_t t_p = tmem_base + COL_P + x * 64;
        const uint32_t t_o = tmem_base + COL_O + x * 64;
        if (j != nblk - 1) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)                   // 16 key rows = 16 * 128 B = 128 units
            umma_ts(t_o, t_p + k * 8, dv + k * 128, idesc_o, (j | k) != 0);
        } else {
          const int ksteps = last_n / 16;
          for (int k = 0; k < ksteps; ++k) umma_ts(t_o, t_p + k * 8, dv + k * 128, idesc_o, (j | k) != 0);
        }
        umma_commit(&pv_done[x]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      for (int x = 0; x < n_tiles; ++x) issue_s(x, 0);
      for (int j = 0; j < nblk; ++j) {
        const bool more = j + 1 < nblk;
        if (more) mbar_wait(&kv_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1);
        for (int x = 0; x < n_tiles; ++x) {
          if (more) {
            mbar_wait(&s_free[x], j & 1);                // S_x(j) fully read: next scores first,
            tc_fence_after();                            // the softmax warps wait on these
            issue_s(x, j + 1);
          }
          mbar_wait(&p_full[x], j & 1);                  // P_x(j) in TMEM, O_x rescaled if needed
          tc_fence_after();
          issue_pv(x, j);
        }
        umma_commit(&kv_empty[j % KV_STAGES]);           // K_j / V_j consumed by both tiles
      }
    }
  } else {
    // ============================================================== softmax / output warpgroups
    const int x = (warp - 4) >> 2;                     // 0: tile A, 1: tile B
    if (x == 0 || has_b) {
      const int quad = warp & 3;
      const int r_tile = quad * 32 + lane;             // query row inside the tile
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      const uint32_t t_s = tmem_base + lane_off + COL_S + x * 128;
      const uint32_t t_p = tmem_base + lane_off + COL_P + x * 64;
      const uint32_t t_o = tmem_base + lane_off + COL_O + x * 64;
      float m_ref = -INFINITY, l = 0.f;

      for (int j = 0; j < nblk; ++j) {
        const bool tail = (j == nblk - 1) && last_valid < BKV;
        const int n_chunks = tail ? (last_n + 31) / 32 : 4;
        mbar_wait(&s_full[x], j & 1);
        tc_fence_after();
        if (!tail) softmax_block_classic<false>(t_s, t_p, t_o, &pv_done[x], &s_free[x], lane, j, 4, BKV, scale_log2, m_ref, l);
        else softmax_block_classic<true>(t_s, t_p, t_o, &pv_done[x], &s_free[x], lane, j, n_chunks, last_valid, scale_log2, m_ref, l);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);           // one arrival per warp
      }
      // epilogue: O / l -> bf16 rows
      mbar_wait(&pv_done[x], (nblk - 1) & 1);
      tc_fence_after();
      const int q_row = (qp * 2 + x) * BQ + r_tile;
      const float inv = 1.0f / l;
      __nv_bfloat16* o_dst = out + static_cast<size_t>(row0 + q_row) * d + h * HD;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld_32x32p(t_o + c * 32, o);
        tmem_ld_wait();
        if (q_row < T) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
            pk.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
            pk.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
            pk.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
            *reinterpret_cast<uint4*>(o_dst + c * 32 + i) = pk;
          }
        }
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

}  // namespace vb200

using namespace vb200;

extern "C" int vb200_flash_attn_varlen(void* out_bf16, const void* qkv_bf16, const int32_t* cu_rows,
                                       int32_t B, int32_t max_T, int32_t M, int32_t n_heads,
                                       float scale, vb200_stream_t stream) {
  using namespace attn;
  VB_REQUIRE(out_bf16 && qkv_bf16 && cu_rows, "flash_attn: null pointer");
  VB_REQUIRE(B >= 1 && B <= 65535 && n_heads >= 1 && n_heads <= 65535 && max_T >= 1 && M >= 0,
             "flash_attn: bad sizes B=%d heads=%d max_T=%d M=%d", B, n_heads, max_T, M);
  if (M == 0) return VB200_OK;
  const int d = n_heads * HD;
  CUtensorMap tm;
  int rc = cached_tmap(&tm, VB200_BF16, qkv_bf16, static_cast<uint64_t>(3) * d, M,
                       static_cast<uint64_t>(3) * d * 2, HD, 128);
  if (rc != VB200_OK) return rc;
  dim3 grid((max_T + 2 * BQ - 1) / (2 * BQ), n_heads, B);
  const float sl2 = scale * 1.4426950408889634f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out_bf16);
  static bool configured = false;
  if (!configured) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  flash_attn_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tm, o, cu_rows, n_heads, sl2);
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}
