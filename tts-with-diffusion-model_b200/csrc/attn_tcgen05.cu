// Non-causal variable-length flash attention, head_dim 64 (reference
// base.py:112-127 materialises (b, i, j, h) scores + mask + softmax in HBM).
//
// One CTA = (utterance, head, one 128-query tile); it walks the utterance's keys in blocks of 128.
// TWO CTAs share an SM (256 TMEM columns and ~81 KB of shared memory each), so one tile's
// start-up, drain and barrier round trips are covered by the other tile's steady state, and the
// nearly empty last tile of a ragged utterance only idles half an SM.  6 warps:
//   warp 0        : TMA producer  — Q once, then K and V blocks through two 2-slot rings (K_j is
//                   released as soon as S(j) has retired, V_j when O_blk(j) has)
//   warp 1        : MMA issuer (one elected lane)
//                                   S = Q K^T       (tcgen05.mma 128xNx16, SS, both K-major)
//                                   O_blk = P V     (128x64x16, A = P from TMEM, B = V straight
//                                                    from the TMA tile as an MN-major operand)
//   warps 2..5    : softmax — thread = one query row.
// Measured on B200 (profiles/): a softmax -> MMA -> softmax round trip (mbarrier hops + MMA
// latency) costs ~900 cycles, more than the MMAs of a block, and with one score buffer per tile
// that round trip sits on the critical path of every key block.  So the tile owns TWO 128-column
// TMEM buffers:
//     buffer (j & 1):  S(j) [128 cols]  ->  P(j) bf16 in cols 0..63 (written over the consumed
//     scores)  ->  O_blk(j) = P(j) V_j fp32 in cols 64..127
// S(j+1) is computed into the other buffer while block j is still being exponentiated, so a
// softmax warp normally finds its next scores ready.  O is accumulated in registers (fp32, the
// standard online-softmax recurrence o = o * alpha_j + O_blk(j)); block j-1's O_blk is folded in
// halfway through block j's exp pass, which is also what frees that buffer for S(j+1).
// The tensor pipe is not the limit: a block's 4 + 8 MMAs sustain ~570 cycles (tools/mma_bench.cu)
// against ~1500 for the softmax of a 128x128 block, which is co-limited by instruction issue and
// MUFU.EX2 — hence the packed f32x2 arithmetic and the polynomial exp2 below.
// Keys past the utterance end are masked to -inf — the reference's key-padding mask
// (base.py:119-124) in the packed-row layout — and the last key block only issues the MMAs
// (N resp. K rounded up to 16) its valid keys need.
#include "common.cuh"

namespace vb200 {

namespace attn {
constexpr int BQ = 128, BKV = 128, HD = 64, KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB: 128 rows x 128 B
constexpr uint32_t TMEM_COLS = 256;         // two CTAs share an SM
constexpr int THREADS = 6 * 32;           // TMA warp, MMA warp, four softmax warps
constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * KV_STAGES) + 1024 + 256;
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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// exp2 on the FMA / ALU pipes (no MUFU), two at once with packed f32x2 arithmetic: round-to-nearest
// split x = n + r with the 1.5 * 2^23 magic constant, 2^r by a minimax cubic on [-0.5, 0.5]
// (|rel err| < 7.5e-5, far below the bf16 rounding of P), exponent patched in with an integer add.
// MUFU.EX2 runs at 16 lanes / clk / SM on B200 (tools/mufu_bench.cu; the f16x2 form is two MUFU
// operations, no faster), which makes the softmax of a 128x128 block co-limited by MUFU and
// instruction issue; moving a share of the exponentials here rebalances the two (the FA4 trick).
__device__ __forceinline__ void ex2_poly2(uint64_t x, float& p0, float& p1) {
  float x0, x1;
  unpack2(x, x0, x1);
  x = pack2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t t = fadd2(x, pack2(12582912.0f, 12582912.0f));
  const uint64_t u = fadd2(t, pack2(-12582912.0f, -12582912.0f));
  const uint64_t r = ffma2(u, pack2(-1.0f, -1.0f), x);
  uint64_t p = ffma2(r, pack2(0.05517167f, 0.05517167f), pack2(0.24261113f, 0.24261113f));
  p = ffma2(r, p, pack2(0.69326097f, 0.69326097f));
  p = ffma2(r, p, pack2(0.99992806f, 0.99992806f));
  float t0, t1, q0, q1;
  unpack2(t, t0, t1);
  unpack2(p, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
#ifndef VB200_ATTN_EMU_PAIRS
#define VB200_ATTN_EMU_PAIRS 4      // of the 16 score pairs of a chunk, how many take ex2_poly2 instead of MUFU
#endif

// 32 scores -> 32 probabilities (bf16 pairs in pk); row sum tracked on two packed chains.
__device__ __forceinline__ void exp_chunk(const uint32_t (&s)[32], uint32_t (&pk)[16], uint64_t scale2,
                                          uint64_t mneg2, uint64_t (&ps)[2]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t x = ffma2(pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), scale2, mneg2);
    const bool emulate = ((i + 1) * VB200_ATTN_EMU_PAIRS) / 16 != (i * VB200_ATTN_EMU_PAIRS) / 16;
    float p0, p1;
    if (emulate) {
      ex2_poly2(x, p0, p1);
    } else {
      float x0, x1;
      unpack2(x, x0, x1);
      p0 = ex2_approx(x0);
      p1 = ex2_approx(x1);
    }
    ps[i & 1] = fadd2(ps[i & 1], pack2(p0, p1));
    pk[i] = pack_bf16x2(p0, p1);
  }
}
__device__ __forceinline__ void exp_store_chunk(const uint32_t (&s)[32], uint32_t t_p_chunk, uint64_t scale2,
                                                uint64_t mneg2, uint64_t (&ps)[2]) {
  uint32_t pk[16];
  exp_chunk(s, pk, scale2, mneg2, ps);
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

// o = o * alpha + O_blk, O_blk read from TMEM columns [t_oblk, t_oblk + 64)
__device__ __forceinline__ void fold_o_block(uint64_t (&o)[32], uint32_t t_oblk, float alpha) {
  const uint64_t alpha2 = pack2(alpha, alpha);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_32x32p(t_oblk + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      o[c * 16 + i] = ffma2(o[c * 16 + i], alpha2, pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])));
  }
}

// One key block j of one query row.  t_buf = this block's TMEM buffer (scores in, P out),
// t_prev = the other buffer, whose columns 64..127 hold O_blk(j-1).
template <bool TAIL>
__device__ __forceinline__ void softmax_block(uint32_t t_buf, uint32_t t_prev, uint64_t* pv_done_prev,
                                              uint32_t pv_parity, uint64_t* buf_free_prev, int lane, int j,
                                              int n_chunks, int last_valid, float scale_log2, float& m,
                                              float& l, float& alpha_prev, uint64_t (&o)[32]) {
  // pass 1: block max -> new running max, rescale factor of everything accumulated so far
  float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    if (TAIL && c >= n_chunks) break;
    uint32_t s[32];
    tmem_ld_32x32p(t_buf + c * 32, s);
    tmem_ld_wait();
    max_chunk<TAIL>(s, c * 32, last_valid, bm);
  }
  const float m_new = fmaxf(m, fmaxf(fmaxf(bm[0], bm[1]), fmaxf(bm[2], bm[3])));
  const float alpha = ex2_approx((m - m_new) * scale_log2);   // 0 on the first block (m = -inf)
  m = m_new;
  const float mneg = -m_new * scale_log2;
  const uint64_t scale2 = pack2(scale_log2, scale_log2), mneg2 = pack2(mneg, mneg);
  uint64_t ps[2] = {0ull, 0ull};
  float unused[4] = {0.f, 0.f, 0.f, 0.f};
  // pass 2, first half
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    if (TAIL && c >= n_chunks) break;
    uint32_t s[32];
    tmem_ld_32x32p(t_buf + c * 32, s);
    tmem_ld_wait();
    if (TAIL) max_chunk<true>(s, c * 32, last_valid, unused);     // re-apply the key mask
    exp_store_chunk(s, t_buf + c * 16, scale2, mneg2, ps);
  }
  // fold in O_blk(j-1) (its MMAs were issued a whole block ago) and hand its buffer back to the MMA warp
  if (j > 0) {
    mbar_wait(pv_done_prev, pv_parity);
    tc_fence_after();
    fold_o_block(o, t_prev + 64, alpha_prev);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(buf_free_prev);
  }
  // pass 2, second half (scores of columns 64..127; their P goes to columns 32..63, already consumed)
#pragma unroll 1
  for (int c = 2; c < 4; ++c) {
    if (TAIL && c >= n_chunks) break;
    uint32_t s[32];
    tmem_ld_32x32p(t_buf + c * 32, s);
    tmem_ld_wait();
    if (TAIL) max_chunk<true>(s, c * 32, last_valid, unused);
    exp_store_chunk(s, t_buf + c * 16, scale2, mneg2, ps);
  }
  float s0, s1, s2, s3;
  unpack2(ps[0], s0, s1);
  unpack2(ps[1], s2, s3);
  l = fmaf(l, alpha, (s0 + s1) + (s2 + s3));
  alpha_prev = alpha;
}

// 64-column TMEM load / 32-column store for the full-block path below
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
      "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_lo(uint32_t taddr, const uint32_t (&r)[64]) {
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

// A FULL key block (128 valid keys) handled in 64-score halves: twice the independent work between
// two TMEM round trips as the 32-score chunks of the tail path.  The packed pair i replaces s[i]
// (already consumed: pair i reads s[2i], s[2i+1]), so a half never needs more than its own 64
// registers beside the 64 accumulators.
__device__ __forceinline__ void softmax_block_full(uint32_t t_buf, uint32_t t_prev, uint64_t* pv_done_prev,
                                                   uint32_t pv_parity, uint64_t* buf_free_prev, int lane, int j,
                                                   float scale_log2, float& m, float& l, float& alpha_prev,
                                                   uint64_t (&o)[32]) {
  float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
  for (int hb = 0; hb < 2; ++hb) {
    uint32_t s[64];
    tmem_ld_32x64(t_buf + hb * 64, s);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      bm[0] = fmaxf(bm[0], __uint_as_float(s[i]));     bm[1] = fmaxf(bm[1], __uint_as_float(s[i + 1]));
      bm[2] = fmaxf(bm[2], __uint_as_float(s[i + 2])); bm[3] = fmaxf(bm[3], __uint_as_float(s[i + 3]));
    }
  }
  const float m_new = fmaxf(m, fmaxf(fmaxf(bm[0], bm[1]), fmaxf(bm[2], bm[3])));
  const float alpha = ex2_approx((m - m_new) * scale_log2);   // 0 on the first block (m = -inf)
  m = m_new;
  const float mneg = -m_new * scale_log2;
  const uint64_t scale2 = pack2(scale_log2, scale_log2), mneg2 = pack2(mneg, mneg);
  uint64_t ps[2] = {0ull, 0ull};
#pragma unroll 1
  for (int hb = 0; hb < 2; ++hb) {
    if (hb == 1 && j > 0) {
      // fold in O_blk(j-1) (its MMAs were issued a whole block ago) and hand its buffer back to the MMA warp
      mbar_wait(pv_done_prev, pv_parity);
      tc_fence_after();
      fold_o_block(o, t_prev + 64, alpha_prev);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(buf_free_prev);
    }
    uint32_t s[64];
    tmem_ld_32x64(t_buf + hb * 64, s);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const uint64_t x = ffma2(pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), scale2, mneg2);
      const int ii = i & 15;
      const bool emulate = ((ii + 1) * VB200_ATTN_EMU_PAIRS) / 16 != (ii * VB200_ATTN_EMU_PAIRS) / 16;
      float p0, p1;
      if (emulate) {
        ex2_poly2(x, p0, p1);
      } else {
        float x0, x1;
        unpack2(x, x0, x1);
        p0 = ex2_approx(x0);
        p1 = ex2_approx(x1);
      }
      ps[i & 1] = fadd2(ps[i & 1], pack2(p0, p1));
      s[i] = pack_bf16x2(p0, p1);
    }
    tmem_st_32x32_lo(t_buf + hb * 32, s);
  }
  float s0, s1, s2, s3;
  unpack2(ps[0], s0, s1);
  unpack2(ps[1], s2, s3);
  l = fmaf(l, alpha, (s0 + s1) + (s2 + s3));
  alpha_prev = alpha;
}

__global__ void __launch_bounds__(attn::THREADS, 2) flash_attn_kernel(
    const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
    const int32_t* __restrict__ cu_rows, int n_heads, float scale_log2) {
  using namespace attn;
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  const int row0 = cu_rows[b];
  const int T = cu_rows[b + 1] - row0;
  if (qt * BQ >= T) return;                       // uniform early exit, before any allocation
  const int nblk = (T + BKV - 1) / BKV;
  const int last_valid = T - (nblk - 1) * BKV;    // keys inside the utterance in the last block
  const int last_n = (last_valid + 15) & ~15;     // MMA extent of the last block (multiple of 16)
  const int d = n_heads * HD;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* s_q = smem;                            // Q tile
  uint8_t* s_k = smem + TILE_BYTES;               // K ring [KV_STAGES]
  uint8_t* s_v = s_k + KV_STAGES * TILE_BYTES;    // V ring [KV_STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_v + KV_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;                    // [KV_STAGES]
  uint64_t* k_empty = k_full + KV_STAGES;         // [KV_STAGES]  S(j) retired: K_j no longer needed
  uint64_t* v_full = k_empty + KV_STAGES;         // [KV_STAGES]
  uint64_t* v_empty = v_full + KV_STAGES;         // [KV_STAGES]  O_blk(j) retired: V_j no longer needed
  // per TMEM buffer bf; every one of them completes once per two blocks
  uint64_t* s_full = v_empty + KV_STAGES;         // [2]  scores of the block in this buffer complete
  uint64_t* p_full = s_full + 2;                  // [2]  P written (one arrival per softmax warp)
  uint64_t* pv_done = p_full + 2;                 // [2]  O_blk = P V retired
  uint64_t* buf_free = pv_done + 2;               // [2]  O_blk folded into registers: buffer reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(buf_free + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1); mbar_init(&buf_free[i], 4);
    }
    fence_barrier_init();
    // Q and the first K / V block go out before the block-wide sync below (TMEM allocation, barrier
    // visibility for the other warps): the first TMA round trip is on every CTA's critical path.
    pdl_launch_dependents();
    pdl_wait();                                   // qkv is the previous kernel's output
    mbar_arrive_expect_tx(q_full, TILE_BYTES);
    tma_load_2d(s_q, &tm_qkv, q_full, h * HD, row0 + qt * BQ);
    mbar_arrive_expect_tx(&k_full[0], TILE_BYTES);
    tma_load_2d(s_k, &tm_qkv, &k_full[0], d + h * HD, row0);
    mbar_arrive_expect_tx(&v_full[0], TILE_BYTES);
    tma_load_2d(s_v, &tm_qkv, &v_full[0], 2 * d + h * HD, row0);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: cu_rows (static batch layout) was read above; qkv and `out` only from here on
  pdl_launch_dependents();
  pdl_wait();

  // The two control warps run their loops WARP-UNIFORMLY (all 32 lanes compute the same
  // addresses / descriptors, one elected lane executes the TMA / MMA / commit instructions).
  // Issuing from a divergent `if (lane == 0)` region makes the compiler wrap every tcgen05.mma in
  // an elect / R2UR "waterfall" loop: measured 129 cycles per MMA regardless of shape instead of
  // 56-99 (tools/mma_bench.cu), which alone made this kernel tensor-issue bound.
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    for (int j = 1; j < nblk; ++j) {              // Q and block 0 were issued in the prologue
      const int s = j % KV_STAGES;
      const uint32_t ph = (j / KV_STAGES) & 1;
      mbar_wait(&k_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
        tma_load_2d(s_k + s * TILE_BYTES, &tm_qkv, &k_full[s], d + h * HD, row0 + j * BKV);
      }
      __syncwarp();
      mbar_wait(&v_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
        tma_load_2d(s_v + s * TILE_BYTES, &tm_qkv, &v_full[s], 2 * d + h * HD, row0 + j * BKV);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // Everything needed per MMA is a 32-bit add on a precomputed descriptor (smem addresses are
    // < 2^18, so the 14-bit start-address field of the low word never carries).
    const bool leader = elect_one();
    const uint32_t idesc_o = umma_idesc_bf16(BQ, HD, false, true);        // B = V, MN-major
    const uint32_t idesc_s_full = umma_idesc_bf16(BQ, BKV, false, false);
    const uint32_t idesc_s_last = umma_idesc_bf16(BQ, last_n, false, false);
    const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(s_q));
    const uint64_t dk_base = umma_desc_kmajor_sw128(smem_u32(s_k));
    const uint64_t dv_base = umma_desc_mnmajor_sw128(smem_u32(s_v), 1024);
    constexpr uint32_t kTileStep = TILE_BYTES >> 4;          // descriptor units (16 B)
    auto issue_s = [&](int j) {             // S(j) = Q K_j^T into buffer j & 1
      const int st = j % KV_STAGES;
      const uint64_t dk = dk_base + static_cast<uint32_t>(st) * kTileStep;
      const uint32_t idesc_s = (j == nblk - 1) ? idesc_s_last : idesc_s_full;
      const int bi = j & 1;
      const uint32_t t_dst = tmem_base + bi * 128;
      if (leader) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_dst, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[bi]);
        umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int j) {            // O_blk(j) = P(j) V_j, inside buffer j & 1
      const int st = j % KV_STAGES;
      const uint64_t dv = dv_base + static_cast<uint32_t>(st) * kTileStep;
      const int bi = j & 1;
      const uint32_t t_p = tmem_base + bi * 128;
      const uint32_t t_o = t_p + 64;
      if (leader) {
        if (j != nblk - 1) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)                 // 16 key rows = 16 * 128 B = 128 units
            umma_ts(t_o, t_p + k * 8, dv + k * 128, idesc_o, k != 0);
        } else {
          const int ksteps = last_n / 16;
          for (int k = 0; k < ksteps; ++k) umma_ts(t_o, t_p + k * 8, dv + k * 128, idesc_o, k != 0);
        }
        umma_commit(&pv_done[bi]);
        umma_commit(&v_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    for (int j = 0; j < 2 && j < nblk; ++j) {              // both buffers start out free
      mbar_wait(&k_full[j % KV_STAGES], (j / KV_STAGES) & 1);
      tc_fence_after();
      issue_s(j);
    }
    for (int j = 0; j < nblk; ++j) {
      const uint32_t par = (j >> 1) & 1;
      mbar_wait(&v_full[j % KV_STAGES], (j / KV_STAGES) & 1);
      mbar_wait(&p_full[j & 1], par);                      // P(j) written
      tc_fence_after();
      issue_pv(j);
      if (j + 2 < nblk) {
        mbar_wait(&k_full[(j + 2) % KV_STAGES], ((j + 2) / KV_STAGES) & 1);
        mbar_wait(&buf_free[j & 1], par);                  // O_blk(j) folded: buffer j & 1 reusable
        tc_fence_after();
        issue_s(j + 2);
      }
    }
  } else {
    // ============================================================== softmax / output warps 2..5
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int r_tile = quad * 32 + lane;               // query row inside the tile
    const uint32_t t_x = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    if (qt * BQ + quad * 32 >= T) {
      // None of this warp's 32 query rows exists (ragged last tile: T = 1027 leaves 3 rows in the
      // ninth tile).  Keep the barrier protocol going — whatever sits in its P columns only ever
      // feeds its own, never stored, output rows — and leave the issue slots and the MUFU to the
      // other CTA on this SM.
      for (int j = 0; j < nblk; ++j) {
        const int bf = j & 1;
        mbar_wait(&s_full[bf], (j >> 1) & 1);
        if (j > 0) {
          mbar_wait(&pv_done[bf ^ 1], ((j - 1) >> 1) & 1);
          if (lane == 0) mbar_arrive(&buf_free[bf ^ 1]);
        }
        if (lane == 0) mbar_arrive(&p_full[bf]);
      }
      mbar_wait(&pv_done[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);
    } else {
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    uint64_t o[32];                                    // 64 fp32 accumulators as f32x2 pairs
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0ull;

    for (int j = 0; j < nblk; ++j) {
      const bool tail = (j == nblk - 1) && last_valid < BKV;
      const int n_chunks = tail ? (last_n + 31) / 32 : 4;
      const int bf = j & 1;
      mbar_wait(&s_full[bf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t t_buf = t_x + bf * 128, t_prev = t_x + (bf ^ 1) * 128;
      uint64_t* pvd = &pv_done[bf ^ 1];
      uint64_t* bfr = &buf_free[bf ^ 1];
      const uint32_t pv_par = ((j - 1) >> 1) & 1;
      if (!tail) softmax_block_full(t_buf, t_prev, pvd, pv_par, bfr, lane, j, scale_log2, m, l, alpha_prev, o);
      else softmax_block<true>(t_buf, t_prev, pvd, pv_par, bfr, lane, j, n_chunks, last_valid, scale_log2, m, l, alpha_prev, o);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[bf]);         // one arrival per warp
    }
    // last block's O_blk, then O / l -> bf16 rows
    const int jl = nblk - 1;
    mbar_wait(&pv_done[jl & 1], (jl >> 1) & 1);
    tc_fence_after();
    fold_o_block(o, t_x + (jl & 1) * 128 + 64, alpha_prev);
    const int q_row = qt * BQ + r_tile;
    if (q_row < T) {
      const float inv = 1.0f / l;
      __nv_bfloat16* o_dst = out + static_cast<size_t>(row0 + q_row) * d + h * HD;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack2(o[i + k], v[2 * k], v[2 * k + 1]);
        uint4 pk;
        pk.x = pack_bf16x2(v[0] * inv, v[1] * inv);
        pk.y = pack_bf16x2(v[2] * inv, v[3] * inv);
        pk.z = pack_bf16x2(v[4] * inv, v[5] * inv);
        pk.w = pack_bf16x2(v[6] * inv, v[7] * inv);
        *reinterpret_cast<uint4*>(o_dst + 2 * i) = pk;
      }
    }
    }  // warp with valid rows
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
  if (M == 0) return VB200_OK;
  VB_REQUIRE(out_bf16 && qkv_bf16 && cu_rows, "flash_attn: null pointer");
  VB_REQUIRE(B >= 1 && B <= 65535 && n_heads >= 1 && n_heads <= 65535 && max_T >= 1 && M >= 0,
             "flash_attn: bad sizes B=%d heads=%d max_T=%d M=%d", B, n_heads, max_T, M);
  if (M == 0) return VB200_OK;
  const int d = n_heads * HD;
  CUtensorMap tm;
  int rc = cached_tmap(&tm, VB200_BF16, qkv_bf16, static_cast<uint64_t>(3) * d, M,
                       static_cast<uint64_t>(3) * d * 2, HD, 128);
  if (rc != VB200_OK) return rc;
  dim3 grid((max_T + BQ - 1) / BQ, n_heads, B);
  const float sl2 = scale * 1.4426950408889634f;
  VB_CONFIGURE_SMEM(flash_attn_kernel, SMEM_BYTES);
  {
    static std::atomic<uint64_t> carveout_done{0};                         // per device, like the macro above
    int dev = 0;
    VB_CHECK_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(carveout_done.load(std::memory_order_acquire) & bit)) {
      VB_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));   // two CTAs per SM
      carveout_done.fetch_or(bit, std::memory_order_release);
    }
  }
  VB_CHECK_CUDA(launch_pdl(flash_attn_kernel, grid, dim3(THREADS), SMEM_BYTES, static_cast<cudaStream_t>(stream), 1,
                           tm, static_cast<__nv_bfloat16*>(out_bf16), cu_rows, n_heads, sl2));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}
