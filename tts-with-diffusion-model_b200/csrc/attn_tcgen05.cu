// Non-causal variable-length flash attention, head_dim 64 (reference
// base.py:112-127 materialises (b, i, j, h) scores + mask + softmax in HBM).
//
// One CTA = (utterance, head, one 128-query tile); it walks the utterance's keys in blocks of 128.
// TWO CTAs share an SM (256 TMEM columns and ~81 KB of shared memory each).  10 warps:
//   warps 0..7    : softmax.  Warp w owns 16 query rows — TMEM lanes 32 (w & 3) + 16 (w >> 2) + 0..15 —
//                   and a row is split over the two half-warps: lane t < 16 holds keys 0..63 of row t,
//                   lane t + 16 keys 64..127 (tcgen05.ld/st shape 16x32bx2).  Everything a row needs
//                   from its other half is one shuffle.
//   warp 8        : TMA producer — Q once, then K and V blocks through two 2-slot rings
//   warp 9        : MMA issuer (one elected lane)
//                     S = Q K^T       (tcgen05.mma 128xNx16, SS, both K-major)       -> TMEM cols   0..127
//                     O += P V        (128x64x16, A = P from TMEM cols 128..191, B = V straight from the
//                                      TMA tile as an MN-major operand), accumulated IN TMEM cols 192..255
// Why this shape (round-1 kernel: 4 softmax warps per CTA, thread = row, O folded into registers; ncu
// profiles/r1_attn_v9_ncu_raw.txt): the softmax of a 128x128 block needs ~770 MUFU cycles and ~650 issue
// slots per scheduler against ~570 cycles of MMA, but with TWO softmax warps per scheduler whose MUFU-heavy
// exp phases ran in lock step the MUFU pipe idled through everybody's other phases (second TMEM sweep for
// the row max, the fold of O_blk into registers, barrier round trips): 1 811 cycles per key block per SM,
// MUFU 48 % busy, issue slots 49 %.  Here
//   * O lives in TMEM and the P V MMAs accumulate into it: no fold, no per-block TMEM read of O;
//   * a warp reads its scores ONCE (64 registers per thread) and hands the score buffer straight back,
//     so S(j+1) is computed while block j is being exponentiated — the registers are the second buffer;
//   * the running maximum only moves when a block exceeds it by 2^8 (then the warp rescales its rows of
//     O in TMEM and its row sums): after the first block that is rare, and P <= 2^8 is harmless in bf16;
//   * half the registers per thread buy FOUR softmax warps per scheduler, enough independent work to
//     keep the MUFU pipe fed.
// Keys past the utterance end are masked to -inf — the reference's key-padding mask (base.py:119-124) in
// the packed-row layout — and the last key block only issues the MMAs (N resp. K rounded up to 16) its
// valid keys need.
#include "common.cuh"
#include <type_traits>

namespace vb200 {

namespace attn {
constexpr int BQ = 128, BKV = 128, HD = 64, KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB: 128 rows x 128 B
constexpr uint32_t TMEM_COLS = 256;       // two CTAs share an SM
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;
constexpr int SOFTMAX_WARPS = 8;
constexpr int THREADS = (SOFTMAX_WARPS + 2) * 32;
constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * KV_STAGES) + 1024 + 256;
constexpr float RESCALE_LOG2 = 8.0f;      // the reference maximum moves when a block exceeds it by 2^8
}  // namespace attn

// 16 TMEM lanes x 128 columns: thread t < 16 gets lane t, columns [c, c+64); thread t >= 16 gets lane t-16, columns [c+64, c+128)
__device__ __forceinline__ void tmem_ld_16x2_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x32bx2.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
      "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64], 64;"
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
// 16 lanes x 64 columns from r[0..31]: thread t < 16 writes lane t, columns [c, c+32); t >= 16 lane t-16, columns [c+32, c+64)
__device__ __forceinline__ void tmem_st_16x2_x32(uint32_t taddr, const uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x32bx2.x32.b32 [%0], 32, "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 16 lanes x 32 columns: t < 16 -> columns [c, c+16), t >= 16 -> [c+16, c+32)
__device__ __forceinline__ void tmem_ld_16x2_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x32bx2.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16], 16;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// store form of tmem_ld_16x2_x16
__device__ __forceinline__ void tmem_st_16x2_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x32bx2.x16.b32 [%0], 16, "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 16 lanes x 16 columns from 8 registers: t < 16 -> columns [c, c+8), t >= 16 -> [c+8, c+16)
__device__ __forceinline__ void tmem_st_16x2_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x32bx2.x8.b32 [%0], 8, {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
// 16 lanes x 64 columns: t < 16 -> columns [c, c+32), t >= 16 -> [c+32, c+64)
__device__ __forceinline__ void tmem_ld_16x2_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x32bx2.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32], 32;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
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
// operations, no faster), which makes the softmax MUFU bound; moving a share of the exponentials here
// trades MUFU cycles for issue slots (the FA4 trick).
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
#define VB200_ATTN_EMU_PAIRS 4      // of every 16 score pairs, how many take ex2_poly2 instead of MUFU
#endif
#ifndef VB200_ATTN_EXP_CHUNK
#define VB200_ATTN_EXP_CHUNK 4      // score pairs per scheduling chunk of the exp pass (see the kernel)
#endif
namespace attn { constexpr int EXP_CHUNK = VB200_ATTN_EXP_CHUNK; }

#ifdef VB200_ATTN_TRACE
// Debug build only (tools/attn_trace.py): per-CTA clock64() stamps of the softmax warp 0 and the MMA warp.
constexpr int TRACE_CTAS = 1024, TRACE_BLOCKS = 16, TRACE_EVENTS = 8;
__device__ unsigned long long g_attn_trace[TRACE_CTAS * 2 * TRACE_BLOCKS * TRACE_EVENTS];
#define VB_TR(slot, j, ev)                                                                               \
  do {                                                                                                   \
    if (lane == 0 && cta_lin < TRACE_CTAS && (j) < TRACE_BLOCKS)                                         \
      g_attn_trace[((cta_lin * 2 + (slot)) * TRACE_BLOCKS + (j)) * TRACE_EVENTS + (ev)] = clock64();     \
  } while (0)
#else
#define VB_TR(slot, j, ev) do {} while (0)
#endif

__global__ void __launch_bounds__(attn::THREADS, 2) flash_attn_kernel(
    const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
    const int32_t* __restrict__ cu_rows, int n_heads, float scale_log2, float opaque_zero) {
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
  uint64_t* v_empty = v_full + KV_STAGES;         // [KV_STAGES]  P V(j) retired: V_j no longer needed
  // one completion per key block each (parity j & 1)
  uint64_t* s_full = v_empty + KV_STAGES;         // scores of block j complete in TMEM
  uint64_t* s_free = s_full + 1;                  // every softmax warp holds its scores in registers
  uint64_t* p_full = s_free + 1;                  // P(j) written (and O rescaled where needed)
  uint64_t* pv_done = p_full + 1;                 // O += P(j) V_j retired: P buffer reusable, O readable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
#ifdef VB200_ATTN_TRACE
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (threadIdx.x == 0 && cta_lin < TRACE_CTAS) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_attn_trace[((cta_lin * 2 + 1) * TRACE_BLOCKS + TRACE_BLOCKS - 1) * TRACE_EVENTS + 7] = smid;
    g_attn_trace[((cta_lin * 2 + 1) * TRACE_BLOCKS + TRACE_BLOCKS - 1) * TRACE_EVENTS + 6] = clock64();
  }
#endif

  if (threadIdx.x == SOFTMAX_WARPS * 32) {        // first lane of the TMA warp
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, SOFTMAX_WARPS); mbar_init(p_full, SOFTMAX_WARPS); mbar_init(pv_done, 1);
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
  if (warp == SOFTMAX_WARPS + 1) {
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
  // 56-99 (tools/mma_bench.cu).
  if (warp == SOFTMAX_WARPS) {
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
  } else if (warp == SOFTMAX_WARPS + 1) {
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
    const uint32_t t_s = tmem_base + COL_S, t_p = tmem_base + COL_P, t_o = tmem_base + COL_O;
    auto issue_s = [&](int j) {             // S(j) = Q K_j^T
      const int st = j % KV_STAGES;
      const uint64_t dk = dk_base + static_cast<uint32_t>(st) * kTileStep;
      const uint32_t idesc_s = (j == nblk - 1) ? idesc_s_last : idesc_s_full;
      if (leader) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int j) {            // O (+)= P(j) V_j
      const int st = j % KV_STAGES;
      const uint64_t dv = dv_base + static_cast<uint32_t>(st) * kTileStep;
      const uint32_t acc0 = j != 0;
      if (leader) {
        if (j != nblk - 1) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)                 // 16 key rows = 16 * 128 B = 128 units
            umma_ts(t_o, t_p + k * 8, dv + k * 128, idesc_o, acc0 | (k != 0));
        } else {
          const int ksteps = last_n / 16;
          for (int k = 0; k < ksteps; ++k) umma_ts(t_o, t_p + k * 8, dv + k * 128, idesc_o, acc0 | (k != 0));
        }
        umma_commit(pv_done);
        umma_commit(&v_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nblk; ++j) {
      if (j + 1 < nblk) {                                    // S(j+1) runs under block j's exponentials
        mbar_wait(&k_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1);
        mbar_wait(s_free, j & 1);                            // S(j) is in the softmax warps' registers
        tc_fence_after();
        VB_TR(1, j, 0);
        issue_s(j + 1);
      }
      mbar_wait(&v_full[j % KV_STAGES], (j / KV_STAGES) & 1);
      mbar_wait(p_full, j & 1);                              // P(j) written
      tc_fence_after();
      VB_TR(1, j, 1);
      issue_pv(j);
    }
  } else {
    // ============================================================== softmax warps 0..7
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int half_rows = warp >> 2;                   // which 16 lanes of the quadrant
    const int t0 = lane & 15, t1 = lane >> 4;          // row inside the warp's 16, key half
    const int r_tile = quad * 32 + half_rows * 16 + t0;
    const uint32_t t_x = tmem_base + (static_cast<uint32_t>(quad * 32 + half_rows * 16) << 16);
    if (qt * BQ + quad * 32 + half_rows * 16 >= T) {
      // None of this warp's 16 query rows exists (ragged last tile).  Keep the barrier protocol going —
      // whatever sits in its P rows only ever feeds its own, never stored, output rows — and leave the
      // issue slots and the MUFU to the other warps on this SM.
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(s_full, j & 1);
        if (lane == 0) mbar_arrive(s_free);
        if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
        if (lane == 0) mbar_arrive(p_full);
      }
      mbar_wait(pv_done, (nblk - 1) & 1);
    } else {
      float m_ref = -INFINITY, l = 0.f;                  // reference maximum of the row; this half's share of the row sum
      const uint64_t zero2 = pack2(opaque_zero, opaque_zero);
      // The reference maximum moves when a block exceeds it by 2^8: rescales this warp's rows of O (TMEM) and l.
      auto move_reference = [&](float bm, int j) {
        if (j == 0) {
          m_ref = bm;                                      // block 0 holds at least one valid key
          return;
        }
        const bool grow = (bm - m_ref) * scale_log2 > RESCALE_LOG2;
        if (__any_sync(0xffffffffu, grow)) {
          // rare after the first block
          const float m_new = grow ? bm : m_ref;
          const float alpha = ex2_approx((m_ref - m_new) * scale_log2);      // 1 for rows that stay
          m_ref = m_new;
          l *= alpha;
          mbar_wait(pv_done, (j - 1) & 1);                 // O complete up to block j-1, P V(j) not issued yet
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[16];
            tmem_ld_16x2_x16(t_x + COL_O + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_16x2_x16(t_x + COL_O + c * 32, o);
          }
        }
      };
      // One 128-key block.  TAIL (compile time): keys >= last_valid are masked.  The full blocks run in
      // their own loop so that their code carries nothing of the ragged end.
      auto key_block = [&](int j, auto tail_tag) {
        constexpr bool tail = decltype(tail_tag)::value;
        if (warp == 0) VB_TR(0, j, 0);
        mbar_wait(s_full, j & 1);
        tc_fence_after();
        if (warp == 0) VB_TR(0, j, 1);
        uint32_t s[64];
        tmem_ld_16x2_x64(t_x + COL_S, s);
        tmem_ld_wait();
        if (warp == 0) VB_TR(0, j, 2);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);              // S(j+1) may overwrite the buffer
        if (tail) {                                      // keys past the utterance end (and stale columns)
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (t1 * 64 + i >= last_valid) s[i] = __float_as_uint(-INFINITY);
        }
        float bm0 = -INFINITY, bm1 = -INFINITY, bm2 = -INFINITY, bm3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          bm0 = fmaxf(bm0, __uint_as_float(s[i]));     bm1 = fmaxf(bm1, __uint_as_float(s[i + 1]));
          bm2 = fmaxf(bm2, __uint_as_float(s[i + 2])); bm3 = fmaxf(bm3, __uint_as_float(s[i + 3]));
        }
        float bm = fmaxf(fmaxf(bm0, bm1), fmaxf(bm2, bm3));
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));       // the row's other half
        move_reference(bm, j);
        const float mneg = -m_ref * scale_log2;
#ifdef VB200_ATTN_TRACE
        if (warp == 0 && mneg != 12345.f) VB_TR(0, j, 3);     // after the row maximum (data dependent: stays in place)
#endif
        const uint64_t scale2 = pack2(scale_log2, scale_log2), mneg2 = pack2(mneg, mneg);
        // The 32 score pairs go in chunks of EXP_CHUNK pairs, and chunk k's offset is made to DEPEND on the
        // sum of chunk k-2 (an FFMA2 with a zero the compiler cannot see through).  Without it ptxas hoists
        // all 64 scale FFMA2s and the polynomial exponentials in front of one run of 48 MUFU.EX2: the warps
        // of a CTA (released together by s_full) then all do FMA work with the MUFU idle, then all queue on
        // the MUFU with the FMA pipe idle (ncu source page: mio / wait stalls on every MUFU, pipe 52 % busy).
        uint64_t ps[2] = {0ull, 0ull};
#pragma unroll
        for (int k = 0; k < 32 / EXP_CHUNK; ++k) {
          const uint64_t mn2 = k >= 2 ? ffma2(ps[k & 1], zero2, mneg2) : mneg2;
#pragma unroll
          for (int i = k * EXP_CHUNK; i < (k + 1) * EXP_CHUNK; ++i) {
            const uint64_t x = ffma2(pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), scale2, mn2);
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
            ps[k & 1] = fadd2(ps[k & 1], pack2(p0, p1));
            s[i] = pack_bf16x2(p0, p1);                  // pair i replaces s[i] (already consumed)
          }
        }
        float a0, a1, a2, a3;
        unpack2(ps[0], a0, a1);
        unpack2(ps[1], a2, a3);
        l += (a0 + a1) + (a2 + a3);
#ifdef VB200_ATTN_TRACE
        if (warp == 0 && l != -1.f) VB_TR(0, j, 4);            // after the exp pass
#endif
        if (j > 0) {
          mbar_wait(pv_done, (j - 1) & 1);               // P V(j-1) has consumed the P buffer
          tc_fence_after();
        }
        if (warp == 0) VB_TR(0, j, 5);
        tmem_st_16x2_x32(t_x + COL_P, s);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);              // one arrival per warp
        if (warp == 0) VB_TR(0, j, 6);
      };
      // Last block of at most 32 keys (an utterance of T = 1027 rows ends in 3): one 32-column chunk, lane
      // t < 16 holds keys 0..15 of its row, lane t + 16 keys 16..31, instead of a masked 128-key pass.
      auto short_block = [&](int j) {
        mbar_wait(s_full, j & 1);
        tc_fence_after();
        uint32_t sq[16];
        tmem_ld_16x2_x16(t_x + COL_S, sq);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        float bm = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (t1 * 16 + i >= last_valid) sq[i] = __float_as_uint(-INFINITY);
          bm = fmaxf(bm, __uint_as_float(sq[i]));
        }
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
        move_reference(bm, j);
        const float mneg = -m_ref * scale_log2;
        uint32_t pq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sq[2 * i]), scale_log2, mneg));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sq[2 * i + 1]), scale_log2, mneg));
          l += p0 + p1;
          pq[i] = pack_bf16x2(p0, p1);
        }
        if (j > 0) {
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st_16x2_x8(t_x + COL_P, pq);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      };
      const int n_full = last_valid < BKV ? nblk - 1 : nblk;
#pragma unroll 1
      for (int j = 0; j < n_full; ++j) key_block(j, std::false_type{});
      if (n_full < nblk) {
        if (last_valid <= 32) short_block(n_full);
        else key_block(n_full, std::true_type{});
      }
      // O / l -> bf16: this thread's half (32 of the 64 dims) of its row
      mbar_wait(pv_done, (nblk - 1) & 1);
      tc_fence_after();
      l += __shfl_xor_sync(0xffffffffu, l, 16);
      uint32_t o[32];
      tmem_ld_16x2_x32(t_x + COL_O, o);
      tmem_ld_wait();
      const int q_row = qt * BQ + r_tile;
      if (q_row < T) {
        const float inv = 1.0f / l;
        __nv_bfloat16* o_dst = out + static_cast<size_t>(row0 + q_row) * d + h * HD + t1 * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          pk.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          pk.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(o_dst + i) = pk;
        }
      }
    }  // warp with valid rows
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SOFTMAX_WARPS + 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace vb200

using namespace vb200;

#ifdef VB200_ATTN_TRACE
extern "C" int vb200_debug_attn_trace(unsigned long long* host_dst, size_t n_words) {
  const size_t n = sizeof(g_attn_trace) / sizeof(unsigned long long);
  cudaError_t e = cudaMemcpyFromSymbol(host_dst, g_attn_trace, (n_words < n ? n_words : n) * sizeof(unsigned long long));
  return e == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int vb200_flash_attn_varlen(void* out_bf16, const void* qkv_bf16, const int32_t* cu_rows,
                                       int32_t B, int32_t max_T, int32_t M, int32_t n_heads,
                                       float scale, vb200_stream_t stream) {
  using namespace attn;
  if (M == 0) return VB200_OK;
  VB_REQUIRE(out_bf16 && qkv_bf16 && cu_rows, "flash_attn: null pointer");
  VB_REQUIRE(B >= 1 && B <= 65535 && n_heads >= 1 && n_heads <= 65535 && max_T >= 1 && M >= 0,
             "flash_attn: bad sizes B=%d heads=%d max_T=%d M=%d", B, n_heads, max_T, M);
  const int d = n_heads * HD;
  CUtensorMap tm;
  int rc = cached_tmap(&tm, VB200_BF16, qkv_bf16, static_cast<uint64_t>(3) * d, M,
                       static_cast<uint64_t>(3) * d * 2, HD, 128);
  if (rc != VB200_OK) return rc;
  dim3 grid((max_T + BQ - 1) / BQ, n_heads, B);
  const float sl2 = scale * 1.4426950408889634f;
  // A grid that fits the SMs once is padded to one CTA per SM (shared memory only one CTA can hold): launched
  // early (PDL), two small CTAs would otherwise be packed onto the first SMs that drain.
  constexpr int SMEM_SOLO = 116 * 1024;
  const bool solo = static_cast<long long>(grid.x) * grid.y * grid.z <= num_sms();
  const int smem_bytes = solo ? SMEM_SOLO : SMEM_BYTES;
  VB_CONFIGURE_SMEM(flash_attn_kernel, SMEM_SOLO);
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
  PdlTag pdl_tag(solo ? 16 : 64);
  VB_CHECK_CUDA(launch_pdl(flash_attn_kernel, grid, dim3(THREADS), smem_bytes, static_cast<cudaStream_t>(stream), 1,
                           tm, static_cast<__nv_bfloat16*>(out_bf16), cu_rows, n_heads, sl2, 0.0f));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}
