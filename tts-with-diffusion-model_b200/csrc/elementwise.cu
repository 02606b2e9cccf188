// HBM-bound kernels of the denoiser: embedding gather-sum (rows E1-E4, T), AdaLN / LayerNorm
// (rows N1, N2) and the response-row gather in front of the classifier (row H1).
// One warp per row; 16-byte vector accesses; fp32 math.
#include "common.cuh"

#ifndef VB200_ADALN_EARLY_PARAMS
#define VB200_ADALN_EARLY_PARAMS 1
#endif

namespace vb200 {

__device__ __forceinline__ void add_bf16x8(float (&acc)[8], const __nv_bfloat16* p) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[2 * i] += __uint_as_float(w[i] << 16);
    acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
  }
}

// x[r] = emb(row r) + pe[pos]   (base.py:427-436)
__global__ void __launch_bounds__(256) embed_gather_kernel(
    float* __restrict__ x_out, const __nv_bfloat16* __restrict__ text_w,
    const __nv_bfloat16* __restrict__ prom_w, const __nv_bfloat16* __restrict__ resp_w,
    const __nv_bfloat16* __restrict__ sep, const __nv_bfloat16* __restrict__ time_w,
    const float* __restrict__ pe, const int32_t* __restrict__ text_ids,
    const int32_t* __restrict__ prom_ids, const int32_t* __restrict__ resp_ids,
    const int32_t* __restrict__ utt, const int32_t* __restrict__ row_utt,
    const int32_t* __restrict__ t_utt, int M, int d, int K, int resp_levels_in) {
  pdl_launch_dependents();
  pdl_wait();                                   // everything below reads / writes activations
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < M; r += gridDim.x * wpb) {
    const int b = row_utt[r];
    const int32_t* u = utt + static_cast<size_t>(b) * VB200_U_STRIDE;
    const int pos = r - u[VB200_U_ROW0];
    const int t_txt = u[VB200_U_TTXT], t_prom = u[VB200_U_TPROM];
    // segment: [text][sep][prom][sep][resp]
    int kind, idx;
    if (pos < t_txt) { kind = 0; idx = pos; }
    else if (pos == t_txt) { kind = 1; idx = 0; }
    else if (pos < t_txt + 1 + t_prom) { kind = 2; idx = pos - t_txt - 1; }
    else if (pos == t_txt + 1 + t_prom) { kind = 1; idx = 0; }
    else { kind = 3; idx = pos - t_txt - t_prom - 2; }

    const __nv_bfloat16* rows[9];
    int n_rows = 0;
    if (kind == 0) {
      rows[n_rows++] = text_w + static_cast<size_t>(text_ids[u[VB200_U_TXT0] + idx]) * d;
    } else if (kind == 1) {
      rows[n_rows++] = sep;
    } else if (kind == 2) {
      const int32_t* ids = prom_ids + static_cast<size_t>(u[VB200_U_PROM0] + idx) * 8;
#pragma unroll
      for (int l = 0; l < 8; ++l)
        rows[n_rows++] = prom_w + (static_cast<size_t>(l) * K + ids[l]) * d;
    } else {
      const int32_t* ids = resp_ids + static_cast<size_t>(u[VB200_U_RESP0] + idx) * resp_levels_in;
      for (int l = 0; l < resp_levels_in; ++l)
        rows[n_rows++] = resp_w + (static_cast<size_t>(l) * K + ids[l]) * d;
      if (time_w) rows[n_rows++] = time_w + static_cast<size_t>(t_utt[b]) * d;
    }
    const float* pe_row = pe + static_cast<size_t>(pos) * d;
    float* out = x_out + static_cast<size_t>(r) * d;
    for (int c = lane * 8; c < d; c += 256) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < n_rows; ++i) add_bf16x8(acc, rows[i] + c);
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(pe_row + c));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(pe_row + c) + 1);
      float4 o0 = make_float4(acc[0] + p0.x, acc[1] + p0.y, acc[2] + p0.z, acc[3] + p0.w);
      float4 o1 = make_float4(acc[4] + p1.x, acc[5] + p1.y, acc[6] + p1.z, acc[7] + p1.w);
      reinterpret_cast<float4*>(out + c)[0] = o0;
      reinterpret_cast<float4*>(out + c)[1] = o1;
    }
  }
}

// Same sum for d = 256 * NV <= 1024, with every source row in a register-resident slot and all
// loads of a row issued before the first use: the generic kernel above walks its (up to 9) source
// rows one dependent 16-byte load after the other, which at one utterance at a time (M ~ 1000, one
// row per warp) is pure latency — 24 us of a 0.87 ms denoise step.
template <int NV>
__global__ void __launch_bounds__(256) embed_gather_rows_kernel(
    float* __restrict__ x_out, const __nv_bfloat16* __restrict__ text_w,
    const __nv_bfloat16* __restrict__ prom_w, const __nv_bfloat16* __restrict__ resp_w,
    const __nv_bfloat16* __restrict__ sep, const __nv_bfloat16* __restrict__ time_w,
    const float* __restrict__ pe, const int32_t* __restrict__ text_ids,
    const int32_t* __restrict__ prom_ids, const int32_t* __restrict__ resp_ids,
    const int32_t* __restrict__ utt, const int32_t* __restrict__ row_utt,
    const int32_t* __restrict__ t_utt, int M, int K, int resp_levels_in) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int d = NV * 256;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < M; r += gridDim.x * wpb) {
    const int b = row_utt[r];
    const int4 u0 = __ldg(reinterpret_cast<const int4*>(utt + static_cast<size_t>(b) * VB200_U_STRIDE));
    const int4 u1 = __ldg(reinterpret_cast<const int4*>(utt + static_cast<size_t>(b) * VB200_U_STRIDE) + 1);
    const int pos = r - u0.x, t_txt = u0.y, t_prom = u0.z;        // ROW0, TTXT, TPROM | TXT0, PROM0, RESP0
    const __nv_bfloat16* src[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) src[i] = nullptr;
    if (pos < t_txt) {
      src[0] = text_w + static_cast<size_t>(text_ids[u1.x + pos]) * d;
    } else if (pos == t_txt || pos == t_txt + 1 + t_prom) {
      src[0] = sep;
    } else if (pos < t_txt + 1 + t_prom) {
      const int32_t* ids = prom_ids + static_cast<size_t>(u1.y + pos - t_txt - 1) * 8;
      const int4 i0 = __ldg(reinterpret_cast<const int4*>(ids)), i1 = __ldg(reinterpret_cast<const int4*>(ids) + 1);
      const int id[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
      for (int l = 0; l < 8; ++l) src[l] = prom_w + (static_cast<size_t>(l) * K + id[l]) * d;
    } else {
      const int32_t* ids = resp_ids + static_cast<size_t>(u1.z + pos - t_txt - t_prom - 2) * resp_levels_in;
#pragma unroll
      for (int l = 0; l < 8; ++l)
        if (l < resp_levels_in) src[l] = resp_w + (static_cast<size_t>(l) * K + ids[l]) * d;
      if (time_w) src[8] = time_w + static_cast<size_t>(t_utt[b]) * d;
    }
    float acc[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(pe + static_cast<size_t>(pos) * d + v * 256 + lane * 8));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(pe + static_cast<size_t>(pos) * d + v * 256 + lane * 8) + 1);
      acc[v][0] = p0.x; acc[v][1] = p0.y; acc[v][2] = p0.z; acc[v][3] = p0.w;
      acc[v][4] = p1.x; acc[v][5] = p1.y; acc[v][6] = p1.z; acc[v][7] = p1.w;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (src[i] != nullptr) {                       // warp-uniform
        uint4 raw[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) raw[v] = __ldg(reinterpret_cast<const uint4*>(src[i] + v * 256 + lane * 8));
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const uint32_t w[4] = {raw[v].x, raw[v].y, raw[v].z, raw[v].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[v][2 * q] += __uint_as_float(w[q] << 16);
            acc[v][2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
          }
        }
      }
    }
    float* out = x_out + static_cast<size_t>(r) * d;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      reinterpret_cast<float4*>(out + v * 256 + lane * 8)[0] = make_float4(acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
      reinterpret_cast<float4*>(out + v * 256 + lane * 8)[1] = make_float4(acc[v][4], acc[v][5], acc[v][6], acc[v][7]);
    }
  }
}

// Row statistics + normalisation.  MODE 0: AdaLN (base.py:145-158), MODE 1: affine LayerNorm.
// d <= 32 * 8 * MAXV elements are kept in registers between the two passes.
template <int MODE>
__global__ void __launch_bounds__(256) norm_kernel(
    __nv_bfloat16* __restrict__ out, const float* __restrict__ x, const float* __restrict__ p0,
    const float* __restrict__ p1, const int32_t* __restrict__ level_utt,
    const int32_t* __restrict__ row_utt, int M, int d, float eps, float k, float c, int out_f16) {
  pdl_launch_dependents();
  pdl_wait();                                   // everything below reads / writes activations
  constexpr int MAXV = 8;  // d <= 2048
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nv = d / 256 + ((d % 256) ? 1 : 0);
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < M; r += gridDim.x * wpb) {
    const float* xr = x + static_cast<size_t>(r) * d;
    float v[MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int col = i * 256 + lane * 8;
      if (i < nv && col < d) {
        const float4 a = *reinterpret_cast<const float4*>(xr + col);
        const float4 bb = *reinterpret_cast<const float4*>(xr + col + 4);
        v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
        v[i][4] = bb.x; v[i][5] = bb.y; v[i][6] = bb.z; v[i][7] = bb.w;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[i][e];
      }
    }
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int col = i * 256 + lane * 8;
      if (i < nv && col < d) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float dv = v[i][e] - mean; q += dv * dv; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
    const float* g;   // gamma (AdaLN: exp(log gamma) row) / LN weight
    const float* bt;  // beta
    if (MODE == 0) {
      const int l = level_utt[row_utt[r]];
      g = p0 + static_cast<size_t>(l) * 2 * d;
      bt = g + d;
    } else {
      g = p0; bt = p1;
    }
    __nv_bfloat16* orow = out + static_cast<size_t>(r) * d;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int col = i * 256 + lane * 8;
      if (i < nv && col < d) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + col));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + col + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bt + col));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bt + col + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float h = (v[i][e] - mean) * rstd;
          if (MODE == 0) h = c * (1.f - k * h) * h;
          y[e] = gg[e] * h + bb[e];
        }
        uint4 o;
        o.x = pack_act16x2(y[0], y[1], out_f16); o.y = pack_act16x2(y[2], y[3], out_f16);
        o.z = pack_act16x2(y[4], y[5], out_f16); o.w = pack_act16x2(y[6], y[7], out_f16);
        *reinterpret_cast<uint4*>(orow + col) = o;
      }
    }
  }
}

// AdaLN for d = NV*256 with the [gamma | beta] row of the current timestep cached in registers:
// each warp walks a CONTIGUOUS range of rows (same utterance => same table row), so the table is
// re-read only at utterance boundaries and the steady state moves 4d bytes in, 2d bytes out.
template <int NV>
__global__ void __launch_bounds__(256) adaln_rows_kernel(
    __nv_bfloat16* __restrict__ out, const float* __restrict__ x, const float* __restrict__ table,
    const int32_t* __restrict__ level_utt, const int32_t* __restrict__ row_utt, int M, int rows_per_warp,
    float eps, float k, float c, int out_f16) {
  pdl_launch_dependents();
  constexpr int d = NV * 256;
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int r_begin = warp_global * rows_per_warp;
  const int r_end = min(r_begin + rows_per_warp, M);
  float g[NV][8], bt[NV][8];
  int cached = -1;
#if VB200_ADALN_EARLY_PARAMS
  // The [gamma | beta] row of the first row's level is fetched BEFORE the wait on the previous kernel (row_utt ->
  // level_utt -> table: three dependent loads): the layout and the levels / timesteps are written once per
  // denoise step, never by the kernel right in front of an AdaLN.
  if (r_begin < r_end) {
    cached = level_utt[row_utt[r_begin]];
    adaln_load_params<NV>(table, cached, lane, g, bt);
  }
#endif
  pdl_wait();                                   // everything below reads / writes activations
  for (int r = r_begin; r < r_end; ++r) {
    const float* xr = x + static_cast<size_t>(r) * d;
    float v[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(xr + i * 256 + lane * 8));
      const float4 b = __ldcs(reinterpret_cast<const float4*>(xr + i * 256 + lane * 8 + 4));
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
      v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
    }
    const int lvl = level_utt[row_utt[r]];
    if (lvl != cached) {                               // warp-uniform
      adaln_load_params<NV>(table, lvl, lane, g, bt);
      cached = lvl;
    }
    adaln_row_finish<NV>(v, g, bt, out + static_cast<size_t>(r) * d, lane, eps, k, c, out_f16 != 0);
  }
}

__global__ void __launch_bounds__(256) gather_rows_bf16_kernel(
    __nv_bfloat16* __restrict__ out, const float* __restrict__ x,
    const int32_t* __restrict__ row_index, int n_rows, int d, int out_f16) {
  pdl_launch_dependents();
  pdl_wait();                                   // everything below reads / writes activations
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rows; r += gridDim.x * wpb) {
    const float* xr = x + static_cast<size_t>(row_index[r]) * d;
    __nv_bfloat16* orow = out + static_cast<size_t>(r) * d;
    for (int c = lane * 8; c < d; c += 256) {
      const float4 a = *reinterpret_cast<const float4*>(xr + c);
      const float4 b = *reinterpret_cast<const float4*>(xr + c + 4);
      uint4 o;
      o.x = pack_act16x2(a.x, a.y, out_f16); o.y = pack_act16x2(a.z, a.w, out_f16);
      o.z = pack_act16x2(b.x, b.y, out_f16); o.w = pack_act16x2(b.z, b.w, out_f16);
      *reinterpret_cast<uint4*>(orow + c) = o;
    }
  }
}

// codes (sum T_resp, L) int32 -> (B, L, T_max) int64, one thread per (utterance, frame): the L codes of
// a frame are one contiguous read, every level's row of the output is written coalesced along t.
__global__ void __launch_bounds__(256) codes_to_bqt_kernel(
    int64_t* __restrict__ out, const int32_t* __restrict__ codes, const int32_t* __restrict__ utt,
    int n_levels, int T_max, int64_t pad) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T_max) return;
  const int32_t* u = utt + b * VB200_U_STRIDE;
  const bool valid = t < u[VB200_U_TRESP];
  const int32_t* src = codes + (static_cast<size_t>(u[VB200_U_RESP0]) + t) * n_levels;
  int64_t* dst = out + static_cast<size_t>(b) * n_levels * T_max + t;
  for (int l = 0; l < n_levels; ++l) dst[static_cast<size_t>(l) * T_max] = valid ? static_cast<int64_t>(src[l]) : pad;
}

static int row_grid(int rows, int wpb) {
  int grid = (rows + wpb - 1) / wpb;
  const int cap = num_sms() * 16;
  return grid > cap ? cap : (grid < 1 ? 1 : grid);
}

}  // namespace vb200

using namespace vb200;

extern "C" int vb200_embed_gather(float* x_out, const void* text_w, const void* prom_w,
                                  const void* resp_w, const void* sep, const void* time_w,
                                  const float* pe, const int32_t* text_ids,
                                  const int32_t* prom_ids, const int32_t* resp_ids,
                                  const int32_t* utt, const int32_t* row_utt,
                                  const int32_t* t_utt, int32_t M, int32_t d, int32_t K,
                                  int32_t resp_levels_in, vb200_stream_t stream) {
  if (M <= 0) return VB200_OK;
  VB_REQUIRE(x_out && text_w && prom_w && resp_w && sep && pe && utt && row_utt,
             "embed_gather: null pointer");
  VB_REQUIRE(!time_w || t_utt, "embed_gather: time_w given without t_utt");
  VB_REQUIRE(d > 0 && d % 8 == 0, "embed_gather: d=%d must be a positive multiple of 8", d);
  VB_REQUIRE(resp_levels_in >= 1 && resp_levels_in <= 8, "embed_gather: resp_levels_in=%d not in 1..8", resp_levels_in);
  if (M <= 0) return VB200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16 *tw = static_cast<const __nv_bfloat16*>(text_w), *pw = static_cast<const __nv_bfloat16*>(prom_w),
                      *rw = static_cast<const __nv_bfloat16*>(resp_w), *sp = static_cast<const __nv_bfloat16*>(sep),
                      *ti = static_cast<const __nv_bfloat16*>(time_w);
  const bool rows_form = d % 256 == 0 && d <= 1024 && resp_levels_in <= 8 &&
                         (reinterpret_cast<uintptr_t>(utt) & 15) == 0 && (reinterpret_cast<uintptr_t>(prom_ids) & 15) == 0;
#define VB_EMB(NV)                                                                                          \
  VB_CHECK_CUDA(launch_pdl(embed_gather_rows_kernel<NV>, dim3(row_grid(M, 8)), dim3(256), 0, st, 1, x_out, tw, pw, rw, \
                           sp, ti, pe, text_ids, prom_ids, resp_ids, utt, row_utt, t_utt, M, K, resp_levels_in))
  if (rows_form) {
    switch (d / 256) {
      case 1: VB_EMB(1); break;
      case 2: VB_EMB(2); break;
      case 3: VB_EMB(3); break;
      default: VB_EMB(4); break;
    }
  } else {
    VB_CHECK_CUDA(launch_pdl(embed_gather_kernel, dim3(row_grid(M, 8)), dim3(256), 0, st, 1, x_out, tw, pw, rw, sp, ti,
                             pe, text_ids, prom_ids, resp_ids, utt, row_utt, t_utt, M, d, K, resp_levels_in));
  }
#undef VB_EMB
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}

extern "C" int vb200_adaln(void* out_bf16, vb200_dtype out_dtype, const float* x, const float* table,
                           const int32_t* level_utt, const int32_t* row_utt, int32_t M, int32_t d,
                           float eps, float k, float c, vb200_stream_t stream) {
  VB_REQUIRE(out_dtype == VB200_BF16 || out_dtype == VB200_F16, "adaln: out_dtype must be bf16 or f16");
  const int f16 = out_dtype == VB200_F16;
  if (M <= 0) return VB200_OK;
  VB_REQUIRE(out_bf16 && x && table && level_utt && row_utt, "adaln: null pointer");
  VB_REQUIRE(d > 0 && d % 8 == 0 && d <= 2048, "adaln: d=%d must be a multiple of 8, <= 2048", d);
  if (M <= 0) return VB200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out_bf16);
  if (d % 256 == 0 && d <= 1024) {
    // contiguous row ranges per warp; enough warps for ~4 resident blocks per SM.  Small M (one
    // utterance at a time) gets one row per warp: there the kernel is latency-, not bandwidth-bound.
    const int total_warps = num_sms() * 4 * 8;
    const int rpw = (M + total_warps - 1) / total_warps;
    const int warps = (M + rpw - 1) / rpw;
    const int grid = (warps + 7) / 8;
    // one block per SM for a grid that fits the SMs once (see flash_attn: no packing under PDL)
    constexpr int SMEM_SOLO = 116 * 1024;
    const bool solo = grid <= num_sms() && d == 1024;
    PdlTag pdl_tag(solo ? 2 : 128);
    if (solo) {
      VB_CONFIGURE_SMEM(adaln_rows_kernel<4>, SMEM_SOLO);
      VB_CHECK_CUDA(launch_pdl(adaln_rows_kernel<4>, dim3(grid), dim3(256), SMEM_SOLO, st, 1, o, x, table, level_utt, row_utt, M, rpw, eps, k, c, f16));
      VB_CHECK_CUDA(cudaGetLastError());
      return VB200_OK;
    }
    switch (d / 256) {
      case 1: VB_CHECK_CUDA(launch_pdl(adaln_rows_kernel<1>, dim3(grid), dim3(256), 0, st, 1, o, x, table, level_utt, row_utt, M, rpw, eps, k, c, f16)); break;
      case 2: VB_CHECK_CUDA(launch_pdl(adaln_rows_kernel<2>, dim3(grid), dim3(256), 0, st, 1, o, x, table, level_utt, row_utt, M, rpw, eps, k, c, f16)); break;
      case 3: VB_CHECK_CUDA(launch_pdl(adaln_rows_kernel<3>, dim3(grid), dim3(256), 0, st, 1, o, x, table, level_utt, row_utt, M, rpw, eps, k, c, f16)); break;
      default: VB_CHECK_CUDA(launch_pdl(adaln_rows_kernel<4>, dim3(grid), dim3(256), 0, st, 1, o, x, table, level_utt, row_utt, M, rpw, eps, k, c, f16)); break;
    }
  } else {
    VB_CHECK_CUDA(launch_pdl(norm_kernel<0>, dim3(row_grid(M, 8)), dim3(256), 0, st, 1, o, x, table,
                             static_cast<const float*>(nullptr), level_utt, row_utt, M, d, eps, k, c, f16));
  }
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}

extern "C" int vb200_layernorm(void* out_bf16, vb200_dtype out_dtype, const float* x, const float* weight,
                               const float* bias, int32_t M, int32_t d, float eps,
                               vb200_stream_t stream) {
  VB_REQUIRE(out_dtype == VB200_BF16 || out_dtype == VB200_F16, "layernorm: out_dtype must be bf16 or f16");
  const int f16 = out_dtype == VB200_F16;
  if (M <= 0) return VB200_OK;
  VB_REQUIRE(out_bf16 && x && weight && bias, "layernorm: null pointer");
  VB_REQUIRE(d > 0 && d % 8 == 0 && d <= 2048, "layernorm: d=%d must be a multiple of 8, <= 2048", d);
  if (M <= 0) return VB200_OK;
  VB_CHECK_CUDA(launch_pdl(norm_kernel<1>, dim3(row_grid(M, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
      static_cast<__nv_bfloat16*>(out_bf16), x, weight, bias, static_cast<const int32_t*>(nullptr),
      static_cast<const int32_t*>(nullptr), M, d, eps, 0.f, 0.f, f16));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}

extern "C" int vb200_gather_rows_bf16(void* out_bf16, vb200_dtype out_dtype, const float* x,
                                      const int32_t* row_index, int32_t n_rows, int32_t d,
                                      vb200_stream_t stream) {
  VB_REQUIRE(out_dtype == VB200_BF16 || out_dtype == VB200_F16, "gather_rows: out_dtype must be bf16 or f16");
  const int f16 = out_dtype == VB200_F16;
  if (n_rows <= 0) return VB200_OK;
  VB_REQUIRE(out_bf16 && x && row_index, "gather_rows: null pointer");
  VB_REQUIRE(d > 0 && d % 8 == 0, "gather_rows: d=%d must be a multiple of 8", d);
  if (n_rows <= 0) return VB200_OK;
  VB_CHECK_CUDA(launch_pdl(gather_rows_bf16_kernel, dim3(row_grid(n_rows, 8)), dim3(256), 0,
                           static_cast<cudaStream_t>(stream), 1, static_cast<__nv_bfloat16*>(out_bf16), x, row_index, n_rows, d, f16));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}

extern "C" int vb200_codes_to_bqt(int64_t* out, const int32_t* codes, const int32_t* utt, int32_t B,
                                  int32_t n_levels, int32_t T_max, int64_t pad, vb200_stream_t stream) {
  VB_REQUIRE(B >= 0 && T_max >= 0 && n_levels > 0, "codes_to_bqt: bad sizes B=%d n_levels=%d T_max=%d", B, n_levels, T_max);
  if (B == 0 || T_max == 0) return VB200_OK;
  VB_REQUIRE(out && codes && utt, "codes_to_bqt: null pointer");
  VB_REQUIRE(B <= 65535, "codes_to_bqt: B=%d exceeds the grid's y extent", B);
  VB_CHECK_CUDA(launch_pdl(codes_to_bqt_kernel, dim3((T_max + 255) / 256, B), dim3(256), 0,
                           static_cast<cudaStream_t>(stream), 1, out, codes, utt, n_levels, T_max, pad));
  VB_CHECK_CUDA(cudaGetLastError());
  return VB200_OK;
}
