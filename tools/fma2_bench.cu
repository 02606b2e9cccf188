// Micro-benchmark: issue rate per SM sub-partition of the fp32 instructions the attention softmax
// is made of — FFMA (register and uniform-operand forms), the packed FFMA2 / FADD2 (f32x2),
// FMNMX3, F2FP (bf16x2 pack) and LEA — at 1, 2 and 4 warps per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void kern(float* out, long long* cyc, int iters, float a, float b) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3f + i;
  uint64_t* v2 = reinterpret_cast<uint64_t*>(v);
  uint64_t ab, bb;
  {
    float2 t = make_float2(a, a), u = make_float2(b, b);
    ab = *reinterpret_cast<uint64_t*>(&t);
    bb = *reinterpret_cast<uint64_t*>(&u);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b));
    } else if (OP == 1) {      // 8 packed = 16 floats
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v2[i]) : "l"(ab), "l"(bb));
    } else if (OP == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v2[i]) : "l"(ab));
    } else if (OP == 3) {      // 3-input max, 16 results
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b));
    } else if (OP == 4) {      // pack pairs: 8 F2FP + nothing else
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t p;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        v[2 * i] = __uint_as_float(p);
      }
    } else if (OP == 5) {      // fadd scalar
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(a));
    } else if (OP == 7) {      // packed half exponentials: 8 instructions = 16 results
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t* h = reinterpret_cast<uint32_t*>(&v[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(*h));
      }
    } else if (OP == 8) {      // f32 pair -> f16x2 pack
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t p;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        v[2 * i] = __uint_as_float(p);
      }
    } else if (OP == 9) {      // half2 add
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t* h = reinterpret_cast<uint32_t*>(&v[i]);
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(*h) : "r"(0x3c003c00u));
      }
    } else if (OP == 6) {      // mixed: per 4 floats  1 FFMA2x2 + MUFU x3 ... the real ratio: see attn kernel
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v2[i]) : "l"(ab), "l"(bb));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[2 * i]));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int inst_per_iter, float* o, long long* c) {
  for (int warps : {4, 8, 16}) {
    const int iters = 2048;
    for (int rep = 0; rep < 2; ++rep) {
      kern<OP><<<148, warps * 32>>>(o, c, iters, 1.0001f, 1e-7f);
      cudaDeviceSynchronize();
    }
    long long cyc;
    cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
    const double per_smsp_inst = double(iters) * inst_per_iter * (warps / 4.0);
    printf("%-28s %2d warps/SM: %.2f cycles per warp instruction per scheduler\n", name, warps, cyc / per_smsp_inst);
  }
}

int main() {
  float* o;
  long long* c;
  cudaMalloc(&o, 148 * 1024 * 4);
  cudaMalloc(&c, 8);
  run<0>("FFMA", 16, o, c);
  run<1>("FFMA2 (f32x2)", 8, o, c);
  run<2>("FADD2 (f32x2)", 8, o, c);
  run<3>("FMNMX3", 16, o, c);
  run<4>("F2FP.BF16 pack", 8, o, c);
  run<5>("FADD", 16, o, c);
  run<6>("FFMA2 + MUFU.EX2 pair", 16, o, c);
  run<7>("MUFU.EX2 f16x2", 8, o, c);
  run<8>("F2FP.F16 pack", 8, o, c);
  run<9>("HADD2", 16, o, c);
  return 0;
}
