"""Quick kernel timing probe (CUDA events) for bring-up: GEMM shapes of the full denoiser,
attention at T=1027/2527, posterior kernel.  Prints TFLOP/s and GB/s."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
from vall_e.b200 import lib as L  # noqa: E402

L.load()
dev = "cuda"


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def gemm(M, N, K, epi, dt):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, dtype=dt, device=dev)
    res = out if epi == L.EPI_BIAS_RESIDUAL else None
    ms = timeit(lambda: L.gemm_bf16(out, A, W, bias, res, epi))
    ref = timeit(lambda: torch.matmul(A, W.t()))
    print(f"gemm M={M} N={N} K={K} epi={epi}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TFLOP/s | cuBLAS {ref:.3f} ms {2*M*N*K/ref/1e9:.0f}", flush=True)


def attn(lens, heads, variant):
    M = sum(lens)
    d = heads * 64
    qkv = torch.randn(M, 3 * d, device=dev).bfloat16()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    out = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: L.flash_attn_varlen(out, qkv, cu, max(lens), heads, 0.125, variant=variant))
    fl = sum(4 * T * T * d for T in lens)
    print(f"attn[{variant}] B={len(lens)} T={lens[0]} heads={heads}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for B in (1, 32):
        M = 1027 * B
        gemm(M, 3072, 1024, L.EPI_NONE, torch.bfloat16)
        gemm(M, 1024, 1024, L.EPI_BIAS_RESIDUAL, torch.float32)
        gemm(M, 4096, 1024, L.EPI_BIAS_GELU, torch.bfloat16)
        gemm(M, 1024, 4096, L.EPI_BIAS_RESIDUAL, torch.float32)
    gemm(750 * 32, 8192, 1024, L.EPI_BIAS, torch.float16)
    for v in ("tmem",):
        try:
            attn([1027] * 32, 16, v)
            attn([2527] * 8, 16, v)
        except Exception as e:  # bring-up: keep going
            print("attn", v, "failed:", e)
