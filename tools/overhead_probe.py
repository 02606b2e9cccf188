"""Where the fixed cost of the batch-1 (one utterance, M = 1027) kernels goes: each kernel is replayed 48 times
back to back inside one CUDA graph, at its real shape and at a shape with (almost) no work, so that
    real - empty = the part that scales with the data,   empty = launch gap + prologue + epilogue.
    python tools/overhead_probe.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
from vall_e.b200 import lib as L  # noqa: E402

L.load()
dev, M, reps = "cuda", 1027, 48
torch.manual_seed(0)


def in_graph(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3 / reps


def gemm(N, K, epi, dt):
    A = (torch.randn(M, K, device=dev) * 0.1).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(N, device=dev)
    out = torch.zeros(M, N, dtype=dt, device=dev)
    return lambda: L.gemm_bf16(out, A, W, None if epi == L.EPI_NONE else bias, out if epi == L.EPI_BIAS_RESIDUAL else None, epi)


for name, N, K, epi, dt in (("qkv", 3072, 1024, L.EPI_NONE, torch.bfloat16), ("to_out", 1024, 1024, L.EPI_BIAS_RESIDUAL, torch.float32),
                            ("ffn1", 4096, 1024, L.EPI_BIAS_GELU, torch.bfloat16), ("ffn2", 1024, 4096, L.EPI_BIAS_RESIDUAL, torch.float32)):
    real = in_graph(gemm(N, K, epi, dt))
    short = in_graph(gemm(N, 64, epi, dt))
    print(f"gemm {name:7s} N={N} K={K}: {real:6.2f} us   with K=64 (one k-block): {short:6.2f} us", flush=True)

# attention: one utterance of 1027 rows vs one whose every query tile sees a single key block
for T in (1027, 128, 16):
    d = 1024
    qkv = torch.randn(T, 3 * d, device=dev).bfloat16()
    cu = torch.tensor([0, T], dtype=torch.int32, device=dev)
    out = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
    print(f"attention T={T}: {in_graph(lambda: L.flash_attn_varlen(out, qkv, cu, T, 16, 0.125)):6.2f} us", flush=True)

# AdaLN over the utterance's rows, and over 8 rows
for rows in (M, 8):
    d = 1024
    x = torch.randn(rows, d, device=dev)
    table = torch.randn(51, 2 * d, device=dev)
    lvl = torch.zeros(1, dtype=torch.int32, device=dev)
    row_utt = torch.zeros(rows, dtype=torch.int32, device=dev)
    h = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
    print(f"adaln rows={rows}: {in_graph(lambda: L.adaln(h, x, table, lvl, row_utt)):6.2f} us", flush=True)

# the cheapest kernel there is: one block decrementing the timesteps
t = torch.full((1,), 1000000, dtype=torch.int32, device=dev)
print(f"step_timesteps (1 block): {in_graph(lambda: L.step_timesteps(t, -1)):6.2f} us", flush=True)
